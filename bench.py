#!/usr/bin/env python
"""bench.py — variants/sec of the per-variant score-test hot path (BASELINE.json metric).

Workload (N = 1 and every N of the scaling run): BASELINE.json configs[1] — continuous trait, sparse GRM
(related samples), n = 100,000 samples, 10 M variants on 8 GPUs — sharded by variants: every rank holds
10 M / 8 = 1.25 M variants as 2-bit packed dosage columns resident in HBM (31 GB, >> L2), plus a replica of the
null model broadcast once over NCCL.  One step = one pass of the fused path (flip/impute/AF -> G'[r | Sigma_iX] ->
sparse quadratic form -> per-variant statistics) over the rank's whole shard.  `scaling` = weak.

One JSON line on stdout (rank 0): value (device-resident), e2e (host FP64 dosage chunks through the C ABI, H2D/D2H
inside the timed region), roofline (dominant kernel, live CUDA-event time), cpu_baseline (the reference's own
sources, oracle/_ref, on a bounded sample on this box's host cores), clocks, gpu_launches.

`--impl reference` times the reference's CPU implementation (oracle/_ref: /root/reference/src/*.cpp compiled
against stand-in headers; falls back to the oracle port) on the R driver's call sequence for this workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SAMPLES = 100_000
Q_COVARIATES = 13
VARIANTS_PER_GPU = 1_250_000           # 10 M variants / 8 GPUs (BASELINE.json configs[1])
CHUNK = 5_000                          # R driver's subset_variants_num (R/Individual_Analysis.R:45)
METRIC = "variants/sec of score tests (n=100k samples)"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception as e:  # noqa: BLE001
            log("clock sampler unavailable:", e)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t_begin=None, t_end=None):
        """rows that arrived inside the timed region [t_begin, t_end] (the sampler is started before the warm-up so that
        nvidia-smi is already running when the region begins); if the region was shorter than one sample period, the
        rows of the warm-up steps — the same kernels under the same load — are used and counted as such"""
        if self.proc:
            self.proc.terminate()
        inside = [r for t, r in self.rows if t_begin is None or (t_begin <= t <= t_end + 0.05)]
        source = "timed region"
        if not inside:
            inside = [r for _, r in self.rows]
            source = "warm-up and timed region"
        rows = inside
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        self.source = source
        for r in rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "sampled_during": self.source}


# ------------------------------------------------------------------------------------------------ reference CPU path
def reference_chunk(ref, G, nm, mac_cutoff=20):
    """The reference's native calls for one chunk, in the R driver's order (R/Individual_Analysis.R:189-443):
    matrix_flip_mean -> common variants: Individual_Score_Test (dense G) -> rare: Individual_Score_Test_sp."""
    import scipy.sparse as sp
    n = G.shape[0]
    t0 = time.perf_counter()
    fl = ref.matrix_flip_mean(G)
    t_native = time.perf_counter() - t0
    maf = fl["MAF"]
    common = maf >= 0.05
    rare = (maf > (mac_cutoff - 0.5) / n / 2) & (maf < 0.05)
    tested = int(common.sum() + rare.sum())
    if common.any():
        Gc = np.asfortranarray(fl["Geno"][:, common])
        t0 = time.perf_counter()
        ref.Individual_Score_Test(Gc, nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
        t_native += time.perf_counter() - t0
    if rare.sum() >= 2:
        Gr = sp.csc_matrix(fl["Geno"][:, rare])          # as(Geno_rare, "dgCMatrix"): R-side, not timed
        t0 = time.perf_counter()
        ref.Individual_Score_Test_sp(Gr, nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
        t_native += time.perf_counter() - t0
    return t_native, tested


def load_reference():
    from oracle.pyoracle import Oracle, have_reference
    if have_reference():
        ref = Oracle("reference")
        blas = ref.set_blas()
        return ref, "reference", blas
    return Oracle("port"), "port", False


def host_chunk(n, p, first_variant, seed):
    """FP64 `$dosage` block on the host from the same hash RNG as the device generator (NaN = missing)."""
    from staarpipeline_b200 import synth
    prm = synth.variant_params(n, first_variant, p, seed)
    out = np.empty((n, p), order="F")
    step = 256
    for a in range(0, p, step):
        b = min(p, a + step)
        sub = synth.VariantParams(prm.thr_hom[a:b], prm.thr_het[a:b], prm.thr_miss[a:b], prm.store_alt[a:b], prm.alt_freq[a:b])
        out[:, a:b] = synth.codes_to_dosage(synth.dosage_codes(n, first_variant + a, b - a, seed, params=sub))
    return out


def run_reference(args):
    """--impl reference: rank 0 only; each step = one R-driver chunk on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from staarpipeline_b200 import synth
    ref, kind, blas = load_reference()
    cores = os.cpu_count() or 1
    n, p = args.n, args.ref_chunk
    nm = synth.nullmodel_sparse_grm(n, seed=synth.BASE_SEED, q=Q_COVARIATES, shuffle=True)
    G = host_chunk(n, p, 0, synth.BASE_SEED)
    times, tested = [], 0
    for s in range(args.warmup + args.steps):
        t, tested = reference_chunk(ref, G, nm)
        if s >= args.warmup:
            times.append(t)
        log(f"[reference] step {s}: {t:.2f} s native for {p} variants ({tested} tested)")
    total = float(np.sum(times))
    value = p * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "variants/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, p),
        "cpu_baseline": {"value": value, "unit": "variants/s", "cores": cores if blas else 1, "kind": kind,
                         "sample": f"{args.steps} chunks of {p} variants x {n} samples; native calls of the R driver "
                                   f"(matrix_flip_mean, Individual_Score_Test, Individual_Score_Test_sp); dense GEMM on "
                                   f"{'OpenBLAS, ' + str(cores) + ' threads' if blas else 'in-order loops'}; "
                                   "Armadillo/Rmath are stand-ins (oracle/ref_stub)"},
        "e2e": {"value": value, "unit": "variants/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(json.dumps(line))


def workload_config(args, variants_per_step):
    return {"workload": "BASELINE configs[1]: Individual_Analysis continuous trait, sparse GRM (related samples), "
                        f"n={args.n}, q={Q_COVARIATES}; per-GPU shard of the 10M-variant job",
            "n_samples": args.n, "variants_per_gpu_per_step": variants_per_step, "genotype_bits": 2,
            "impute": "mean", "parallelism": f"variant-sharded x{args.gpus}, null model replicated",
            "l2": "inputs (31 GB/GPU) larger than L2; no flush needed"}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from staarpipeline_b200 import api, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    n, V = args.n, args.variants_per_gpu
    seed = synth.BASE_SEED

    # ---- null model: built on rank 0, replicated by NCCL broadcast over NVLink once per run
    from staarpipeline_b200 import shard
    t0 = time.perf_counter()
    parts = None
    if rank == 0:
        nm = synth.nullmodel_sparse_grm(n, seed=seed, q=Q_COVARIATES, shuffle=True)
        parts = shard.pack_nullmodel(nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
    parts = shard.broadcast_nullmodel(parts, rank, world, dev)
    Sigma_i, Sigma_iX, cov, residuals = shard.unpack_nullmodel(parts)
    ctx = api.Context(local)
    ctx.set_nullmodel_sparse(Sigma_i, Sigma_iX, cov, residuals)
    log(f"[rank {rank}] null model ready in {time.perf_counter() - t0:.1f} s (nnz(Sigma_i) = {Sigma_i.nnz})")

    # ---- this rank's shard of packed dosage columns, generated on the device
    t0 = time.perf_counter()
    stride = api.packed_stride(n)
    order = ctx.nullmodel_sample_order()           # resident columns are in the null-model sample order
    d_order = torch.from_numpy(order).to(dev)
    first = rank * V                       # weak scaling: rank r owns variants [r*V, (r+1)*V) of the N*V-variant job
    packed = torch.empty((V, stride), dtype=torch.uint8, device=dev)
    gen = 125_000
    for a in range(0, V, gen):
        b = min(V, a + gen)
        prm = synth.variant_params(n, first + a, b - a, seed)
        th = torch.from_numpy(prm.thr_hom.view(np.int64)).to(dev)
        te = torch.from_numpy(prm.thr_het.view(np.int64)).to(dev)
        tm = torch.from_numpy(prm.thr_miss.view(np.int64)).to(dev)
        sa = torch.from_numpy(prm.store_alt).to(dev)
        torch.cuda.synchronize()
        ctx.synth_packed_device(packed[a:b].data_ptr(), stride, n, first + a, b - a, seed, th.data_ptr(), te.data_ptr(),
                                tm.data_ptr(), sa.data_ptr(), d_order.data_ptr())
    outs = torch.zeros((7, V), dtype=torch.float64, device=dev)
    optrs = [outs[i].data_ptr() for i in range(7)]
    torch.cuda.synchronize()
    log(f"[rank {rank}] {V} packed variants x {n} samples resident ({V * stride / 1e9:.1f} GB) in {time.perf_counter() - t0:.1f} s")

    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)

    def step():
        ctx.packed_score_device(packed.data_ptr(), stride, n, V, api.IMPUTE_MEAN, optrs)

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    t_begin = time.perf_counter()
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(args.steps):
        step()
    e1.record(ext)
    barrier()
    launches = ctx.launch_count - launches0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop(t_begin, time.perf_counter()) if rank == 0 else None
    ms_total = float(ms.item())
    value = world * V * args.steps / (ms_total / 1e3)
    # rank 0 owns variants [0, V) at every N: the hash of its seven output arrays must be the same in the N = 1, 2, 4, 8
    # lines of the scaling run (no floating-point atomics, fixed reduction trees)
    outputs_sha256 = outputs_sha256_arrays = None
    if rank == 0:
        import hashlib
        outs_host = outs.cpu().numpy()
        outputs_sha256 = hashlib.sha256(outs_host.tobytes()).hexdigest()
        outputs_sha256_arrays = {k: hashlib.sha256(np.ascontiguousarray(outs_host[i]).tobytes()).hexdigest()[:16]
                                 for i, k in enumerate(("AF", "MAF", "Score", "Score_se", "pvalue_log", "Est", "Est_se"))}
        if os.environ.get("STAAR_B200_BENCH_DUMP"):
            np.save(os.environ["STAAR_B200_BENCH_DUMP"], outs_host)
        del outs_host

    # ---- per-kernel device times (CUDA events on the launching stream) for the roofline figure
    ctx.set_profiling(True)
    stage = {"scan": [], "contract": [], "finalize": []}
    for _ in range(3):
        step()
        for k, v in ctx.last_stage_ms().items():
            stage[k].append(v)
    ctx.set_profiling(False)
    stage_ms = {k: float(np.mean(v)) for k, v in stage.items()}
    dominant = max(stage_ms, key=stage_ms.get)
    bytes_per_variant = n * 2 / 8 + 56                 # n*E/8 + 7 doubles out, E = 2 bits
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = V * bytes_per_variant / (stage_ms[dominant] / 1e3) / 1e9
    traffic = None                                      # DRAM bytes per launch of the dominant kernel (ncu --set full capture)
    try:
        key = "fused" if (ctx.last_packed_path_was_fused and dominant == "contract") else dominant
        per_variant = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[key]["dram_bytes_per_variant"]
        traffic = per_variant * V
    except Exception:  # noqa: BLE001
        pass
    fused = ctx.last_packed_path_was_fused
    names = ({"scan": "scan_block_kernel (redo list)", "contract": "fused_tc_kernel", "finalize": "fused_finalize_kernel"} if fused
             else {"scan": "scan_block_kernel (+ expand_lists_kernel and the detection-path launch for truncated lists)",
                   "contract": "contract_tc_kernel", "finalize": "finalize_packed_kernel"})
    roofline = {"bound": "hbm", "kernel": names[dominant],
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                "stage_ms": stage_ms, "step_frac": stage_ms[dominant] / sum(stage_ms.values()),
                "whole_step": {"achieved": V * bytes_per_variant / (ms_total / args.steps / 1e3) / 1e9,
                               "frac": V * bytes_per_variant / (ms_total / args.steps / 1e3) / 1e9 / hbm_peak}}
    # secondary roof: the INT8 tensor pipe the contraction runs on (tcgen05.mma.kind::i8).  Denominator = the rate
    # measured on this pool's B200 for the contraction's instruction shape (M = 128, N = 112, K = 32, A in TMEM;
    # tools/pipe_peaks.cu -> profiles/r02_pipe_peaks.json); nominal 4.5 POP/s only if that file is missing.
    ops = 2.0 * V * n * 112                             # 7 digit slices x 16 columns per genotype
    int8_peak, int8_src = 4500.0, "nominal dense int8/fp8 (B200_PROFILING.md)"
    try:
        int8_peak = float(json.load(open(os.path.join(ROOT, "profiles", "r02_pipe_peaks.json")))["tcgen05_i8_ts_n112_tops"])
        int8_src = "measured: tcgen05.mma kind::i8, M=128 N=112 K=32, A in TMEM (profiles/r02_pipe_peaks.json)"
    except Exception:  # noqa: BLE001
        pass
    roofline["int8_tensor"] = {"kernel": "contract_tc_kernel", "achieved_tops": ops / (stage_ms["contract"] / 1e3) / 1e12,
                               "peak_tops": int8_peak, "peak_source": int8_src,
                               "frac": ops / (stage_ms["contract"] / 1e3) / 1e12 / int8_peak}

    # ---- end to end: host FP64 dosage chunks (what the Rcpp boundary delivers) through the C-ABI chunk driver
    e2e = None
    if not args.no_e2e:
        p = CHUNK
        G_pageable = np.asfortranarray(host_chunk_from_device(packed[:p], n, dev, d_order))   # what R hands over: pageable FP64
        host = torch.from_numpy(np.ascontiguousarray(G_pageable.T)).pin_memory()              # (p, n) C-order, pinned
        G = host.numpy().T                                                                    # n x p, F-order view
        out7 = [np.zeros(p) for _ in range(7)]
        nchunk = max(args.steps, 16)            # chunks per e2e leg: a 30 ms chunk needs more than K = 5 repetitions for a stable rate
        # headline leg: the block lies in PAGEABLE memory, as an R numeric matrix does (host threads pack everything)
        out7g = [np.zeros(p) for _ in range(7)]
        for _ in range(3):
            ctx.individual_analysis_chunk(G_pageable, api.IMPUTE_MEAN, 0, out7g)
        barrier()
        t0 = time.perf_counter()
        for _ in range(nchunk):
            ctx.individual_analysis_chunk(G_pageable, api.IMPUTE_MEAN, 0, out7g)
        barrier()
        dt_page = time.perf_counter() - t0
        for _ in range(max(3, args.warmup)):             # also lets the host/GPU packing split settle
            ctx.individual_analysis_chunk(G, api.IMPUTE_MEAN, 0, out7)
        barrier()
        t0 = time.perf_counter()
        for _ in range(nchunk):
            ctx.individual_analysis_chunk(G, api.IMPUTE_MEAN, 0, out7)
        barrier()
        dt_sync = time.perf_counter() - t0
        # the same chunks through the pipelined form of the call (submit / wait, two slots): the host packs chunk i+1
        # while the GPU copies and scores chunk i; every chunk still has its H2D and D2H inside the timed region
        out7p = [np.zeros(p) for _ in range(7)]
        for w_ in range(max(3, args.warmup)):
            ctx.chunk_submit(G, api.IMPUTE_MEAN, 0, w_ & 1); ctx.chunk_wait(w_ & 1, out7p)
        barrier()
        t0 = time.perf_counter()
        ctx.chunk_submit(G, api.IMPUTE_MEAN, 0, 0)
        for s_ in range(nchunk):
            if s_ + 1 < nchunk:
                ctx.chunk_submit(G, api.IMPUTE_MEAN, 0, (s_ + 1) & 1)
            ctx.chunk_wait(s_ & 1, out7p)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        p_gpu = min(p, int(ctx.gpu_pack_share() * p))       # variants the GPU packs itself over PCIe (share after re-balancing)
        dt_sync_t = torch.tensor([dt_sync], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt_sync_t, op=dist.ReduceOp.MAX)
        dt_sync = float(dt_sync_t.item())
        dt_page_t = torch.tensor([dt_page], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt_page_t, op=dist.ReduceOp.MAX)
        dt_page = float(dt_page_t.item())
        e2e = {"value": world * p * nchunk / dt_page, "unit": "variants/s", "chunks_timed": nchunk,
               "h2d_bytes_per_step": int(p * stride), "d2h_bytes_per_step": int(7 * p * 8),
               "call": "staar_individual_analysis_chunk (host FP64 n x 5000 dosage block in PAGEABLE memory, as R hands it "
                       "over; 2-bit packing by host threads, H2D of the packed columns, re-order, contraction, scan, "
                       "finaliser, D2H of 7 result arrays — all inside the timed region)",
               "ms_per_chunk": 1e3 * dt_page / nchunk, "host_bytes_read_per_step": int(p * n * 8),
               "host_read_gbs": world * p * n * 8 * nchunk / dt_page / 1e9,
               "pinned_block": {"value": world * p * nchunk / dt_sync, "unit": "variants/s", "ms_per_chunk": 1e3 * dt_sync / nchunk,
                                "h2d_bytes_per_step": int((p - p_gpu) * stride + p_gpu * n * 8),
                                "host_read_gbs": world * p * n * 8 * nchunk / dt_sync / 1e9,
                                "call": "same call, block in pinned memory: the GPU packs a share of the variants itself, "
                                        f"reading the host matrix over PCIe ({p_gpu} of {p} variants after re-balancing)"},
               "gpu_pack_share": ctx.gpu_pack_share(),
               "pipelined_calls": {"value": world * p * nchunk / float(dt.item()), "unit": "variants/s",
                                   "ms_per_chunk": 1e3 * float(dt.item()) / nchunk,
                                   "call": "staar_chunk_submit / staar_chunk_wait, two chunks in flight"}}
        # parity of the device-resident step against the host-buffer path on the same variants (bit-identical)
        same = all(np.array_equal(outs[i, :p].cpu().numpy(), out7[i], equal_nan=True) and
                   np.array_equal(out7[i], out7p[i], equal_nan=True) and
                   np.array_equal(out7[i], out7g[i], equal_nan=True) for i in range(7))
        e2e["matches_resident_path"] = bool(same)
        # the same chunk as SeqArray's raw dosage bytes (1 B/genotype on the host instead of the 8 B Rcpp widens it to)
        Gu8 = np.where(np.isnan(G), 255, G).astype(np.uint8, order="F")
        out7b = [np.zeros(p) for _ in range(7)]
        for w_ in range(3):
            ctx.chunk_submit(Gu8, api.IMPUTE_MEAN, 0, w_ & 1); ctx.chunk_wait(w_ & 1, out7b)
        barrier()
        t0 = time.perf_counter()
        ctx.chunk_submit(Gu8, api.IMPUTE_MEAN, 0, 0)
        for s_ in range(nchunk):
            if s_ + 1 < nchunk:
                ctx.chunk_submit(Gu8, api.IMPUTE_MEAN, 0, (s_ + 1) & 1)
            ctx.chunk_wait(s_ & 1, out7b)
        barrier()
        dtb_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dtb_t, op=dist.ReduceOp.MAX)
        dtb = float(dtb_t.item())
        e2e["raw_u8_input"] = {"value": world * p * nchunk / dtb, "unit": "variants/s", "ms_per_chunk": 1e3 * dtb / nchunk,
                               "host_bytes_read_per_step": int(p * n),
                               "matches_fp64_input": bool(all(np.array_equal(out7[i], out7b[i], equal_nan=True) for i in range(7)))}

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only): bounded sample of the same workload
    cpu = None
    pc = args.cpu_chunk
    nm_like = argparse.Namespace(Sigma_i=Sigma_i, Sigma_iX=Sigma_iX, cov=cov, residuals=residuals)
    Gc = np.asfortranarray(host_chunk_from_device(packed[:pc], n, dev, d_order)) if (e2e is not None or (rank == 0 and world == 1)) else None
    if e2e is not None:
        # the R driver's native call sequence on one chunk through this library's frozen-signature entry points (what the
        # Rcpp shim of INTEGRATION.md binds: FP64 G in, flipped FP64 Geno back out, dense + CSC marshalling, pageable
        # memory, host operands on every call); every rank runs it, time = max over ranks
        reference_chunk(ctx, Gc, nm_like)
        barrier()
        tf, tested_f = reference_chunk(ctx, Gc, nm_like)
        tf_t = torch.tensor([tf], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tf_t, op=dist.ReduceOp.MAX)
        tf = float(tf_t.item())
        e2e["frozen_signatures"] = {"value": world * pc / tf, "unit": "variants/s", "ms_per_chunk": 1e3 * tf, "variants": pc,
                                    "tested": tested_f,
                                    "call": "staar_matrix_flip_mean -> staar_individual_score_test (common, dense FP64 G) -> "
                                            "staar_individual_score_test_sp (rare, CSC G): the R driver's sequence "
                                            "through the 13 unchanged signatures, host operands on every call"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ref, kind, blas = load_reference()
        t, tested = reference_chunk(ref, Gc, nm_like)
        cpu = {"value": pc / t, "unit": "variants/s", "cores": (os.cpu_count() or 1) if blas else 1, "kind": kind,
               "sample": f"one chunk of {pc} variants x {n} samples ({tested} pass the MAC filter): native calls of the R "
                         f"driver (matrix_flip_mean, Individual_Score_Test, Individual_Score_Test_sp), {t:.1f} s; "
                         "reference sources with stand-in Armadillo/Rmath (oracle/ref_stub), dense GEMM on "
                         f"{'OpenBLAS all cores' if blas else 'in-order loops'}"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "variants/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "int8 x int8 -> int32 exact contraction of 54-bit sliced f64 operands; f64 statistics",
                "data": "synthetic", "config": workload_config(args, V), "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "outputs_sha256": outputs_sha256, "outputs_sha256_arrays": outputs_sha256_arrays,
                "packed_path": "fused" if ctx.last_packed_path_was_fused else "split"}
        emit(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ other configurations
def run_config(args):
    """--config 0 / 2 / 3 / 4: the other BASELINE.json configurations on resident 2-bit columns, one JSON line of the same
    shape (value = device-resident throughput of the configuration's flow; roofline against the roof that binds: HBM for
    the packed score paths, the measured FP64 pipe for the dense-GRM GEMM and the SPA passes).  Weak scaling: every rank
    runs the same shard size; time = max over ranks."""
    import torch
    import torch.distributed as dist
    from staarpipeline_b200 import api, synth
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    ctx = api.Context(local)
    peaks, pipes = {}, {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    try:
        pipes = json.load(open(os.path.join(ROOT, "profiles", "r02_pipe_peaks.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0)); fp64_peak = float(pipes.get("fp64_fma_tflops", 37.0))
    cfg = args.config
    seed = synth.BASE_SEED

    def resident(n, V, order=None):
        stride = api.packed_stride(n)
        packed = torch.empty((V, stride), dtype=torch.uint8, device=dev)
        first = rank * V
        for a in range(0, V, 125_000):
            b = min(V, a + 125_000)
            prm = synth.variant_params(n, first + a, b - a, seed)
            t = lambda x: torch.from_numpy(x.view(np.int64) if x.dtype == np.uint64 else x).to(dev)
            th, te, tm, sa = t(prm.thr_hom), t(prm.thr_het), t(prm.thr_miss), t(prm.store_alt)
            torch.cuda.synchronize()
            ctx.synth_packed_device(packed[a:b].data_ptr(), stride, n, first + a, b - a, seed, th.data_ptr(), te.data_ptr(),
                                    tm.data_ptr(), sa.data_ptr(), order.data_ptr() if order is not None else 0)
        return packed, stride

    def timed(step, sync):
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        for _ in range(args.warmup):
            step()
        sync()
        if world > 1:
            dist.barrier()
        ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_begin = time.perf_counter()
        l0 = ctx.launch_count
        e0.record(ext)
        for _ in range(args.steps):
            step()
        e1.record(ext)
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        clocks = sampler.stop(t_begin, time.perf_counter()) if rank == 0 else None
        return float(ms.item()) / args.steps, clocks, (ctx.launch_count - l0) // max(args.steps, 1)

    def sync_all():
        ctx.synchronize(); torch.cuda.synchronize()

    extra = {}
    if cfg == 0:
        n, V, q = 10_000, args.variants_per_gpu if args.variants_per_gpu != VARIANTS_PER_GPU else 2_000_000, Q_COVARIATES
        nm = synth.nullmodel_unrelated(n, seed=seed, q=q)
        ctx.set_nullmodel_sparse(nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
        packed, stride = resident(n, V)
        outs = torch.zeros((7, V), dtype=torch.float64, device=dev); optrs = [outs[i].data_ptr() for i in range(7)]
        ms, clocks, launches = timed(lambda: ctx.packed_score_device(packed.data_ptr(), stride, n, V, api.IMPUTE_MEAN, optrs), sync_all)
        bytes_v = n * 2 / 8 + 56
        roof = {"bound": "hbm", "achieved": V * bytes_v / (ms / 1e3) / 1e9, "peak": hbm_peak, "unit": "GB/s", "traffic": None,
                "kernel": "whole step (contract_tc_kernel + expand_lists_kernel + scan_block_kernel + finalize_packed_kernel)"}
        work = "BASELINE configs[0]: Individual_Analysis continuous trait, unrelated samples, n=10,000 (resident 2-bit columns)"
        dtype = "int8 x int8 -> int32 exact contraction of 54-bit sliced f64 operands; f64 statistics"
    elif cfg == 2:
        n, V = 20_000, args.variants_per_gpu if args.variants_per_gpu != VARIANTS_PER_GPU else 20_000
        rng = np.random.default_rng(seed)
        W = rng.normal(size=(n, 64))
        P = -(W @ W.T) / 64.0
        P[np.diag_indices(n)] += 3.0 + rng.random(n)
        P = np.asfortranarray(0.5 * (P + P.T)); res = rng.normal(size=n)
        ctx.set_nullmodel_dense(P, res)
        packed, stride = resident(n, V)
        outs = torch.zeros((7, V), dtype=torch.float64, device=dev); optrs = [outs[i].data_ptr() for i in range(7)]
        ms, clocks, launches = timed(lambda: ctx.packed_score_dense_device(packed.data_ptr(), stride, n, V, api.IMPUTE_MEAN, optrs), sync_all)
        af = outs[0].cpu().numpy(); maf = np.minimum(af, 1 - af)
        # variants with more than n/8 non-zero genotypes take P*G (2 n^2 flop each); the rest the carrier form
        dense_share = float((1 - (1 - maf) ** 2 > 1 / 8).mean())
        flops = V * dense_share * 2.0 * n * n
        roof = {"bound": "tensor", "achieved": flops / (ms / 1e3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s", "traffic": None,
                "kernel": "dgemm_PG_kernel (FP64 DMMA) for the dense share of the variants", "dense_share": dense_share,
                "peak_source": "measured FP64 FMA / DMMA rate (profiles/r02_pipe_peaks.json)"}
        work = "BASELINE configs[2]: Individual_Analysis with dense GRM (_denseGRM), n=20,000 (resident 2-bit columns; P 3.2 GB resident)"
        dtype = "f64"
    elif cfg == 3:
        n, k, V = 100_000, 3, args.variants_per_gpu if args.variants_per_gpu != VARIANTS_PER_GPU else 50_000
        nm = synth.nullmodel_multi(n, k, seed=seed, q=Q_COVARIATES)
        ctx.set_nullmodel_sparse(nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
        packed, stride = resident(n, V)
        af = torch.zeros(V, dtype=torch.float64, device=dev); maf = torch.zeros_like(af)
        score = torch.zeros(V * k, dtype=torch.float64, device=dev); pl = torch.zeros(V, dtype=torch.float64, device=dev)
        ms, clocks, launches = timed(lambda: ctx.packed_score_multi_device(packed.data_ptr(), stride, n, V, k, api.IMPUTE_MEAN, af.data_ptr(),
                                                                          maf.data_ptr(), score.data_ptr(), pl.data_ptr()), sync_all)
        bytes_v = n * 2 / 8 + 8 * (k + 3)
        roof = {"bound": "hbm", "achieved": V * bytes_v / (ms / 1e3) / 1e9, "peak": hbm_peak, "unit": "GB/s", "traffic": None,
                "kernel": "contract_multi_packed_kernel + quad_multi_packed_kernel (one CTA per variant, carrier lists)"}
        work = "BASELINE configs[3]: 3 correlated traits jointly (_multi), sparse GRM, n=100,000 (resident 2-bit columns)"
        dtype = "f64"
    else:
        n, V = 400_000, args.variants_per_gpu if args.variants_per_gpu != VARIANTS_PER_GPU else 20_000
        nb = synth.nullmodel_binary(n, seed=seed, q=Q_COVARIATES)
        ctx.set_nullmodel_sparse(nb.Sigma_i, nb.Sigma_iX, nb.cov, nb.residuals)
        ctx.set_nullmodel_spa(nb.XW, nb.XXWX_inv, nb.residuals, nb.muhat)
        order = torch.from_numpy(ctx.nullmodel_sample_order()).to(dev)
        packed, stride = resident(n, V, order)
        idx = torch.zeros(V, dtype=torch.int64, device=dev); sel = torch.zeros(V, dtype=torch.int64, device=dev)
        outs = torch.zeros((7, V), dtype=torch.float64, device=dev); optrs = [outs[i].data_ptr() for i in range(7)]
        pv = torch.zeros(V, dtype=torch.float64, device=dev); tr = torch.zeros((V, 4), dtype=torch.int64, device=dev)
        tol = float(np.finfo(np.float64).eps ** 0.25)
        stat = {}

        def flow():
            nt, nsel = ctx.packed_analysis_device(packed.data_ptr(), stride, n, V, api.IMPUTE_MEAN, 20.0, 0.05, idx.data_ptr(), optrs, sel.data_ptr())
            cols = idx[sel[:nsel]].contiguous()
            torch.cuda.synchronize()
            ctx.packed_spa_device(packed.data_ptr(), stride, n, cols.data_ptr(), nsel, api.IMPUTE_MEAN, tol, 1000, pv.data_ptr(), tr.data_ptr())
            stat.update(tested=int(nt), spa_variants=int(nsel))
        ms, clocks, launches = timed(flow, sync_all)
        passes = int(tr[:stat["spa_variants"], 0].sum().item())
        flops = passes * float(n) * 66.0                       # c_K ~ 2q (recompute) + ~40 per sample and pass
        roof = {"bound": "tensor", "achieved": flops / (ms / 1e3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s", "traffic": None,
                "kernel": "spa_kernel (FP64 ALU; one cluster per variant)", "cgf_passes": passes,
                "peak_source": "measured FP64 FMA rate (profiles/r02_pipe_peaks.json); flop count = passes x n x 66 (SURVEY 8d)"}
        extra = stat
        work = ("BASELINE configs[4]: dichotomous trait 1:100 with SPA, n=400,000: score test + MAC filter + SPA pre-filter, "
                "then SPA on the selected resident columns (resident 2-bit columns)")
        dtype = "f64 (SPA), exact int8 contraction (stage 1)"
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof.setdefault("peak_source", "measured (MEASURED_PEAKS.json)" if peaks else "fallback")
    if rank == 0:
        line = {"metric": METRIC.replace("n=100k", f"n={n}"), "value": world * V / (ms / 1e3), "unit": "variants/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": dtype, "data": "synthetic",
                "config": {"workload": work, "n_samples": n, "variants_per_gpu_per_step": V, "genotype_bits": 2, "impute": "mean",
                           "parallelism": f"variant-sharded x{world}", "l2": "inputs larger than L2" if V * stride > 2e8 else "inputs fit L2 partly", **extra},
                "e2e": None, "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": None}
        emit(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def host_chunk_from_device(packed_rows, n, dev, d_order):
    """unpack device-resident 2-bit columns (null-model sample order) into the FP64 `$dosage` matrix in the caller's
    sample order (n x p, F-order, NaN = missing)"""
    import torch
    p = packed_rows.shape[0]
    sh = torch.tensor([0, 2, 4, 6], dtype=torch.uint8, device=dev)
    by_pos = ((packed_rows.unsqueeze(-1) >> sh) & 3).reshape(p, -1)[:, :n]
    codes = torch.empty_like(by_pos)
    codes[:, d_order.long()] = by_pos              # position j holds sample order[j]
    G = codes.to(torch.float64)
    G[codes == 3] = float("nan")
    return G.cpu().numpy().T          # (n, p) view of a C-order (p, n) array == F-order n x p


_json_out = None


def emit(line: str):
    """the bench line, on the process's original stdout"""
    out = _json_out or sys.stdout
    out.write(line + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=N_SAMPLES)
    ap.add_argument("--variants-per-gpu", type=int, default=VARIANTS_PER_GPU)
    ap.add_argument("--cpu-chunk", type=int, default=2000, help="variants in the bounded cpu_baseline sample")
    ap.add_argument("--ref-chunk", type=int, default=2000, help="variants per step of --impl reference")
    ap.add_argument("--config", type=int, default=1, choices=[0, 1, 2, 3, 4],
                    help="BASELINE.json configs[k]; 1 (default) is the headline workload, the others print the same line shape")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    # stdout carries the ONE JSON line and nothing else: anything a library writes to fd 1 (NCCL prints its version
    # banner there when NCCL_DEBUG is set on the box) goes to stderr for the rest of the run
    global _json_out
    sys.stdout.flush()
    _json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    elif args.config != 1:
        run_config(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
