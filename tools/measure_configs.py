#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations' kernels (not the headline bench).
Prints one JSON line per measurement.  Run on a B200: python tools/measure_configs.py"""
import json, sys, time, os
import numpy as np
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
import torch
from staarpipeline_b200 import api, synth

dev = torch.device("cuda:0")
ctx = api.Context(0)

def packed_resident(n, V, nm, label):
    ctx.set_nullmodel_sparse(nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
    stride = api.packed_stride(n)
    order = torch.from_numpy(ctx.nullmodel_sample_order()).to(dev)
    packed = torch.empty((V, stride), dtype=torch.uint8, device=dev)
    prm = synth.variant_params(n, 0, V)
    t = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).to(dev)
    th, te, tm, sa = t(prm.thr_hom), t(prm.thr_het), t(prm.thr_miss), t(prm.store_alt)
    torch.cuda.synchronize()
    ctx.synth_packed_device(packed.data_ptr(), stride, n, 0, V, synth.BASE_SEED, th.data_ptr(), te.data_ptr(), tm.data_ptr(), sa.data_ptr(), order.data_ptr())
    outs = torch.zeros((7, V), dtype=torch.float64, device=dev)
    ptrs = [outs[i].data_ptr() for i in range(7)]
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    for _ in range(2):
        ctx.packed_score_device(packed.data_ptr(), stride, n, V, api.IMPUTE_MEAN, ptrs)
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(3):
        ctx.packed_score_device(packed.data_ptr(), stride, n, V, api.IMPUTE_MEAN, ptrs)
    e1.record(ext); ctx.synchronize(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    ctx.set_profiling(True); ctx.packed_score_device(packed.data_ptr(), stride, n, V, api.IMPUTE_MEAN, ptrs); st = ctx.last_stage_ms(); ctx.set_profiling(False)
    print(json.dumps({"what": label, "n": n, "variants": V, "ms": ms, "variants_per_s": V / ms * 1e3,
                      "hbm_frac": V * (stride + 56) / (ms / 1e3) / 1e9 / 6542.1, "stage_ms": st}), flush=True)
    del packed, outs

if len(sys.argv) > 2 and sys.argv[1] == "n":          # python tools/measure_configs.py n <samples> [variants]: one size, sparse GRM
    nn = int(sys.argv[2]); vv = int(sys.argv[3]) if len(sys.argv) > 3 else 400_000
    packed_resident(nn, vv, synth.nullmodel_sparse_grm(nn), "packed resident, sparse GRM, n=%d, STAAR_B200_SCAN_WARPS=%s" % (nn, os.environ.get("STAAR_B200_SCAN_WARPS", "auto")))
    ctx.close(); sys.exit(0)
# cfg1: unrelated, n = 10,000
packed_resident(10_000, 2_000_000, synth.nullmodel_unrelated(10_000), "cfg1 packed resident (unrelated, n=1e4)")
if len(sys.argv) > 1 and sys.argv[1] == "cfg1":
    ctx.close(); sys.exit(0)
# cfg5 stage 1 sample size: n = 400,000 (unrelated: Sigma_i diagonal, as fit_nullmodel gives for a binary trait without GRM)
nb = synth.nullmodel_binary(400_000)
packed_resident(400_000, 100_000, nb, "cfg5 stage-1 packed resident (binary trait weights, n=4e5)")
# cfg5 SPA kernel at n = 400,000 on 296 variants (host FP64 G through the frozen-signature entry point: includes H2D)
n, p = 400_000, 296
G = synth.codes_to_dosage(synth.dosage_codes(n, 0, p)); G = np.nan_to_num(G, nan=0.0); G = np.asfortranarray(2.0 - G)
tol = float(np.finfo(np.float64).eps ** 0.25)
pv, tr = ctx.Individual_Score_Test_SPA(G, nb.XW, nb.XXWX_inv, nb.residuals, nb.muhat, tol, 1000, return_trace=True)
t0 = time.perf_counter(); pv, tr = ctx.Individual_Score_Test_SPA(G, nb.XW, nb.XXWX_inv, nb.residuals, nb.muhat, tol, 1000, return_trace=True); dt = time.perf_counter() - t0
passes = float(tr[:, 0].mean())
print(json.dumps({"what": "cfg5 SPA entry point (host FP64 G, H2D inside)", "n": n, "variants": p, "s": dt, "variants_per_s": p / dt,
                  "mean_passes_over_n": passes, "approx_fp64_gflops": p * passes * n * 60 / dt / 1e9}), flush=True)
ctx.close()
