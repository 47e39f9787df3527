"""GPU diagnostic for the fused single-pass kernel (csrc/fused_tc.cu): compares it, stage by stage, with the split
route (contraction + block-per-variant scan) and with numpy expectations computed from Sigma_i, and prints where the
first differences are.  Usage: python tools/fused_check.py [--time]"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from staarpipeline_b200 import api, synth  # noqa: E402
import cases  # noqa: E402

KEYS = ("AF", "MAF", "Score", "Score_se", "pvalue_log", "Est", "Est_se")


def expected_scan(codes, Sigma_i, flip):
    """2 * sum d_j [g_j == 2] + sum_{j != k} s_jk g_j g_k over post-flip codes, missing -> 0 (what acc + 2 dval cnt is)"""
    g = codes.astype(np.float64)
    mis = codes == 3
    g = np.where(flip[None, :], 2.0 - g, g)
    g[mis] = 0.0
    d = Sigma_i.diagonal()
    off = Sigma_i - sp.diags(d)
    hom = 2.0 * (d[:, None] * (g == 2.0)).sum(axis=0)
    kin = np.einsum("ij,ij->j", g, off @ g)
    return hom + kin


def check(ctx, n, p, q, seed, related=True, impute="mean", miss=None, tiles=2):
    nm = synth.nullmodel_sparse_grm(n, seed=seed, q=q, shuffle=True) if related else synth.nullmodel_unrelated(n, seed=seed, q=q)
    G0 = cases.raw_dosage(n, p, seed)
    if miss:
        G0[np.random.default_rng(seed).random(G0.shape) < miss] = np.nan
    codes = np.where(np.isnan(G0) | (G0 <= -1), 3, G0).astype(np.uint8)
    packed = synth.pack_codes(np.asfortranarray(codes))
    P = packed.shape[0]
    flag = api.IMPUTE_MEAN if impute == "mean" else api.IMPUTE_MINOR
    ctx.set_nullmodel_sparse(nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
    ctx.set_packed_path("split")
    a = ctx.packed_score_host(packed, n, flag)
    da = ctx.packed_debug_dump(P)
    ctx.set_packed_path("fused")
    b = ctx.packed_score_host(packed, n, flag)
    fused = ctx.last_packed_path_was_fused
    tag = f"n={n} p={P} q={q} related={related} impute={impute} miss={miss}"
    if not fused:
        print(f"[{tag}] fused route not taken"); return False
    db = ctx.packed_debug_dump(P)
    df = ctx.packed_debug_fused(P)
    ok = True
    for k in ("hi", "lo"):
        bad = np.flatnonzero((da[k] != db[k]).any(axis=1))
        if bad.size:
            ok = False
            i = bad[0]
            print(f"[{tag}] {k}: {bad.size} variants differ, first {i}: split {da[k][i][:4]} fused {db[k][i][:4]}; rows {bad[:10]}")
    cm = (codes == 3).sum(axis=0)
    if not np.array_equal(df["cm"], cm):
        ok = False
        bad = np.flatnonzero(df["cm"] != cm)
        print(f"[{tag}] cm: {bad.size} differ, first {bad[0]}: got {df['cm'][bad[0]]} want {cm[bad[0]]}")
    flip = (da["flip"] & 1) != 0
    mism = np.flatnonzero(df["guess"].astype(bool) != flip)
    # expected scanner sums where the guess was right
    dval = nm.Sigma_i.diagonal()
    want = expected_scan(codes, sp.csc_matrix(nm.Sigma_i), df["guess"].astype(bool))
    got = df["acc"]
    # acc excludes the uniform-diagonal count part: add it back with the last diagonal value in null-model order
    order = ctx.nullmodel_sample_order()
    dlast = dval[order[-1]]
    got_total = got + 2.0 * dlast * df["cnt"]
    err = np.abs(got_total - want) / np.maximum(1e-300, np.abs(want).max())
    if err.max() > 1e-12:
        ok = False
        i = int(np.argmax(err))
        print(f"[{tag}] scan sums: max rel err {err.max():.3e} at {i}: got {got_total[i]!r} want {want[i]!r} (acc {got[i]!r}, cnt {df['cnt'][i]}, guess {df['guess'][i]})")
        print("   bad variants:", np.flatnonzero(err > 1e-12)[:20])
    for k in KEYS:
        x, y = a[k], b[k]
        if k in ("AF", "MAF"):
            good = np.array_equal(x, y, equal_nan=True)
            e = 0.0 if good else 1.0
        else:
            scale = np.maximum(np.abs(x), 1e-3 * np.abs(x).max() if k in ("Score", "Est") else 1e-300)
            with np.errstate(invalid="ignore"):
                ee = np.abs(x - y) / scale
            ee[np.isnan(ee)] = 0.0
            e = float(ee.max()); good = e < 1e-11 and np.array_equal(np.isnan(x), np.isnan(y))
        if not good:
            ok = False
            i = int(np.argmax(np.abs(np.nan_to_num(x) - np.nan_to_num(y))))
            print(f"[{tag}] {k}: max err {e:.3e}, worst variant {i}: split {x[i]!r} fused {y[i]!r} (cm {cm[i]}, flip {flip[i]}, guess {df['guess'][i]})")
    print(f"[{tag}] {'OK' if ok else 'FAIL'}  redo {df['redo_count']} (guess mismatches {mism.size}, cm>cap or mismatch)")
    return ok


def timing(ctx, n=100_000, p=250_000):
    import torch
    dev = torch.device("cuda:0")
    nm = synth.nullmodel_sparse_grm(n, seed=synth.BASE_SEED, q=13, shuffle=True)
    ctx.set_nullmodel_sparse(nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
    stride = api.packed_stride(n)
    prm = synth.variant_params(n, 0, p)
    t = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).to(dev)
    th, te, tm, sa = t(prm.thr_hom), t(prm.thr_het), t(prm.thr_miss), t(prm.store_alt)
    order = torch.from_numpy(ctx.nullmodel_sample_order().astype(np.int32)).to(dev)
    cols = torch.zeros((p, stride), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    ctx.synth_packed_device(cols.data_ptr(), stride, n, 0, p, synth.BASE_SEED, th.data_ptr(), te.data_ptr(), tm.data_ptr(),
                            sa.data_ptr(), order.data_ptr())
    outs = torch.zeros((7, p), dtype=torch.float64, device=dev)
    ctx.set_profiling(True)
    res = {}
    runs = [("split", 2, 0), ("fused", 2, 0)]
    if "--dbg" in sys.argv:
        extra_dbg = sys.argv[sys.argv.index("--dbg") + 1:]
        dbgs = [int(x) for x in extra_dbg if x.isdigit()] or [2 + 32, 2 + 32 + 256, 2 + 32 + 512, 6]
        runs += [("fused", 2, d) for d in dbgs]
    for path, tiles, dbg in runs:
        ctx.set_packed_path(path)
        os.environ["STAAR_B200_FUSED_TILES"] = str(tiles)
        os.environ["STAAR_B200_FUSED_DBG"] = str(dbg)
        for it in range(3):
            t0 = time.perf_counter()
            ctx.packed_score_device(cols.data_ptr(), stride, n, p, api.IMPUTE_MEAN, [outs[i].data_ptr() for i in range(7)])
            ctx.synchronize()
            dt = time.perf_counter() - t0
        ms = ctx.last_stage_ms()
        extra = ""
        if path == "fused":
            extra = f" redo_count {ctx.packed_debug_fused(p)['redo_count']}"
        print(f"time {path} tiles={tiles} dbg={dbg}: {dt * 1e3:.2f} ms per {p} variants = {dt / p * 1e9:.2f} ns/variant; stages {ms}{extra}", flush=True)
        if dbg == 0 and tiles == 2:
            res[path] = outs.cpu().numpy().copy()
    for i, k in enumerate(KEYS):
        x, y = res["split"][i], res["fused"][i]
        with np.errstate(invalid="ignore", divide="ignore"):
            e = np.abs(x - y) / np.maximum(np.abs(x), 1e-3 * np.nanmax(np.abs(x)))
        print(f"   full-size {k}: max rel diff split vs fused {np.nanmax(e):.3e}")


if __name__ == "__main__":
    ctx = api.Context(0)
    allok = True
    for args in ((255, 9, 1, 1, True, "mean", None), (1000, 70, 13, 2, True, "mean", None), (777, 40, 5, 3, False, "mean", None),
                 (2049, 300, 13, 4, True, "minor", None), (5000, 200, 13, 5, True, "mean", 0.05),
                 (20000, 300, 13, 6, True, "mean", 0.002)):
        try:
            allok &= check(ctx, *args)
        except Exception as e:  # noqa: BLE001
            allok = False
            print("EXCEPTION", args, repr(e))
            break
    print("ALL OK" if allok else "SOME FAILED", flush=True)
    if "--time" in sys.argv and allok:
        timing(ctx)
    ctx.close()
