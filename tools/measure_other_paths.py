#!/usr/bin/env python
"""Throughput of the FP64-signature entry points of BASELINE configs 3 (dense GRM), 4 (three traits) and the
conditional test of config 5, through the frozen-signature ABI with host inputs (H2D inside).  Not the headline
bench; one JSON line per measurement.  Run on a B200: python tools/measure_other_paths.py [n_dense]"""
import json, sys, time, os
import numpy as np, scipy.sparse as sp
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
from staarpipeline_b200 import api, synth

ctx = api.Context(0)

def flipped_block(n, first, p):
    G = synth.codes_to_dosage(synth.dosage_codes(n, first, p))
    mu = np.nanmean(G, axis=0); mu = np.where(np.isnan(mu), 0.0, mu)
    G = np.where(np.isnan(G), mu[None, :], G)
    flip = G.mean(axis=0) / 2 > 0.5
    G[:, flip] = 2.0 - G[:, flip]
    return np.asfortranarray(G)

def timed(fn, reps=2):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps): out = fn()
    return (time.perf_counter() - t0) / reps, out

# ---- config 3: dense GRM
n3 = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
p3 = 2000
t0 = time.perf_counter(); nd = synth.nullmodel_dense_grm(n3); tgen = time.perf_counter() - t0
G = flipped_block(n3, 0, p3)
dt, out = timed(lambda: ctx.Individual_Score_Test_denseGRM(G, nd.P, nd.residuals))
print(json.dumps({"what": "cfg3 Individual_Score_Test_denseGRM (host FP64 G, P cached on device)", "n": n3, "variants": p3, "s": dt,
                  "variants_per_s": p3 / dt, "fp64_tflops": 2.0 * n3 * n3 * p3 / dt / 1e12, "gen_s": tgen}), flush=True)
# the R driver's split (R/Individual_Analysis.R:242,324,372): common variants -> dense entry, rare (MAF < 0.05) -> sparse entry
maf = G.mean(axis=0) / 2
rare = maf < 0.05
Gs = sp.csc_matrix(G[:, rare]); Gc = np.asfortranarray(G[:, ~rare])
dt_r, out_r = timed(lambda: ctx.Individual_Score_Test_sp_denseGRM(Gs, nd.P, nd.residuals))
dt_c, out_c = timed(lambda: ctx.Individual_Score_Test_denseGRM(Gc, nd.P, nd.residuals))
print(json.dumps({"what": "cfg3 split as the R driver does: rare -> Individual_Score_Test_sp_denseGRM (carrier form), common -> _denseGRM (P*G)",
                  "n": n3, "rare": int(rare.sum()), "common": int((~rare).sum()), "rare_s": dt_r, "common_s": dt_c,
                  "rare_variants_per_s": rare.sum() / dt_r, "common_variants_per_s": (~rare).sum() / max(dt_c, 1e-9),
                  "variants_per_s": p3 / (dt_r + dt_c), "mean_carriers_rare": Gs.nnz / max(int(rare.sum()), 1),
                  "max_rel_diff_vs_dense_entry": float(np.nanmax(np.abs(out_r["Score_se"] - out["Score_se"][rare]) / np.maximum(out["Score_se"][rare], 1e-300)))}), flush=True)
Gs8 = sp.hstack([Gs] * 8, format="csc")
dt8, _ = timed(lambda: ctx.Individual_Score_Test_sp_denseGRM(Gs8, nd.P, nd.residuals))
print(json.dumps({"what": "cfg3 rare route, 8x the columns in one call (fixed per-call cost vs per-variant cost)", "variants": Gs8.shape[1],
                  "s": dt8, "variants_per_s": Gs8.shape[1] / dt8}), flush=True)
del nd, G, Gs, Gs8

# ---- config 4: three traits, sparse GRM, Kronecker-expanded sparse G as R/Individual_Analysis.R:267 builds it
n4, k, pp = 100_000, 3, 2000
nm = synth.nullmodel_multi(n4, k)
G0 = sp.csc_matrix(flipped_block(n4, 0, pp))
GK = sp.kron(sp.identity(k, format="csc"), G0, format="csc")
dt, out = timed(lambda: ctx.Individual_Score_Test_sp_multi(GK, nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals, k))
print(json.dumps({"what": "cfg4 Individual_Score_Test_sp_multi (k=3, host CSC Kronecker G), all variants in one call (dense route if any column is common)",
                  "n": n4, "variants": pp, "s": dt, "variants_per_s": pp / dt, "nnz_frac": G0.nnz / (n4 * pp)}), flush=True)
maf4 = np.asarray(G0.mean(axis=0)).ravel() / 2
rare4 = maf4 < 0.05
GKr = sp.kron(sp.identity(k, format="csc"), G0[:, rare4], format="csc")
dt, out_r = timed(lambda: ctx.Individual_Score_Test_sp_multi(GKr, nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals, k))
print(json.dumps({"what": "cfg4 rare variants (MAF < 0.05, as the R driver routes them) -> Individual_Score_Test_sp_multi, carrier form",
                  "n": n4, "variants": int(rare4.sum()), "s": dt, "variants_per_s": rare4.sum() / dt,
                  "max_abs_diff_pvalue_log_vs_dense_route": float(np.nanmax(np.abs(out_r["pvalue_log"] - out["pvalue_log"][rare4])))}), flush=True)
del nm, G0, GK

# ---- config 5: conditional test on 10 known variants (optimal adjustment: X_adj = [known | X])
n5 = 400_000
nb = synth.nullmodel_binary(n5)
known = flipped_block(n5, 5000, 10)
X_adj, res_c = synth.conditional_design(nb, known, "optimal")
for p5 in (1, 256):
    G = flipped_block(n5, 0, p5)
    dt, out = timed(lambda: ctx.Individual_Score_Test_cond(G, nb.Sigma_i, nb.Sigma_iX, nb.cov, X_adj, res_c))
    print(json.dumps({"what": "cfg5 Individual_Score_Test_cond (m=%d adjustment columns)" % X_adj.shape[1], "n": n5, "variants": p5, "s": dt,
                      "variants_per_s": p5 / dt}), flush=True)
ctx.close()
