"""Dense-GRM entry points once each (tools; for ncu captures of dgemm_PG_kernel / quad_pairs_carriers_P_kernel).
usage: python tools/dense_grm_time.py n p"""
import sys, time, os
import numpy as np, scipy.sparse as sp
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
from staarpipeline_b200 import api, synth
n, p = int(sys.argv[1]), int(sys.argv[2])
ctx = api.Context(0)
nd = synth.nullmodel_dense_grm(n, markers=500)
G = synth.codes_to_dosage(synth.dosage_codes(n, 0, p)); G = np.nan_to_num(G, nan=0.0)
flip = G.mean(axis=0) / 2 > 0.5; G[:, flip] = 2.0 - G[:, flip]; G = np.asfortranarray(G)
rare = G.mean(axis=0) / 2 < 0.05
for _ in range(2):
    t0 = time.perf_counter(); ctx.Individual_Score_Test_denseGRM(np.asfortranarray(G[:, ~rare]), nd.P, nd.residuals); t1 = time.perf_counter()
    ctx.Individual_Score_Test_sp_denseGRM(sp.csc_matrix(G[:, rare]), nd.P, nd.residuals); t2 = time.perf_counter()
    print(n, "common", int((~rare).sum()), f"{(t1-t0)*1e3:.1f} ms", "rare", int(rare.sum()), f"{(t2-t1)*1e3:.1f} ms", flush=True)
