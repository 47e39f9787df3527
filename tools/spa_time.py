"""SPA entry point timing (tools; not part of the product).  usage: python tools/spa_time.py n p"""
import sys, time, os
import numpy as np
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
from staarpipeline_b200 import api, synth
n, p = int(sys.argv[1]), int(sys.argv[2])
ctx = api.Context(0)
nb = synth.nullmodel_binary(n)
tol = float(np.finfo(np.float64).eps ** 0.25)
G = synth.codes_to_dosage(synth.dosage_codes(n, 0, p)); G = np.nan_to_num(G, nan=0.0); G = np.asfortranarray(2.0 - G)
for _ in range(2):
    t0 = time.perf_counter(); pv, tr = ctx.Individual_Score_Test_SPA(G, nb.XW, nb.XXWX_inv, nb.residuals, nb.muhat, tol, 1000, return_trace=True); dt = time.perf_counter() - t0
    print(n, p, f"{dt*1e3:.1f} ms", "passes", tr[:, 0].mean(), "max", tr[:, 0].max(), flush=True)
