import json, sys, time, os
import numpy as np
sys.path[:0] = ["/root/repo"]
import torch
from staarpipeline_b200 import api, synth
ctx = api.Context(0)
for n in (100_000, 400_000):
    nb = synth.nullmodel_binary(n)
    tol = float(np.finfo(np.float64).eps ** 0.25)
    for p in (148, 296, 592):
        G = synth.codes_to_dosage(synth.dosage_codes(n, 0, p)); G = np.nan_to_num(G, nan=0.0); G = np.asfortranarray(2.0 - G)
        ctx.Individual_Score_Test_SPA(G, nb.XW, nb.XXWX_inv, nb.residuals, nb.muhat, tol, 1000)
        t0 = time.perf_counter(); pv, tr = ctx.Individual_Score_Test_SPA(G, nb.XW, nb.XXWX_inv, nb.residuals, nb.muhat, tol, 1000, return_trace=True); dt = time.perf_counter() - t0
        print(n, p, f"{dt*1e3:.1f} ms", "passes", tr[:,0].mean(), "per-variant ms", dt*1e3/p, flush=True)
