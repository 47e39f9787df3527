#!/usr/bin/env python
"""BASELINE configs[4] (three correlated traits, sparse GRM, n = 100,000) on resident packed genotypes (tools).
usage: python tools/cfg4_resident.py [V]"""
import json, sys, time, os
import numpy as np
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
import torch
from staarpipeline_b200 import api, synth

V = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000
n, k = 100_000, 3
dev = torch.device("cuda:0")
ctx = api.Context(0)
nm = synth.nullmodel_multi(n, k)
ctx.set_nullmodel_sparse(nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
stride = api.packed_stride(n)
packed = torch.empty((V, stride), dtype=torch.uint8, device=dev)
prm = synth.variant_params(n, 0, V)
t = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).to(dev)
th, te, tm, sa = t(prm.thr_hom), t(prm.thr_het), t(prm.thr_miss), t(prm.store_alt)
torch.cuda.synchronize()
ctx.synth_packed_device(packed.data_ptr(), stride, n, 0, V, synth.BASE_SEED, th.data_ptr(), te.data_ptr(), tm.data_ptr(), sa.data_ptr(), 0)
af = torch.zeros(V, dtype=torch.float64, device=dev); maf = torch.zeros_like(af)
score = torch.zeros(V * k, dtype=torch.float64, device=dev); pl = torch.zeros(V, dtype=torch.float64, device=dev)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.packed_score_multi_device(packed.data_ptr(), stride, n, V, k, api.IMPUTE_MEAN, af.data_ptr(), maf.data_ptr(), score.data_ptr(), pl.data_ptr())
    dt = time.perf_counter() - t0
print(json.dumps({"what": "cfg4 three traits on resident packed columns (Kronecker carrier lists on the device)", "n": n, "k": k,
                  "variants": V, "s": dt, "variants_per_s": V / dt, "common_frac_maf_ge_0.05": float((maf.cpu().numpy() >= 0.05).mean())}), flush=True)
ctx.close()
