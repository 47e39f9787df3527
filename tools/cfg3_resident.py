#!/usr/bin/env python
"""BASELINE configs[3] (dense GRM) on resident packed genotypes (tools; not the headline bench).
usage: python tools/cfg3_resident.py [n] [V]"""
import json, sys, time, os
import numpy as np
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
import torch
from staarpipeline_b200 import api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000
V = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000
dev = torch.device("cuda:0")
ctx = api.Context(0)
nd = synth.nullmodel_dense_grm(n)
ctx.set_nullmodel_dense(nd.P, nd.residuals)
stride = api.packed_stride(n)
packed = torch.empty((V, stride), dtype=torch.uint8, device=dev)
prm = synth.variant_params(n, 0, V)
t = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).to(dev)
th, te, tm, sa = t(prm.thr_hom), t(prm.thr_het), t(prm.thr_miss), t(prm.store_alt)
torch.cuda.synchronize()
ctx.synth_packed_device(packed.data_ptr(), stride, n, 0, V, synth.BASE_SEED, th.data_ptr(), te.data_ptr(), tm.data_ptr(), sa.data_ptr(), 0)
outs = torch.zeros((7, V), dtype=torch.float64, device=dev)
ptrs = [outs[i].data_ptr() for i in range(7)]
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.packed_score_dense_device(packed.data_ptr(), stride, n, V, api.IMPUTE_MEAN, ptrs)
    dt = time.perf_counter() - t0
maf = outs[1].cpu().numpy()
print(json.dumps({"what": "cfg3 dense GRM on resident packed columns (carrier form for <= n/8 non-zero genotypes, P*G for the rest)",
                  "n": n, "variants": V, "s": dt, "variants_per_s": V / dt, "common_frac_maf_ge_0.05": float((maf >= 0.05).mean())}), flush=True)
ctx.close()
