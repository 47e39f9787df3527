#!/usr/bin/env python
"""BASELINE configs[5] flow on resident packed genotypes (tools; not the headline bench): n = 400,000, binary trait at
~1 % cases, stage 1 = score test + MAF/MAC filter + SPA pre-filter (p < 0.05) on the device, stage 2 = SPA on the
selected columns decoded on the device.  usage: python tools/cfg5_resident.py [V]"""
import json, sys, time, os
import numpy as np
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
import torch
from staarpipeline_b200 import api, synth

V = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000
n = 400_000
dev = torch.device("cuda:0")
ctx = api.Context(0)
nb = synth.nullmodel_binary(n)
ctx.set_nullmodel_sparse(nb.Sigma_i, nb.Sigma_iX, nb.cov, nb.residuals)
ctx.set_nullmodel_spa(nb.XW, nb.XXWX_inv, nb.residuals, nb.muhat)
stride = api.packed_stride(n)
order = torch.from_numpy(ctx.nullmodel_sample_order()).to(dev)
packed = torch.empty((V, stride), dtype=torch.uint8, device=dev)
prm = synth.variant_params(n, 0, V)
t = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).to(dev)
th, te, tm, sa = t(prm.thr_hom), t(prm.thr_het), t(prm.thr_miss), t(prm.store_alt)
torch.cuda.synchronize()
ctx.synth_packed_device(packed.data_ptr(), stride, n, 0, V, synth.BASE_SEED, th.data_ptr(), te.data_ptr(), tm.data_ptr(), sa.data_ptr(), order.data_ptr())
idx = torch.zeros(V, dtype=torch.int64, device=dev); sel = torch.zeros(V, dtype=torch.int64, device=dev)
outs = torch.zeros((7, V), dtype=torch.float64, device=dev)
ptrs = [outs[i].data_ptr() for i in range(7)]
tol = float(np.finfo(np.float64).eps ** 0.25)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    nt, nsel = ctx.packed_analysis_device(packed.data_ptr(), stride, n, V, api.IMPUTE_MEAN, 20.0, 0.05, idx.data_ptr(), ptrs, sel.data_ptr())
    t1 = time.perf_counter()
    cols = idx[sel[:nsel]].contiguous()
    pv = torch.zeros(nsel, dtype=torch.float64, device=dev); tr = torch.zeros((nsel, 4), dtype=torch.int64, device=dev)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    ctx.packed_spa_device(packed.data_ptr(), stride, n, cols.data_ptr(), nsel, api.IMPUTE_MEAN, tol, 1000, pv.data_ptr(), tr.data_ptr())
    t3 = time.perf_counter()
print(json.dumps({"what": "cfg5 resident flow (n=4e5): stage 1 + SPA stage on the device", "n": n, "variants": V, "tested": nt,
                  "spa_variants": nsel, "stage1_s": t1 - t0, "spa_s": t3 - t2, "spa_variants_per_s": nsel / (t3 - t2),
                  "variants_per_s_overall": V / ((t1 - t0) + (t3 - t2)),
                  "p_min": float(pv.min().item()) if nsel else None,
                  "passes": {k: float(np.percentile(tr[:, 0].cpu().numpy(), k)) for k in (50, 90, 99, 100)},
                  "passes_total": int(tr[:, 0].sum().item()), "fallback_variants": int((tr[:, 3] & 5 != 0).sum().item())}), flush=True)
trh = tr.cpu().numpy(); o = outs.cpu().numpy(); s_ = sel[:nsel].cpu().numpy(); ix = idx.cpu().numpy()
slow = np.flatnonzero(trh[:, 0] > 1000)[:12]
for k in slow:
    r = s_[k]
    print(json.dumps({"variant": int(ix[r]), "AF": o[0, r], "MAF": o[1, r], "mac": o[1, r] * 2 * n, "score": o[2, r], "se": o[3, r],
                      "p_score": float(np.exp(-o[4, r])), "p_spa": float(pv[k].item()), "passes": int(trh[k, 0]), "nr_iter": int(trh[k, 1]), "bits": int(trh[k, 3] & 255)}))
ctx.close()
