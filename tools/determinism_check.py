"""Run the resident packed score path several times on the same columns and report whether the seven output arrays are
bit-identical from run to run (both routes).  usage: python tools/determinism_check.py [p]"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from staarpipeline_b200 import api, synth  # noqa: E402

import torch  # noqa: E402

p = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
n = 100_000
ctx = api.Context(0)
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
names = ("AF", "MAF", "Score", "Score_se", "pvalue_log", "Est", "Est_se")
for path in ("split", "fused"):
    ctx.set_packed_path(path)
    ref = None
    for it in range(3):
        outs = torch.zeros((7, p), dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        ctx.packed_score_device(cols.data_ptr(), stride, n, p, api.IMPUTE_MEAN, [outs[i].data_ptr() for i in range(7)])
        ctx.synchronize()
        a = outs.cpu().numpy()
        h = hashlib.sha256(a.tobytes()).hexdigest()[:16] + " first60k " + hashlib.sha256(np.ascontiguousarray(a[:, :60000]).tobytes()).hexdigest()[:16]
        if ref is None:
            ref = a
        diff = {names[i]: int((a[i].view(np.int64) != ref[i].view(np.int64)).sum()) for i in range(7)}
        print(path, it, h, diff, flush=True)
        if it and any(diff.values()):
            i = [k for k in range(7) if diff[names[k]]][0]
            j = int(np.flatnonzero(a[i].view(np.int64) != ref[i].view(np.int64))[0])
            print("   first difference:", names[i], "variant", j, repr(a[i][j]), repr(ref[i][j]), "AF", a[0][j])
ctx.close()
