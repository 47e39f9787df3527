"""Short run of the resident packed score path for ncu: python tools/fused_prof.py [p] [path]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from staarpipeline_b200 import api, synth  # noqa: E402

import torch  # noqa: E402

p = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 256
path = sys.argv[2] if len(sys.argv) > 2 else "fused"
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
outs = torch.zeros((7, p), dtype=torch.float64, device=dev)
ctx.set_packed_path(path)
ctx.set_profiling(True)
for it in range(3):
    ctx.packed_score_device(cols.data_ptr(), stride, n, p, api.IMPUTE_MEAN, [outs[i].data_ptr() for i in range(7)])
    ctx.synchronize()
print(path, p, ctx.last_stage_ms())
if int(os.environ.get("STAAR_B200_FUSED_DBG", "0")) & 128:
    import ctypes as C
    ctas = (p + 255) // 256
    buf = np.zeros((ctas, 32, 8), dtype=np.uint32)
    ctx._call("staar_packed_debug_timers", C.c_int64(ctas), buf.ctypes.data_as(C.POINTER(C.c_uint32)))
    c = buf.reshape(-1)[:4].astype(np.int64)
    print(f"cells listed {c[0]} ({c[0] / p:.1f} per variant), served in place {c[1]}, batches {c[2]}, k-blocks with cells {c[3]}")
    m = buf.astype(np.float64).mean(axis=0) / 1e3
    np.set_printoptions(linewidth=200, suppress=True, precision=0)
    print("kclk per phase, mean over CTAs; expander warps 0-7: [wait a_empty, wait g_full, work]")
    print(m[:8, :3])
    print("scanner warps 8-23: [wait g_full, phase1, -, consume+push, barrier, batches, loop head, tail]")
    print(m[8:24])
    print("mma issuer: [wait b_full, wait a_full, issue]", m[24, :3], " genotype producer [wait g_empty, issue]", m[26, :2])
    print("max over CTAs scanner total", buf[:, 8:24, :7].sum(axis=2).max() / 1e3, " mean", buf[:, 8:24, :7].sum(axis=2).mean() / 1e3)
ctx.close()
