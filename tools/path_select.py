"""Times the two resident packed-score routes (split: contraction + block-per-variant scan; fused: single pass) on
several null-model shapes (kind:n:variants ...; PATHS=split,fused; STAAR_B200_MISSING_LISTS / STAAR_B200_CONTRACT /
STAAR_B200_SCAN_WARPS select variants of the split route).  The measurements quoted in DESIGN.md section 4."""
import sys, os, time
sys.path.insert(0, ".")
import numpy as np, torch
from staarpipeline_b200 import api, synth
dev = torch.device("cuda:0")
GEN = {"binary": synth.nullmodel_binary, "unrel": synth.nullmodel_unrelated, "sparse": synth.nullmodel_sparse_grm}
def run(kind, n, V, paths=("split", "fused")):
    ctx = api.Context(0)
    nb = GEN[kind](n)
    ctx.set_nullmodel_sparse(nb.Sigma_i, nb.Sigma_iX, nb.cov, nb.residuals)
    stride = api.packed_stride(n)
    prm = synth.variant_params(n, 0, V)
    t = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).to(dev)
    packed = torch.empty((V, stride), dtype=torch.uint8, device=dev)
    th, te, tm, sa = t(prm.thr_hom), t(prm.thr_het), t(prm.thr_miss), t(prm.store_alt)   # kept alive across the call
    order = torch.from_numpy(ctx.nullmodel_sample_order()).to(dev)
    ctx.synth_packed_device(packed.data_ptr(), stride, n, 0, V, synth.BASE_SEED, th.data_ptr(), te.data_ptr(), tm.data_ptr(), sa.data_ptr(), order.data_ptr())
    ctx.synchronize()
    outs = torch.zeros((7, V), dtype=torch.float64, device=dev); ptrs = [outs[i].data_ptr() for i in range(7)]
    ctx.set_profiling(True)
    for pth in paths:
        ctx.set_packed_path(pth)
        for _ in range(3):
            ctx.packed_score_device(packed.data_ptr(), stride, n, V, api.IMPUTE_MEAN, ptrs); ctx.synchronize()
        ms = ctx.last_stage_ms()
        print(kind, n, V, pth, "total %.2f ms" % sum(ms.values()), {k: round(v, 2) for k, v in ms.items()}, "fused" if ctx.last_packed_path_was_fused else "split", flush=True)
    del ctx
if __name__ == "__main__":
    cases = sys.argv[1:] or ["binary:100000:250000", "unrel:100000:250000", "unrel:400000:60000", "sparse:400000:60000", "sparse:200000:120000"]
    for c in cases:
        kind, n, V = c.split(":")
        run(kind, int(n), int(V), tuple(os.environ.get("PATHS", "split,fused").split(",")))
