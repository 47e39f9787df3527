"""Scan / contraction time as a function of allele frequency (tools; not part of the product)."""
import json, sys, os
import numpy as np
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
import torch
from staarpipeline_b200 import api, synth
dev = torch.device("cuda:0"); ctx = api.Context(0)
n, V = 100_000, 100_000
nm = synth.nullmodel_sparse_grm(n, seed=synth.BASE_SEED, q=13, shuffle=True)
ctx.set_nullmodel_sparse(nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
stride = api.packed_stride(n)
order = torch.from_numpy(ctx.nullmodel_sample_order()).to(dev)
packed = torch.empty((V, stride), dtype=torch.uint8, device=dev)
outs = torch.zeros((7, V), dtype=torch.float64, device=dev); ptrs = [outs[i].data_ptr() for i in range(7)]
for f in (0.0003, 0.003, 0.01, 0.03, 0.1, 0.3):
    for miss in (0.001, 0.0):
        two32 = 2.0 ** 32
        th = torch.full((V,), int(f * f * two32), dtype=torch.int64, device=dev)
        te = torch.full((V,), int((f * f + 2 * f * (1 - f)) * two32), dtype=torch.int64, device=dev)
        tm = torch.full((V,), int(miss * two32), dtype=torch.int64, device=dev)
        sa = torch.zeros((V,), dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        ctx.synth_packed_device(packed.data_ptr(), stride, n, 0, V, 1, th.data_ptr(), te.data_ptr(), tm.data_ptr(), sa.data_ptr(), order.data_ptr())
        ctx.set_profiling(True)
        for _ in range(2):
            ctx.packed_score_device(packed.data_ptr(), stride, n, V, api.IMPUTE_MEAN, ptrs)
        st = ctx.last_stage_ms(); ctx.set_profiling(False)
        print(json.dumps({"f": f, "missing": miss, "scan_ns_per_variant": st["scan"] * 1e6 / V, "contract_ns_per_variant": st["contract"] * 1e6 / V}), flush=True)
