"""Per-call times of the R driver's native call sequence through the frozen-signature entry points (tools).
usage: python tools/frozen_sequence_time.py [n] [p]"""
import sys, time, os
import numpy as np, scipy.sparse as sp
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
from staarpipeline_b200 import api, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
ctx = api.Context(0)
nm = synth.nullmodel_sparse_grm(n, q=13, shuffle=True)
G = np.asfortranarray(synth.codes_to_dosage(synth.dosage_codes(n, 0, p)))
for rep in range(3):
    t0 = time.perf_counter(); fl = ctx.matrix_flip_mean(G); t1 = time.perf_counter()
    maf = fl["MAF"]; common = maf >= 0.05; rare = (maf > 19.5 / n / 2) & (maf < 0.05)
    Gc = np.asfortranarray(fl["Geno"][:, common]); Gr = sp.csc_matrix(fl["Geno"][:, rare])
    t2 = time.perf_counter(); ctx.Individual_Score_Test(Gc, nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals); t3 = time.perf_counter()
    ctx.Individual_Score_Test_sp(Gr, nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals); t4 = time.perf_counter()
    print(f"flip {1e3*(t1-t0):.1f} ms ({G.nbytes/1e9:.2f} GB each way) | dense test on {int(common.sum())} common {1e3*(t3-t2):.1f} ms | "
          f"sparse test on {int(rare.sum())} rare ({Gr.nnz} entries) {1e3*(t4-t3):.1f} ms", flush=True)
ctx.close()
