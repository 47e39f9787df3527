import sys, numpy as np
sys.path[:0] = ["/root/repo", "/root/repo/tests"]
import cases
from oracle.pyoracle import Oracle
from staarpipeline_b200 import api, synth
orc = Oracle("port")
n, p, q = 777, 40, 5
nm = synth.nullmodel_unrelated(n, seed=n + p, q=q)
G0 = cases.raw_dosage(n, p, n + p)
codes = np.where(np.isnan(G0) | (G0 <= -1), 3, G0).astype(np.uint8)
ctx = api.Context(0)
ctx.set_nullmodel_sparse(nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
got = ctx.packed_score_host(synth.pack_codes(np.asfortranarray(codes)), n, 0)
P = codes.shape[1]
d = ctx.packed_debug_dump(P)
fl = orc.matrix_flip_mean(G0)
Gf = fl["Geno"]
diag = nm.Sigma_i.diagonal()
s2_true = ((Gf == 2) * diag[:, None]).sum(axis=0)
X = d["hi"].astype(np.float64) * 2.0**24 + d["lo"]
print("scale", d["col_scale"][:8])
print("s2 got ", (X[:, 15] * d["col_scale"][q + 1])[:12])
print("s2 true", s2_true[:12])
nonmiss = ~(np.isnan(G0) | (G0 <= -1))
Gint = np.where(nonmiss, Gf, 0.0)
print("dlin got ", (X[:, q + 1] * d["col_scale"][q + 1])[:6])
print("dlin true", (Gint * diag[:, None]).sum(axis=0)[:6])
print("hi15", d["hi"][:6, 15], "lo15", d["lo"][:6, 15])
