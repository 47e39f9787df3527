"""Generate the committed golden vectors under tests/golden/.

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden.py
Outputs come from oracle/_ref — the REFERENCE's own src/*.cpp compiled against the stand-in
headers (oracle/ref_stub/) — not from the oracle port, so the port is checked against them
(tests/test_golden.py) and they travel to the GPU box where /root/reference does not exist.
Inputs are regenerated from seeds by tests/cases.py and are fingerprinted (sha256) in the file
so a drift of the generator is detected rather than silently compared against stale outputs.
"""
import hashlib
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import cases  # noqa: E402
from oracle.pyoracle import Oracle, build  # noqa: E402


def fingerprint(case: dict) -> str:
    h = hashlib.sha256()
    for k in sorted(case):
        v = case[k]
        if sp.issparse(v):
            v = sp.csc_matrix(v); v.sort_indices()
            for a in (v.data, v.indices.astype(np.int64), v.indptr.astype(np.int64)):
                h.update(np.ascontiguousarray(a).tobytes())
        else:
            h.update(np.ascontiguousarray(np.asarray(v, dtype=np.float64)).tobytes())
    return h.hexdigest()


def all_cases(oracle):
    """name -> (function name on the Oracle object, kwargs)"""
    out = {}
    G0 = cases.raw_dosage(400, 60)
    out["flip_mean"] = ("matrix_flip_mean", dict(G=G0))
    out["flip_minor"] = ("matrix_flip_minor", dict(G=G0))
    c, _ = cases.case_sparse(oracle)
    out["score_sparse"] = ("Individual_Score_Test", c)
    out["score_sparse_sp"] = ("Individual_Score_Test_sp", dict(c, G=sp.csc_matrix(c["G"])))
    c, _ = cases.case_unrelated(oracle)
    out["score_unrelated"] = ("Individual_Score_Test", c)
    c, _ = cases.case_dense(oracle)
    out["score_dense"] = ("Individual_Score_Test_denseGRM", c)
    out["score_dense_sp"] = ("Individual_Score_Test_sp_denseGRM", dict(c, G=sp.csc_matrix(c["G"])))
    c, _, _ = cases.case_multi(oracle)
    out["score_multi_sp"] = ("Individual_Score_Test_sp_multi", c)
    out["score_multi"] = ("Individual_Score_Test_multi", dict(c, G=c["G"].toarray()))
    c, _ = cases.case_dense_multi(oracle)
    out["score_dense_multi_sp"] = ("Individual_Score_Test_sp_denseGRM_multi", c)
    out["score_dense_multi"] = ("Individual_Score_Test_denseGRM_multi", dict(c, G=c["G"].toarray()))
    c, _ = cases.case_cond(oracle)
    out["score_cond"] = ("Individual_Score_Test_cond", c)
    c, _ = cases.case_cond_multi(oracle)
    out["score_cond_multi"] = ("Individual_Score_Test_cond_multi", c)
    c, _ = cases.case_spa(oracle)
    out["score_spa"] = ("Individual_Score_Test_SPA", c)
    return out


def main():
    build(with_ref=True)
    port, ref = Oracle("port"), Oracle("reference")
    blob = {}
    for name, (fn, kw) in all_cases(port).items():
        res = getattr(ref, fn)(**kw)
        if not isinstance(res, dict):
            res = {"pvalue": res}
        if name.startswith("flip"):
            res = {k: v for k, v in res.items()}
        for k, v in res.items():
            blob[f"{name}/{k}"] = np.asarray(v)
        blob[f"{name}/__sha256__"] = np.frombuffer(fingerprint(kw).encode(), dtype=np.uint8)
        print(f"{name:24s} {fn:42s} " + " ".join(f"{k}{tuple(np.shape(v))}" for k, v in res.items()))
    path = os.path.join(HERE, "reference_outputs.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
