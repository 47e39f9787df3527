"""Generate the committed golden vectors under tests/golden/.

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden.py
Outputs come from oracle/_ref — the REFERENCE's own src/*.cpp compiled against the stand-in
headers (oracle/ref_stub/) — not from the oracle port, so the port is checked against them
(tests/test_golden.py) and they travel to the GPU box where /root/reference does not exist.
Genotypes are regenerated from seeds by tests/cases.py (integer hash RNG: identical on every machine) and
fingerprinted (sha256); every other operand (null-model matrices, residuals, X_adj, ...) comes out of
LAPACK/BLAS calls whose last bits depend on the CPU, so those inputs are STORED in the file next to the
outputs and loaded back by `all_cases(oracle, gold)`.
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


def _store_inputs(blob, name, kw):
    for k, v in kw.items():
        if k == "G":
            continue
        if sp.issparse(v):
            v = sp.csc_matrix(v); v.sort_indices()
            blob[f"{name}/in/{k}/data"] = v.data
            blob[f"{name}/in/{k}/indices"] = v.indices.astype(np.int64)
            blob[f"{name}/in/{k}/indptr"] = v.indptr.astype(np.int64)
            blob[f"{name}/in/{k}/shape"] = np.asarray(v.shape, dtype=np.int64)
        else:
            blob[f"{name}/in/{k}"] = np.asarray(v)


def _load_inputs(gold, name, kw):
    """replace every non-genotype operand of `kw` by the stored one"""
    out = dict(kw)
    for k in kw:
        if k == "G":
            continue
        if f"{name}/in/{k}/data" in gold.files:
            shape = tuple(int(x) for x in gold[f"{name}/in/{k}/shape"])
            out[k] = sp.csc_matrix((gold[f"{name}/in/{k}/data"], gold[f"{name}/in/{k}/indices"],
                                    gold[f"{name}/in/{k}/indptr"]), shape=shape)
        elif f"{name}/in/{k}" in gold.files:
            v = gold[f"{name}/in/{k}"]
            out[k] = v.item() if v.ndim == 0 else v
    return out


def all_cases(oracle, gold=None):
    """name -> (function name on the Oracle object, kwargs); with `gold`, non-genotype operands come from the file"""
    out = _all_cases(oracle)
    if gold is not None:
        out = {name: (fn, _load_inputs(gold, name, kw)) for name, (fn, kw) in out.items()}
    return out


def fingerprint_G(kw) -> str:
    return fingerprint({"G": kw["G"]})


def _all_cases(oracle):
    out = {}
    G0 = cases.raw_dosage(400, 60)
    out["flip_mean"] = ("matrix_flip_mean", dict(G=G0))
    out["flip_minor"] = ("matrix_flip_minor", dict(G=G0))
    c, _ = cases.case_sparse(oracle)
    out["score_sparse"] = ("Individual_Score_Test", c)
    out["score_sparse_sp"] = ("Individual_Score_Test_sp", dict(c, G=sp.csc_matrix(c["G"])))
    c, _ = cases.case_unrelated(oracle)
    out["score_unrelated"] = ("Individual_Score_Test", c)
    c, _ = cases.case_dense(oracle, n=120, p=24)
    out["score_dense"] = ("Individual_Score_Test_denseGRM", c)
    out["score_dense_sp"] = ("Individual_Score_Test_sp_denseGRM", dict(c, G=sp.csc_matrix(c["G"])))
    c, _, _ = cases.case_multi(oracle)
    out["score_multi_sp"] = ("Individual_Score_Test_sp_multi", c)
    out["score_multi"] = ("Individual_Score_Test_multi", dict(c, G=c["G"].toarray()))
    c, _ = cases.case_dense_multi(oracle, n=60, p=10)
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
        blob[f"{name}/__sha256__"] = np.frombuffer(fingerprint_G(kw).encode(), dtype=np.uint8)
        _store_inputs(blob, name, kw)
        print(f"{name:24s} {fn:42s} " + " ".join(f"{k}{tuple(np.shape(v))}" for k, v in res.items()))
    path = os.path.join(HERE, "reference_outputs.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
