"""The oracle port against the committed golden vectors (outputs of the reference build, see
tests/golden/make_golden.py).  Runs anywhere — /root/reference is not needed."""
import importlib.util
import os

import numpy as np
import pytest

from util import rel_err

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(spec)
spec.loader.exec_module(make_golden)

GOLD = np.load(os.path.join(HERE, "golden", "reference_outputs.npz"))


@pytest.fixture(scope="module")
def golden_cases(oracle):
    return make_golden.all_cases(oracle, GOLD)


NAMES = sorted({k.split("/")[0] for k in GOLD.files})


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_golden(oracle, golden_cases, name):
    fn, kw = golden_cases[name]
    want_sha = GOLD[f"{name}/__sha256__"].tobytes().decode()
    assert make_golden.fingerprint_G(kw) == want_sha, "seeded genotypes drifted from the ones the golden outputs were made from"
    res = getattr(oracle, fn)(**kw)
    if not isinstance(res, dict):
        res = {"pvalue": res}
    for k, v in res.items():
        if k.startswith("in/"):
            continue
        want = GOLD[f"{name}/{k}"]
        if name.startswith("flip") or name == "score_spa":
            assert np.array_equal(np.asarray(v), want, equal_nan=True), (name, k)   # bit-exact
        else:
            assert rel_err(v, want, 1e-300) <= 1e-9, (name, k)
