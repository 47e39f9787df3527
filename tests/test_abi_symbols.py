"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol that
include/staar_b200.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "staar_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(staar_[a-zA-Z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from staarpipeline_b200 import build
    return ctypes.CDLL(build.build())


def test_header_declares_the_13_reference_entry_points():
    syms = declared_symbols()
    for name in ("matrix_flip_mean", "matrix_flip_minor", "individual_score_test", "individual_score_test_sp",
                 "individual_score_test_denseGRM", "individual_score_test_sp_denseGRM", "individual_score_test_multi",
                 "individual_score_test_sp_multi", "individual_score_test_denseGRM_multi",
                 "individual_score_test_sp_denseGRM_multi", "individual_score_test_cond",
                 "individual_score_test_cond_multi", "individual_score_test_SPA"):
        assert "staar_" + name in syms


def test_library_exports_every_declared_symbol(lib):
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ctx = ctypes.c_void_p()
    rc = lib.staar_ctx_create(0, ctypes.byref(ctx))
    lib.staar_last_error.restype = ctypes.c_char_p
    assert rc == 2 and b"no CPU fallback" in lib.staar_last_error()


def test_packed_stride(lib):
    lib.staar_packed_stride.restype = ctypes.c_size_t
    lib.staar_packed_stride.argtypes = [ctypes.c_int64]
    assert lib.staar_packed_stride(1) == 16 and lib.staar_packed_stride(64) == 16 and lib.staar_packed_stride(65) == 32
    assert lib.staar_packed_stride(100000) == 25008


def test_host_packer_cpu():
    import numpy as np
    import cases
    from staarpipeline_b200 import api, synth
    G0 = cases.raw_dosage(130, 9)
    codes = np.where(np.isnan(G0) | (G0 <= -1), 3, G0).astype(np.uint8)
    assert np.array_equal(api.pack_dosage_host(G0, threads=2), synth.pack_codes(np.asfortranarray(codes)))
