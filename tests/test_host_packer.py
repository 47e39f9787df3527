"""Host side of the genotype ingest (no GPU needed): FP64 / int32 / uint8 `$dosage` blocks -> 2-bit packed columns,
AVX-512 and scalar code paths, every byte of the stride written, bad values rejected."""
import os

import numpy as np
import pytest

import cases
from staarpipeline_b200 import api, synth


def _variants(n):
    G0 = cases.raw_dosage(n, 21)
    codes = np.where(np.isnan(G0) | (G0 <= -1), 3, G0).astype(np.uint8)
    ci = codes.astype(np.int32)
    Gi = np.where(ci == 3, np.int32(np.iinfo(np.int32).min), ci).astype(np.int32)
    Gu = np.where(codes == 3, np.uint8(0xFF), codes).astype(np.uint8)
    return G0, Gi, Gu, synth.pack_codes(np.asfortranarray(codes))


@pytest.mark.parametrize("n", [1, 15, 64, 333, 4097])
@pytest.mark.parametrize("simd", [True, False])
def test_pack_matches_numpy(n, simd, monkeypatch):
    if not simd:
        monkeypatch.setenv("STAAR_B200_NO_AVX512", "1")
    G0, Gi, Gu, want = _variants(n)
    for G in (G0, Gi, Gu):
        for threads in (1, 3):
            assert np.array_equal(api.pack_dosage_host(G, threads=threads), want), (G.dtype, threads)


def test_pack_rejects_non_dosage_values():
    G0, Gi, Gu, _ = _variants(200)
    for bad in (G0 * 0.5, np.where(Gi == 1, 7, Gi).astype(np.int32), np.where(Gu == 1, 9, Gu).astype(np.uint8)):
        with pytest.raises(api.StaarError):
            api.pack_dosage_host(bad)
