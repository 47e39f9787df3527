"""-m gpu, needs two GPUs (skipped otherwise): several devices behind one C-ABI handle (include/staar_b200.h section C).
Run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def multi():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from staarpipeline_b200.api import MultiContext
    m = MultiContext([0, 1])
    yield m
    m.close()


def test_multi_chunk_matches_single_gpu_bit_for_bit(multi, gpu_ctx, oracle):
    """the chunk driver over two devices (null model built once, cloned device to device; columns sharded) returns
    exactly the arrays of the one-GPU call, and both match the oracle"""
    from staarpipeline_b200 import api, synth
    n, p, q = 20_000, 301, 13
    nm = synth.nullmodel_sparse_grm(n, seed=3, q=q, shuffle=True)
    G0 = cases.raw_dosage(n, p, 3)
    gpu_ctx.set_nullmodel_sparse(nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
    one = gpu_ctx.individual_analysis_chunk(G0, api.IMPUTE_MEAN)
    multi.set_nullmodel_sparse(nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
    two = multi.individual_analysis_chunk(G0, api.IMPUTE_MEAN)
    for k in one:
        assert np.array_equal(one[k], two[k], equal_nan=True), k
    fl = oracle.matrix_flip_mean(G0)
    assert np.array_equal(two["AF"], fl["AF"]) and np.array_equal(two["MAF"], fl["MAF"])


def test_multi_frozen_signatures(multi, gpu_ctx, oracle):
    """matrix_flip_mean and the dense-G Individual_Score_Test with their columns sharded over two devices"""
    c, _ = cases.case_sparse(oracle, n=3000, p=97, q=6, seed=21)
    G0 = cases.raw_dosage(3000, 97, 21)
    a, b = gpu_ctx.matrix_flip_mean(G0), multi.matrix_flip_mean(G0)
    for k in ("Geno", "AF", "MAF"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k
    one, two = gpu_ctx.Individual_Score_Test(**c), multi.Individual_Score_Test(**c)
    for k in one:
        assert np.array_equal(one[k], two[k], equal_nan=True), k
