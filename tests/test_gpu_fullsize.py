"""-m gpu: parity at the sample sizes BASELINE.json quotes (configs 3, 4, 5), where the n-dependent kernel choices —
scan CTA width, SPA cluster width, carrier vs P*G route, column length vs shared memory — are the ones the bench
numbers come from; plus the fused single-pass route against the two-kernel route, the q > 13 fallback of the chunk
driver and more shapes for the multi-trait conditional test.  A few dozen variants each, oracle = CPU port."""
import numpy as np
import pytest
import scipy.sparse as sp

import cases
from util import assert_spa_pvalues_close, assert_stats_close, rel_err

pytestmark = pytest.mark.gpu

KEYS5 = ("Score", "Score_se", "pvalue_log", "Est", "Est_se")


def score_floor(G, residuals):
    G = G.toarray() if sp.issparse(G) else np.asarray(G)
    return 1e-3 * np.sqrt(((G * np.asarray(residuals)[:, None]) ** 2).sum(axis=0))


def _codes(n, first, p):
    from staarpipeline_b200 import synth
    codes = synth.dosage_codes(n, first, p)
    return codes, synth.codes_to_dosage(codes), synth.pack_codes(np.asfortranarray(codes))


def _fast_dense_P(n, q, seed):
    """symmetric dense n x n 'projection' and residuals without an O(n^3) inverse (n = 20,000 -> 3.2 GB):
    P = D - W W' / r with a rank-64 W; any symmetric P exercises the same kernels"""
    rng = np.random.default_rng(seed)
    W = rng.normal(size=(n, 64))
    P = -(W @ W.T) / 64.0
    P[np.diag_indices(n)] += 3.0 + rng.random(n)
    P = 0.5 * (P + P.T)
    return np.asfortranarray(P), rng.normal(size=n)


def test_config3_full_size_dense_grm(gpu_ctx, oracle):
    """config 3 at n = 20,000 (BASELINE configs[2]): the frozen-signature dense-GRM test (dense G -> P*G DMMA GEMM; sparse
    G -> carrier form) and the resident 2-bit path (carrier route for rare, decoded P*G batches for common variants)"""
    import torch
    from staarpipeline_b200 import api
    n, p = 20_000, 20
    P, res = _fast_dense_P(n, 13, 3)
    codes, G0, packed = _codes(n, 5000, p)
    fl = oracle.matrix_flip_mean(G0)
    want = oracle.Individual_Score_Test_denseGRM(fl["Geno"], P, res)
    floor = score_floor(fl["Geno"], res)
    nnz = (fl["Geno"] != 0).sum(axis=0)
    assert (nnz <= n // 8).any() and (nnz > n // 8).any()                      # both routes of the resident path
    # frozen signatures: dense G, and the dgCMatrix of the rare columns
    got = gpu_ctx.Individual_Score_Test_denseGRM(fl["Geno"], P, res)
    assert_stats_close(got, want, floor)
    rare = np.flatnonzero(nnz <= n // 8)
    got_sp = gpu_ctx.Individual_Score_Test_sp_denseGRM(sp.csc_matrix(fl["Geno"][:, rare]), P, res)
    assert_stats_close(got_sp, {k: want[k][rare] for k in want}, floor[rare])
    # resident path
    gpu_ctx.set_nullmodel_dense(P, res)
    dev = torch.device("cuda:0")
    d_cols = torch.from_numpy(packed).to(dev)
    outs = torch.zeros((7, p), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    gpu_ctx.packed_score_dense_device(d_cols.data_ptr(), packed.shape[1], n, p, api.IMPUTE_MEAN, [outs[i].data_ptr() for i in range(7)])
    r = outs.cpu().numpy()
    assert np.array_equal(r[0], fl["AF"]) and np.array_equal(r[1], fl["MAF"])
    assert_stats_close(dict(zip(KEYS5, r[2:])), want, floor)


def test_config4_full_size_three_traits(gpu_ctx, oracle):
    """config 4 at n = 100,000, k = 3 (BASELINE configs[3]): resident multi-trait path (one CTA per variant, carrier list
    read once for all traits) against Individual_Score_Test_sp_multi on Diagonal(3) (x) Geno"""
    import torch
    from staarpipeline_b200 import api, synth
    n, k, pp = 100_000, 3, 16
    nm = synth.nullmodel_multi(n, k=k, seed=synth.BASE_SEED, q=13)
    codes, G0, packed = _codes(n, 9000, pp)
    fl = oracle.matrix_flip_mean(G0)
    GK = sp.kron(sp.identity(k, format="csc"), sp.csc_matrix(fl["Geno"]), format="csc")
    want = oracle.Individual_Score_Test_sp_multi(GK, nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals, k)
    gpu_ctx.set_nullmodel_sparse(nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
    dev = torch.device("cuda:0")
    d_cols = torch.from_numpy(packed).to(dev)
    af = torch.zeros(pp, dtype=torch.float64, device=dev); maf = torch.zeros_like(af)
    score = torch.zeros(pp * k, dtype=torch.float64, device=dev); pl = torch.zeros(pp, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    gpu_ctx.packed_score_multi_device(d_cols.data_ptr(), packed.shape[1], n, pp, k, api.IMPUTE_MEAN, af.data_ptr(), maf.data_ptr(),
                                      score.data_ptr(), pl.data_ptr())
    assert np.array_equal(af.cpu().numpy(), fl["AF"]) and np.array_equal(maf.cpu().numpy(), fl["MAF"])
    assert rel_err(score.cpu().numpy(), want["Score"], score_floor(GK, nm.residuals)) < 1e-10
    assert np.all(np.abs(pl.cpu().numpy() - want["pvalue_log"]) <= 1e-8 * np.maximum(1.0, np.abs(want["pvalue_log"])))


def test_config5_full_size_binary_trait(gpu_ctx, oracle):
    """config 5 at n = 400,000 (BASELINE configs[4]): packed score test (8-warp scan CTAs, 100 KB columns), conditional
    test on 10 known variants (X_adj n x 23) on selected resident columns, SPA on >= 8 selected variants (six-wide
    clusters)"""
    import torch
    from staarpipeline_b200 import api, synth
    n, p, q = 400_000, 24, 13
    nb = synth.nullmodel_binary(n, seed=synth.BASE_SEED, q=q)
    codes, G0, packed = _codes(n, 20000, p)
    rng = np.random.default_rng(5)
    case_idx = np.flatnonzero(nb.extra["y"] == 1)
    for c in range(3):                                                       # planted signals (Newton-Raphson branch)
        take = rng.choice(case_idx, size=int(len(case_idx) * (0.1 + 0.05 * c)), replace=False)
        codes[take, c] = np.where(codes[take, c] == 3, 3, np.maximum(codes[take, c].astype(np.int64) - 1, 0)).astype(np.uint8)
    G0 = synth.codes_to_dosage(codes)
    packed = synth.pack_codes(np.asfortranarray(codes))
    stride = packed.shape[1]
    fl = oracle.matrix_flip_mean(G0)
    # stage 1: score test, Sigma_i diagonal
    want = oracle.Individual_Score_Test(fl["Geno"], nb.Sigma_i, nb.Sigma_iX, nb.cov, nb.residuals)
    gpu_ctx.set_nullmodel_sparse(nb.Sigma_i, nb.Sigma_iX, nb.cov, nb.residuals)
    for path in ("split", "fused"):
        gpu_ctx.set_packed_path(path)
        got = gpu_ctx.packed_score_host(packed, n, api.IMPUTE_MEAN)
        assert np.array_equal(got["AF"], fl["AF"]) and np.array_equal(got["MAF"], fl["MAF"]), path
        assert_stats_close({k: got[k] for k in want}, want, score_floor(fl["Geno"], nb.residuals))
    gpu_ctx.set_packed_path("split")
    # SPA stage on selected resident columns
    dev = torch.device("cuda:0")
    gpu_ctx.set_nullmodel_spa(nb.XW, nb.XXWX_inv, nb.residuals, nb.muhat)
    d_in = torch.from_numpy(packed).to(dev); d_cols = torch.zeros_like(d_in)
    torch.cuda.synchronize()
    gpu_ctx.packed_reorder_device(d_in.data_ptr(), d_cols.data_ptr(), stride, n, p)
    sel = np.array([0, 1, 2, 5, 8, 13, 17, 21, 23], dtype=np.int64)
    d_sel = torch.from_numpy(sel).to(dev)
    d_p = torch.full((sel.size,), -1.0, dtype=torch.float64, device=dev)
    d_tr = torch.zeros((sel.size, 4), dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    gpu_ctx.packed_spa_device(d_cols.data_ptr(), stride, n, d_sel.data_ptr(), sel.size, api.IMPUTE_MEAN, cases.TOL_SPA,
                              cases.MAX_ITER, d_p.data_ptr(), d_tr.data_ptr())
    want_p, otr = oracle.Individual_Score_Test_SPA(np.asfortranarray(fl["Geno"][:, sel]), nb.XW, nb.XXWX_inv, nb.residuals,
                                                   nb.muhat, cases.TOL_SPA, cases.MAX_ITER, return_trace=True)
    assert_spa_pvalues_close(d_p.cpu().numpy(), want_p)
    assert np.array_equal(d_tr.cpu().numpy()[:, 3], otr[:, 3])
    # conditional test: 10 known variants + 13 covariates
    known = np.arange(10, 20)
    X_adj, res_c = synth.conditional_design(nb, fl["Geno"][:, known], "optimal")
    gpu_ctx.set_nullmodel_sparse(nb.Sigma_i, nb.Sigma_iX, nb.cov, res_c)
    tested = np.array([0, 1, 2, 3, 4, 5, 20, 21, 22, 23], dtype=np.int64)
    d_t = torch.from_numpy(tested).to(dev)
    torch.cuda.synchronize()
    gotc = gpu_ctx.packed_cond(d_cols.data_ptr(), stride, n, d_t.data_ptr(), tested.size, api.IMPUTE_MEAN, X_adj)
    wantc = oracle.Individual_Score_Test_cond(np.asfortranarray(fl["Geno"][:, tested]), nb.Sigma_i, nb.Sigma_iX, nb.cov, X_adj, res_c)
    assert_stats_close(gotc, wantc, score_floor(fl["Geno"][:, tested], res_c))


@pytest.mark.parametrize("n,p,q,related,impute,miss", [(1000, 70, 13, True, "mean", None), (777, 40, 5, False, "minor", None),
                                                     (4099, 200, 13, True, "mean", 0.03), (20000, 150, 7, True, "minor", 0.002)])
def test_fused_route_matches_split_route_and_oracle(gpu_ctx, oracle, n, p, q, related, impute, miss):
    """the single-pass kernel (fused_tc.cu) against the two-kernel route and the oracle: AF / MAF bit-exact, statistics at
    the stated tolerance; heavy missingness sends variants through the redo list (sub-list overflow)"""
    from staarpipeline_b200 import api, synth
    nm = synth.nullmodel_sparse_grm(n, seed=n, q=q, shuffle=True) if related else synth.nullmodel_unrelated(n, seed=n, q=q)
    G0 = cases.raw_dosage(n, p, n % 50)
    if miss:
        G0[np.random.default_rng(n).random(G0.shape) < miss] = np.nan
    codes = np.where(np.isnan(G0) | (G0 <= -1), 3, G0).astype(np.uint8)
    packed = synth.pack_codes(np.asfortranarray(codes))
    flag = api.IMPUTE_MEAN if impute == "mean" else api.IMPUTE_MINOR
    fl = getattr(oracle, f"matrix_flip_{impute}")(G0)
    want = oracle.Individual_Score_Test(fl["Geno"], nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
    gpu_ctx.set_nullmodel_sparse(nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
    res = {}
    try:
        for path in ("split", "fused"):
            gpu_ctx.set_packed_path(path)
            res[path] = gpu_ctx.packed_score_host(packed, n, flag)
            assert gpu_ctx.last_packed_path_was_fused == (path == "fused")
            assert np.array_equal(res[path]["AF"], fl["AF"]) and np.array_equal(res[path]["MAF"], fl["MAF"]), path
            assert_stats_close({k: res[path][k] for k in want}, want, score_floor(fl["Geno"], nm.residuals))
        if miss:
            assert gpu_ctx.packed_debug_fused(packed.shape[0])["redo_count"] > 0
        # twice through the fused route: bit-identical (no floating-point atomics, fixed accumulation order)
        again = gpu_ctx.packed_score_host(packed, n, flag)
        for k in again:
            assert np.array_equal(again[k], res["fused"][k], equal_nan=True), k
    finally:
        gpu_ctx.set_packed_path("split")


def test_fused_route_falls_back_for_large_families(gpu_ctx, oracle):
    """a family of more than 16 members needs relatives' genotypes from anywhere in the column: the fused route must
    decline (two-kernel route, generic record path) and still be right"""
    from staarpipeline_b200 import api, synth
    n, q = 1200, 3
    rng = np.random.default_rng(8)
    perm = rng.permutation(n)
    rows, cols, vals, pos = [], [], [], 0
    for s_ in [2] * 30 + [3] * 10 + [25]:
        A = rng.normal(size=(s_, s_)); inv = np.linalg.inv(A @ A.T / s_ + np.eye(s_))
        idx = perm[pos:pos + s_]; pos += s_
        r, c = np.meshgrid(idx, idx, indexing="ij")
        rows.append(r.ravel()); cols.append(c.ravel()); vals.append(inv.ravel())
    for j in perm[pos:]:
        rows.append(np.array([j])); cols.append(np.array([j])); vals.append(np.array([1.5]))
    Sigma_i = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
    X = synth.covariates(n, rng, q)
    SX = np.asfortranarray(Sigma_i @ X); cov = np.asfortranarray(np.linalg.inv(X.T @ SX))
    res = np.asarray(Sigma_i @ rng.normal(size=n)).ravel()
    G0 = cases.raw_dosage(n, 40, seed=4)
    codes = np.where(np.isnan(G0) | (G0 <= -1), 3, G0).astype(np.uint8)
    packed = synth.pack_codes(np.asfortranarray(codes))
    gpu_ctx.set_nullmodel_sparse(Sigma_i, SX, cov, res)
    try:
        gpu_ctx.set_packed_path("fused")
        got = gpu_ctx.packed_score_host(packed, n, api.IMPUTE_MEAN)
        assert not gpu_ctx.last_packed_path_was_fused
    finally:
        gpu_ctx.set_packed_path("split")
    fl = oracle.matrix_flip_mean(G0)
    want = oracle.Individual_Score_Test(fl["Geno"], Sigma_i, SX, cov, res)
    assert_stats_close({k: got[k] for k in want}, want, score_floor(fl["Geno"], res))


def test_chunk_driver_more_than_13_covariates(gpu_ctx, oracle):
    """q = 14 does not fit the 16-column sliced operand of the packed contraction: the chunk driver must take the FP64
    kernels instead and return the same seven arrays (AF / MAF bit-exact)"""
    from staarpipeline_b200 import api, synth
    n, p, q = 3000, 60, 14
    nm = synth.nullmodel_sparse_grm(n, seed=14, q=q, shuffle=True)
    G0 = cases.raw_dosage(n, p, 14)
    fl = oracle.matrix_flip_mean(G0)
    want = oracle.Individual_Score_Test(fl["Geno"], nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
    gpu_ctx.set_nullmodel_sparse(nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
    got = gpu_ctx.individual_analysis_chunk(G0, api.IMPUTE_MEAN)
    assert np.array_equal(got["AF"], fl["AF"]) and np.array_equal(got["MAF"], fl["MAF"])
    assert_stats_close({k: got[k] for k in want}, want, score_floor(fl["Geno"], nm.residuals))


@pytest.mark.parametrize("n,p,k,q,n_known,seed", [(200, 12, 2, 3, 2, 19), (900, 30, 3, 5, 4, 23), (1501, 9, 2, 13, 10, 29)])
def test_cond_multi_more_shapes(gpu_ctx, oracle, n, p, k, q, n_known, seed):
    """Individual_Score_Test_cond_multi (a9) beyond the one golden case: three traits, 10 known variants, ragged n"""
    c, _ = cases.case_cond_multi(oracle, n=n, p=p, k=k, q=q, n_known=n_known, seed=seed)
    got = gpu_ctx.Individual_Score_Test_cond_multi(**c)
    want = oracle.Individual_Score_Test_cond_multi(**c)
    assert rel_err(got["Score"], want["Score"], score_floor(c["G"], c["residuals"])) < 1e-10
    assert np.all(np.abs(got["pvalue_log"] - want["pvalue_log"]) <= 1e-8 * np.maximum(1.0, np.abs(want["pvalue_log"])))


@pytest.mark.parametrize("n,n_known,method", [(3000, 4, "optimal"), (3000, 3, "naive"), (100_000, 10, "optimal")])
def test_conditional_refit_on_device_vs_lm(gpu_ctx, oracle, n, n_known, method):
    """f3: the refit of R/Individual_Analysis_cond.R:187-197 on the device (known variants decoded from resident columns,
    least squares by scaled normal equations + refinement) followed by the conditional test of a whole group, against
    lm()-equivalent QR least squares (oracle/np_oracle.py) + Individual_Score_Test_cond of the oracle"""
    import torch
    from oracle import np_oracle
    from staarpipeline_b200 import api, synth
    q, p = 13, 40
    nm = synth.nullmodel_sparse_grm(n, seed=n % 91, q=q, shuffle=True)
    codes, G0, packed = _codes(n, 3000, p)
    stride = packed.shape[1]
    fl = oracle.matrix_flip_mean(G0)
    # known loci: the most common variants of the block (a rare known variant would make a near-singular design)
    order = np.argsort(-fl["MAF"])
    known = np.sort(order[:n_known]).astype(np.int64)
    tested = np.setdiff1d(np.arange(p), known)[:20].astype(np.int64)
    X_adj, res_c, coef = np_oracle.conditional_refit(nm.residuals, fl["Geno"][:, known], nm.X, method)
    want = oracle.Individual_Score_Test_cond(np.asfortranarray(fl["Geno"][:, tested]), nm.Sigma_i, nm.Sigma_iX, nm.cov, X_adj, res_c)
    gpu_ctx.set_nullmodel_sparse(nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)          # the ORIGINAL scaled residuals
    dev = torch.device("cuda:0")
    d_in = torch.from_numpy(packed).to(dev); d_cols = torch.zeros_like(d_in)
    torch.cuda.synchronize()
    gpu_ctx.packed_reorder_device(d_in.data_ptr(), d_cols.data_ptr(), stride, n, p)
    d_t, d_k = torch.from_numpy(tested).to(dev), torch.from_numpy(known).to(dev)
    torch.cuda.synchronize()
    got = gpu_ctx.packed_cond_refit(d_cols.data_ptr(), stride, n, d_t.data_ptr(), tested.size, d_k.data_ptr(), known.size,
                                    api.IMPUTE_MEAN, nm.X, method)
    assert rel_err(got["coef"], coef, 1e-3 * np.abs(coef).max()) < 1e-9
    assert_stats_close({k: got[k] for k in want}, want, score_floor(fl["Geno"][:, tested], res_c))


@pytest.mark.parametrize("n,p,q,related,impute", [(20000, 160, 13, True, "mean"), (33000, 96, 5, False, "minor"),
                                                 (5000, 120, 7, True, "mean")])
def test_missing_lists_route_matches_detection_route(monkeypatch, oracle, n, p, q, related, impute):
    """Two-kernel route, missing samples listed by the contraction kernel (default) against the scan's own detection pass
    (STAAR_B200_MISSING_LISTS=0) and the oracle.  Missing rates per variant run from 0 to 60 %, so some lists fit and
    some are truncated (those variants take the detection path inside the default flow); one variant is all missing.
    AF / MAF bit-exact; statistics at the stated tolerance; twice through the default flow bit-identical."""
    from staarpipeline_b200 import api, synth
    nm = synth.nullmodel_sparse_grm(n, seed=n, q=q, shuffle=True) if related else synth.nullmodel_unrelated(n, seed=n, q=q)
    G0 = cases.raw_dosage(n, p, n % 37)
    p = G0.shape[1]                                                    # raw_dosage appends its edge-case columns
    rng = np.random.default_rng(n + 1)
    rates = np.concatenate([[0.0, 0.0005, 0.002, 0.01, 0.05, 0.3, 0.6], rng.choice([0.0, 0.001, 0.004, 0.02], size=p - 7)])
    G0[rng.random(G0.shape) < rates[None, :]] = np.nan
    G0[:, p - 1] = np.nan
    codes = np.where(np.isnan(G0) | (G0 <= -1), 3, G0).astype(np.uint8)
    packed = synth.pack_codes(np.asfortranarray(codes))
    flag = api.IMPUTE_MEAN if impute == "mean" else api.IMPUTE_MINOR
    fl = getattr(oracle, f"matrix_flip_{impute}")(G0)
    want = oracle.Individual_Score_Test(fl["Geno"], nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
    res = {}
    for lists in ("1", "0"):
        monkeypatch.setenv("STAAR_B200_MISSING_LISTS", lists)
        ctx = api.Context(0)
        try:
            ctx.set_nullmodel_sparse(nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
            res[lists] = ctx.packed_score_host(packed, n, flag)
            if lists == "1":
                again = ctx.packed_score_host(packed, n, flag)
                for k in again:
                    assert np.array_equal(again[k], res[lists][k], equal_nan=True), k
        finally:
            ctx.close()
        assert np.array_equal(res[lists]["AF"], fl["AF"], equal_nan=True) and np.array_equal(res[lists]["MAF"], fl["MAF"], equal_nan=True)
        assert_stats_close({k: res[lists][k] for k in want}, want, score_floor(fl["Geno"], nm.residuals))
    for k in ("Score_se", "Est_se"):                                   # both flows hit the exact-zero branch on the same variants
        assert np.array_equal(res["1"][k] == 0, res["0"][k] == 0) or np.array_equal(np.isfinite(res["1"][k]), np.isfinite(res["0"][k])), k


def test_prefix_staged_scan_odd_long_column(gpu_ctx, oracle):
    """n = 90 001 (not a multiple of 4 / 16 / 64, long enough for the list-mode scan to stage only the kinship prefix of a
    column and read the rest from global memory): tail masks of the direct loads, padding, missing samples on both sides
    of the table's end"""
    from staarpipeline_b200 import api, synth
    n, p, q = 90001, 40, 13
    nm = synth.nullmodel_sparse_grm(n, seed=5, q=q, shuffle=True)
    G0 = cases.raw_dosage(n, p, 5)
    rng = np.random.default_rng(6)
    G0[rng.random(G0.shape) < 0.002] = np.nan
    G0[n - 70:, 3] = np.nan                                             # a run of missing samples in the last vectors
    G0[n - 1, :8] = 2.0                                                 # carriers in the very last position
    codes = np.where(np.isnan(G0) | (G0 <= -1), 3, G0).astype(np.uint8)
    packed = synth.pack_codes(np.asfortranarray(codes))
    for impute, flag in (("mean", api.IMPUTE_MEAN), ("minor", api.IMPUTE_MINOR)):
        fl = getattr(oracle, f"matrix_flip_{impute}")(G0)
        want = oracle.Individual_Score_Test(fl["Geno"], nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
        gpu_ctx.set_nullmodel_sparse(nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
        got = gpu_ctx.packed_score_host(packed, n, flag)
        assert np.array_equal(got["AF"], fl["AF"], equal_nan=True) and np.array_equal(got["MAF"], fl["MAF"], equal_nan=True)
        assert_stats_close({k: got[k] for k in want}, want, score_floor(fl["Geno"], nm.residuals))
