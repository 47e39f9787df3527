"""Properties that tie the oracle's diagonal-only evaluation to the reference's literal full-Cov
expressions and to each other."""
import numpy as np
import scipy.sparse as sp

import cases
from oracle import np_oracle
from util import assert_stats_close, rel_err


def test_diag_equals_full_cov_sparse(oracle):
    c, _ = cases.case_sparse(oracle, n=300, p=20)
    got = oracle.Individual_Score_Test(**c)
    Cov = np_oracle.cov_full(c["G"], c["Sigma_i"], c["Sigma_iX"], c["cov"])
    want = np_oracle.single_from_cov(c["residuals"] @ c["G"], Cov, oracle.pchisq_upper_log)
    keep = np.abs(np.diag(Cov)) > 1e-9 * np.abs(np.einsum("ji,ji->i", c["G"], c["Sigma_i"] @ c["G"])).max()
    for k in want:
        assert rel_err(np.asarray(got[k])[keep], np.asarray(want[k])[keep], 1e-300) < 1e-9, k


def test_diag_equals_full_cov_dense(oracle):
    c, _ = cases.case_dense(oracle, n=200, p=16)
    got = oracle.Individual_Score_Test_denseGRM(**c)
    Cov = np_oracle.cov_full_dense_P(c["G"], c["P"])
    want = np_oracle.single_from_cov(c["residuals"] @ c["G"], Cov, oracle.pchisq_upper_log)
    keep = np.diag(Cov) > 1e-9
    for k in want:
        assert rel_err(np.asarray(got[k])[keep], np.asarray(want[k])[keep], 1e-300) < 1e-9, k


def test_cond_equals_literal_five_terms(oracle):
    c, _ = cases.case_cond(oracle, n=300, p=12)
    got = oracle.Individual_Score_Test_cond(**c)
    Cov = np_oracle.cov_full_cond(c["G"], c["Sigma_i"], c["Sigma_iX"], c["cov"], c["X_adj"])
    want = np_oracle.single_from_cov(c["residuals"] @ c["G"], Cov, oracle.pchisq_upper_log)
    for k in want:
        assert rel_err(got[k], want[k], 1e-300) < 1e-8, k


def test_multi_equals_literal(oracle):
    c, _, _ = cases.case_multi(oracle, n=120, p=10)
    got = oracle.Individual_Score_Test_sp_multi(**c)
    Gd = c["G"].toarray()
    Cov = np_oracle.cov_full(Gd, c["Sigma_i"], c["Sigma_iX"], c["cov"])
    want = np_oracle.multi_from_cov(c["residuals"] @ Gd, Cov, c["n_pheno"], oracle.pchisq_upper_log)
    d = np.diag(Cov)[: Gd.shape[1] // c["n_pheno"]]
    keep = d > 1e-9
    assert rel_err(got["Score"], want["Score"], 1e-300) < 1e-10
    assert rel_err(got["pvalue_log"][keep], want["pvalue_log"][keep], 1e-300) < 1e-8


def test_dense_and_sparse_G_agree(oracle):
    c, _ = cases.case_sparse(oracle)
    a = oracle.Individual_Score_Test(**c)
    b = oracle.Individual_Score_Test_sp(**dict(c, G=sp.csc_matrix(c["G"])))
    assert_stats_close(a, b)


def test_flip_invariance_of_variance(oracle):
    """Cov_ii is invariant under g -> 2-g when the intercept is in X (projection kills constants);
    Score changes sign up to the (zero) residual sum."""
    c, nm = cases.case_unrelated(oracle, n=500, p=30)
    a = oracle.Individual_Score_Test(**c)
    b = oracle.Individual_Score_Test(**dict(c, G=np.asfortranarray(2.0 - c["G"])))
    keep = a["Score_se"] > 1e-6
    assert rel_err(b["Score_se"][keep], a["Score_se"][keep]) < 1e-8
    assert np.allclose(b["Score"], 2 * nm.residuals.sum() - a["Score"], rtol=0, atol=1e-9)


def test_flip_semantics(oracle):
    n = 50
    G = cases.raw_dosage(n, 10)
    m = oracle.matrix_flip_mean(G)
    assert np.all(m["MAF"] <= 0.5) and np.all(m["MAF"] >= 0)
    assert not np.isnan(m["Geno"]).any()
    # all-missing column: AF stays 0, imputed to 0, no flip
    j = 10
    assert m["AF"][j] == 0 and m["MAF"][j] == 0 and np.all(m["Geno"][:, j] == 0)
    # AF exactly 0.5 is NOT flipped; just above is
    assert m["AF"][13] == 0.5 and np.array_equal(m["Geno"][:, 13][~np.isnan(G[:, 13])], G[:, 13][~np.isnan(G[:, 13])])
    assert m["AF"][14] > 0.5 and m["MAF"][14] == 1 - m["AF"][14]
    # minor-imputation recomputes AF over all n
    mi = oracle.matrix_flip_minor(G)
    assert np.all(np.isin(mi["Geno"], (0.0, 1.0, 2.0)))
    assert np.allclose(mi["AF"][:10] <= 0.5, m["AF"][:10] <= 0.5)


def test_spa_trace_probe(oracle):
    """E_v of SURVEY §8d: CGF evaluations per variant stay in the tens; NR converges in a few steps."""
    c, _ = cases.case_spa(oracle, n=3000, p=24)
    p, tr = oracle.Individual_Score_Test_SPA(**c, return_trace=True)
    assert np.all((p >= 0) & (p <= 1))
    tested = p < 1.0
    assert tested.sum() >= 10
    evals = tr[tested, :3].sum(axis=1)
    assert evals.min() >= 10 and np.median(evals) < 80


def test_conditional_refit_is_the_least_squares_fit():
    """np_oracle.conditional_refit (the lm() refit of R/Individual_Analysis_cond.R:187-197, checker of the device-side
    refit): residuals orthogonal to the design, equal to the SVD least-squares solution, both model formulas."""
    from staarpipeline_b200 import synth
    n = 2500
    nm = synth.nullmodel_unrelated(n, seed=77, q=6)
    rng = np.random.default_rng(77)
    G_known = rng.binomial(2, 0.2, size=(n, 4)).astype(np.float64)
    for method in ("optimal", "naive"):
        A, res, coef = np_oracle.conditional_refit(nm.residuals, G_known, nm.X, method)
        assert A.shape == (n, 4 + (nm.X.shape[1] if method == "optimal" else 1))
        assert np.abs(A.T @ res).max() < 1e-9 * np.abs(A).max() * np.abs(nm.residuals).max() * n
        A2, res2 = synth.conditional_design(nm, G_known, method)
        assert np.array_equal(A, A2)
        assert rel_err(res, res2, 1e-300) < 1e-10
        want, *_ = np.linalg.lstsq(A, nm.residuals, rcond=None)
        assert rel_err(coef, want, 1e-300) < 1e-9
