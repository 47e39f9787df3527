"""Independent numpy route for the score tests — TEST INFRASTRUCTURE, NOT product code.

Evaluates the reference's Armadillo expressions LITERALLY, including the full
p x p ``Cov`` matrix (reference src/Individual_Score_Test.cpp:39-43), with
numpy/scipy matmul.  Used only to show that the C oracle's diagonal-only
evaluation equals the diagonal of the literal expression
and to give the chi-square tail a second opinion (mpmath).  Small shapes only.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def pchisq_upper_log_mp(x: float, df: float, dps: int = 50) -> float:
    """log of the chi-square upper tail by mpmath (ground truth for the Rmath substitute)."""
    import mpmath as mp
    with mp.workdps(dps):
        q = mp.gammainc(mp.mpf(df) / 2, mp.mpf(x) / 2, mp.inf, regularized=True)
        return float(mp.log(q))


def pnorm_mp(x: float, lower: bool = True, dps: int = 50) -> float:
    import mpmath as mp
    with mp.workdps(dps):
        return float(mp.ncdf(x) if lower else mp.ncdf(-x))


def cov_full(G, Sigma_i, Sigma_iX, cov):
    """Cov = (G' Sigma_i) G - (Sigma_iX' G)' cov (Sigma_iX' G)      [Individual_Score_Test.cpp:42-43]"""
    G = np.asarray(G.todense()) if sp.issparse(G) else np.asarray(G)
    t = Sigma_iX.T @ G
    return (Sigma_i.T @ G).T @ G - t.T @ cov @ t


def cov_full_dense_P(G, P):
    """Cov = (P G)' G                                                 [Individual_Score_Test_denseGRM.cpp:37]"""
    G = np.asarray(G.todense()) if sp.issparse(G) else np.asarray(G)
    return (P @ G).T @ G


def cov_full_cond(G, Sigma_i, Sigma_iX, cov, A):
    """the 5-term conditional covariance                               [Individual_Score_Test_cond.cpp:48-49]"""
    PA = Sigma_i @ A - Sigma_iX @ cov @ (Sigma_iX.T @ A)
    M = np.linalg.inv(A.T @ A)
    SG = Sigma_i @ G
    return (SG.T @ G - (Sigma_iX.T @ G).T @ cov @ Sigma_iX.T @ G
            - (A.T @ G).T @ M @ (PA.T @ G) - (PA.T @ G).T @ M @ (A.T @ G)
            + (A.T @ G).T @ M @ PA.T @ A @ M @ (A.T @ G))


def single_from_cov(U, Cov, logtail):
    d = np.diag(Cov)
    out = {k: np.zeros_like(d) for k in ("Score_se", "pvalue_log", "Est", "Est_se")}
    nz = d != 0
    out["Score"] = U
    out["pvalue_log"][nz] = [-logtail(u * u / c, 1.0) for u, c in zip(U[nz], d[nz])]
    out["Score_se"][nz] = np.sqrt(d[nz])
    out["Est"][nz] = U[nz] / d[nz]
    out["Est_se"][nz] = 1.0 / np.sqrt(d[nz])
    return out


def multi_from_cov(U, Cov, k, logtail):
    pp = U.size // k
    pl = np.zeros(pp)
    for i in range(pp):
        idx = np.arange(k) * pp + i
        C = Cov[np.ix_(idx, idx)]
        if np.linalg.det(C) == 0:
            continue
        pl[i] = -logtail(float(U[idx] @ np.linalg.inv(C) @ U[idx]), float(k))
    return {"Score": U, "pvalue_log": pl}


def conditional_refit(residuals, G_known, X, method: str = "optimal"):
    """The refit R does per tested variant before Individual_Score_Test_cond (R/Individual_Analysis_cond.R:187-197):
    lm(scaled.residuals ~ genotype_adj + X - 1) ("optimal") or lm(scaled.residuals ~ genotype_adj) (intercept);
    returns (X_adj = model.matrix, residuals of the fit, coefficients).  lm() solves by Householder QR; numpy's
    lstsq (SVD) is the same least-squares solution to rounding for a full-rank design."""
    residuals = np.asarray(residuals, dtype=np.float64).ravel()
    G_known = np.asarray(G_known, dtype=np.float64).reshape(residuals.size, -1)
    if method == "optimal":
        A = np.column_stack([G_known, np.asarray(X, dtype=np.float64)])
    else:
        A = np.column_stack([np.ones(residuals.size), G_known])
    Q, R = np.linalg.qr(A)                                    # Householder QR, as lm()
    coef = np.linalg.solve(R, Q.T @ residuals)
    return np.asfortranarray(A), residuals - A @ coef, coef
