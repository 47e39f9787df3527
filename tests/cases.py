"""Seeded small inputs for every function on the hot path.

Shared by the oracle-vs-reference tests, the golden-vector generator
(tests/golden/make_golden.py) and the GPU parity tests, so that all three look
at byte-identical inputs.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from staarpipeline_b200 import synth

TOL_SPA = float(np.finfo(np.float64).eps ** 0.25)     # R/Individual_Analysis.R default tol
MAX_ITER = 1000


def raw_dosage(n: int, p: int, seed: int = 7, adversarial: bool = True) -> np.ndarray:
    """n x (p [+8]) FP64 `$dosage` matrix with NaN missing and the adversarial columns appended."""
    G = synth.codes_to_dosage(synth.dosage_codes(n, 0, p, seed=synth.BASE_SEED + seed))
    if adversarial:
        G = np.column_stack([G, synth.adversarial_columns(n, np.random.default_rng(seed))])
    return np.asfortranarray(G)


def flipped(oracle, n: int, p: int, seed: int = 7, adversarial: bool = True) -> np.ndarray:
    return oracle.matrix_flip_mean(raw_dosage(n, p, seed, adversarial))["Geno"]


def case_sparse(oracle, n=600, p=48, q=6, seed=11, shuffle=True):
    nm = synth.nullmodel_sparse_grm(n, seed=seed, q=q, shuffle=shuffle)
    G = flipped(oracle, n, p, seed)
    return dict(G=G, Sigma_i=nm.Sigma_i, Sigma_iX=nm.Sigma_iX, cov=nm.cov, residuals=nm.residuals), nm


def case_unrelated(oracle, n=800, p=64, q=13, seed=3):
    nm = synth.nullmodel_unrelated(n, seed=seed, q=q)
    G = flipped(oracle, n, p, seed)
    return dict(G=G, Sigma_i=nm.Sigma_i, Sigma_iX=nm.Sigma_iX, cov=nm.cov, residuals=nm.residuals), nm


def case_dense(oracle, n=300, p=40, q=5, seed=5):
    nm = synth.nullmodel_dense_grm(n, seed=seed, q=q, markers=150)
    G = flipped(oracle, n, p, seed)
    return dict(G=G, P=nm.P, residuals=nm.residuals), nm


def case_multi(oracle, n=240, p=24, k=3, q=4, seed=9):
    nm = synth.nullmodel_multi(n, k=k, seed=seed, q=q)
    G0 = flipped(oracle, n, p, seed)
    Gk = sp.kron(sp.identity(k), sp.csc_matrix(G0), format="csc")       # Diagonal(k) %x% G  (R/Individual_Analysis.R:267)
    return dict(G=Gk, Sigma_i=nm.Sigma_i, Sigma_iX=nm.Sigma_iX, cov=nm.cov, residuals=nm.residuals, n_pheno=k), nm, G0


def case_dense_multi(oracle, n=120, p=16, k=2, seed=13):
    rng = np.random.default_rng(seed)
    G0 = flipped(oracle, n, p, seed)
    Gk = sp.kron(sp.identity(k), sp.csc_matrix(G0), format="csc")
    B = rng.normal(size=(n * k, n * k + 8))
    P = B @ B.T / (n * k)                                                  # any SPD nk x nk "projection"
    res = rng.normal(size=n * k)
    return dict(G=Gk, P=np.asfortranarray(P), residuals=res, n_pheno=k), G0


def case_cond(oracle, n=500, p=30, q=5, n_known=4, seed=17, method="optimal"):
    nm = synth.nullmodel_sparse_grm(n, seed=seed, q=q, shuffle=True)
    G = flipped(oracle, n, p + n_known, seed, adversarial=False)
    A, res2 = synth.conditional_design(nm, G[:, :n_known], method)
    return dict(G=np.asfortranarray(G[:, n_known:]), Sigma_i=nm.Sigma_i, Sigma_iX=nm.Sigma_iX, cov=nm.cov,
                X_adj=A, residuals=res2), nm


def case_cond_multi(oracle, n=200, p=12, k=2, q=3, n_known=2, seed=19):
    nm = synth.nullmodel_multi(n, k=k, seed=seed, q=q)
    G = flipped(oracle, n, p + n_known, seed, adversarial=False)
    A0 = np.column_stack([np.ones(n), G[:, :n_known]])
    Ak = np.kron(np.eye(k), A0)                                            # Diagonal(k) %x% X_adj (Individual_Analysis_cond.R:200)
    coef, *_ = np.linalg.lstsq(Ak, nm.residuals, rcond=None)
    Gk = np.kron(np.eye(k), G[:, n_known:])
    return dict(G=np.asfortranarray(Gk), Sigma_i=nm.Sigma_i, Sigma_iX=nm.Sigma_iX, cov=nm.cov,
                X_adj=np.asfortranarray(Ak), residuals=nm.residuals - Ak @ coef, n_pheno=k), nm


def case_spa(oracle, n=4000, p=40, q=5, seed=23, intercept=-3.5, planted=6):
    """binary trait (~3% cases here; cfg5 uses 1%), SPA on all columns; a few planted strong signals
    so both the NR path and tiny p-values are exercised."""
    nm = synth.nullmodel_binary(n, seed=seed, q=q, intercept=intercept)
    G = flipped(oracle, n, p, seed, adversarial=False)
    y = nm.extra["y"]
    rng = np.random.default_rng(seed)
    cases = np.flatnonzero(y == 1)
    for c in range(min(planted, p)):                                       # enrich carriers among cases
        take = rng.choice(cases, size=max(2, int(len(cases) * (0.15 + 0.1 * c))), replace=False)
        G[take, c] = np.minimum(G[take, c] + 1.0, 2.0)
    G[:, p - 1] = 0.0                                                       # monomorphic column -> p = 1 path
    return dict(G=np.asfortranarray(G), XW=nm.XW, XXWX_inv=nm.XXWX_inv, residuals=nm.residuals, muhat=nm.muhat,
                tol=TOL_SPA, max_iter=MAX_ITER), nm
