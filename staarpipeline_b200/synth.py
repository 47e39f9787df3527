"""Deterministic synthetic "GDS" genotypes and null-model objects.

SeqArray/GMMAT are external to the reference and not installable here, so the
inputs the hot path consumes are generated:

* genotypes with ``$dosage`` semantics (count of REF alleles, NA = missing;
  reference R/Individual_Analysis.R:186,199) from a counter-based hash RNG, so
  the SAME bits come out of the numpy generator here and of the device
  generator kernel (csrc/synth.cu) — parity subsets of a device-resident batch
  can be regenerated on the host;
* the null-model objects each BASELINE.json config needs (what
  R/fit_nullmodel.R:71-176 / GMMAT::glmmkin would hand to the score tests).

Nothing here is on the measured path.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

BASE_SEED = 20251019
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


# --------------------------------------------------------------------------
# counter-based RNG shared with csrc/synth.cu (keep the constants in sync)
# --------------------------------------------------------------------------
def hash_u32(seed: int, variant: np.ndarray, sample: np.ndarray, stream: int = 0) -> np.ndarray:
    """splitmix64-style mix of (seed, variant, sample, stream) -> uint32 (top 32 bits)."""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) + np.uint64(stream) * np.uint64(0xD6E8FEB86659FD93)
             + variant.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
             + sample.astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F))
        z ^= z >> np.uint64(30)
        z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27)
        z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    return (z >> np.uint64(32)).astype(np.uint32)


@dataclass
class VariantParams:
    """Per-variant integer thresholds consumed by both generators (exact integer compares)."""
    thr_hom: np.ndarray      # uint64 in [0, 2^32]: u < thr_hom            -> alt count 2
    thr_het: np.ndarray      # uint64:               u < thr_het (>= hom)  -> alt count 1
    thr_miss: np.ndarray     # uint64:               u2 < thr_miss         -> missing
    store_alt: np.ndarray    # uint8: 1 -> store alt count itself instead of REF dosage 2-alt
    alt_freq: np.ndarray     # float64 (informational)


def variant_params(n: int, first_variant: int, p: int, seed: int = BASE_SEED, mac_min: int = 20) -> VariantParams:
    """ALT-allele frequency f ~ 0.5*Beta(0.15,4) truncated to expected MAC >= mac_min; HWE genotypes;
    95% of variants stored as REF dosage (so they need the flip), 5% stored as ALT count;
    1% of variants carry 2% missing entries, the rest 0.1%."""
    v = np.arange(first_variant, first_variant + p, dtype=np.uint64)
    zero = np.zeros(p, dtype=np.uint64)
    u_f = (hash_u32(seed, v, zero, 1).astype(np.float64) + 0.5) / 2.0 ** 32
    u_m = hash_u32(seed, v, zero, 2).astype(np.float64) / 2.0 ** 32
    u_s = hash_u32(seed, v, zero, 3).astype(np.float64) / 2.0 ** 32
    lo = min(0.5, (mac_min + 1.0) / (2.0 * n))
    grid_u, grid_f = _beta_table()
    c_lo = float(np.interp(2.0 * lo, grid_f, grid_u))
    f = 0.5 * np.interp(c_lo + u_f * (1.0 - c_lo), grid_u, grid_f)
    f = np.clip(f, lo, 0.5)
    two32 = 2.0 ** 32
    thr_hom = np.floor(f * f * two32).astype(np.uint64)
    thr_het = np.floor((f * f + 2.0 * f * (1.0 - f)) * two32).astype(np.uint64)
    miss_rate = np.where(u_m < 0.01, 0.02, 0.001)
    thr_miss = np.floor(miss_rate * two32).astype(np.uint64)
    store_alt = (u_s < 0.05).astype(np.uint8)
    return VariantParams(thr_hom, thr_het, thr_miss, store_alt, f)


_BETA_TABLE = None


def _beta_table():
    """inverse CDF of Beta(0.15, 4) on a fixed grid (beta.ppf per variant is far too slow for 1e7 variants)"""
    global _BETA_TABLE
    if _BETA_TABLE is None:
        from scipy.stats import beta
        f = np.concatenate([[0.0], np.logspace(-9, 0, 4096)])
        _BETA_TABLE = (beta.cdf(f, 0.15, 4.0), f)
    return _BETA_TABLE


def dosage_codes(n: int, first_variant: int, p: int, seed: int = BASE_SEED,
                 params: VariantParams | None = None) -> np.ndarray:
    """(n, p) uint8 codes: 0/1/2 = stored dosage, 3 = missing — what the device generator packs."""
    if params is None:
        params = variant_params(n, first_variant, p, seed)
    v = np.arange(first_variant, first_variant + p, dtype=np.uint64)[None, :]
    j = np.arange(n, dtype=np.uint64)[:, None]
    u = hash_u32(seed, v, j, 0).astype(np.uint64)
    u2 = hash_u32(seed, v, j, 4).astype(np.uint64)
    alt = (u < params.thr_hom[None, :]).astype(np.uint8) + (u < params.thr_het[None, :]).astype(np.uint8)
    stored = np.where(params.store_alt[None, :] == 1, alt, 2 - alt).astype(np.uint8)
    stored[u2 < params.thr_miss[None, :]] = 3
    return np.asfortranarray(stored)


def codes_to_dosage(codes: np.ndarray) -> np.ndarray:
    """uint8 codes -> the FP64 matrix seqGetData('$dosage') would return (NaN = NA)."""
    G = codes.astype(np.float64)
    G[codes == 3] = np.nan
    return np.asfortranarray(G)


def pack_codes(codes: np.ndarray, stride_bytes: int | None = None) -> np.ndarray:
    """(n, p) codes -> (p, stride_bytes) uint8, variant-major, sample j in byte j//4 bits 2*(j%4)."""
    n, p = codes.shape
    if stride_bytes is None:
        stride_bytes = packed_stride(n)
    padded = np.zeros((stride_bytes * 4, p), dtype=np.uint8)
    padded[:n] = codes
    q = padded.reshape(stride_bytes, 4, p)
    b = q[:, 0] | (q[:, 1] << 2) | (q[:, 2] << 4) | (q[:, 3] << 6)
    return np.ascontiguousarray(b.T)


def packed_stride(n: int) -> int:
    """bytes per packed variant column: ceil(n/4) rounded up to 16 B (TMA / 128-bit loads)."""
    return ((n + 3) // 4 + 15) // 16 * 16


def adversarial_columns(n: int, rng: np.random.Generator) -> np.ndarray:
    """Columns the reference's branches care about: all missing; monomorphic 0 and 2;
    AF exactly 0.5; AF one ulp-ish above 0.5; a single carrier; missing coded as a <= -1 sentinel."""
    cols = []
    cols.append(np.full(n, np.nan))                                  # all missing -> AF stays 0
    cols.append(np.zeros(n))                                         # monomorphic 0
    cols.append(np.full(n, 2.0))                                     # monomorphic 2 -> flips to all 0
    half = np.zeros(n); half[: n // 2] = 2.0
    if n % 2:                                                        # AF == 0.5 exactly needs an even count
        half[-1] = np.nan
    cols.append(half)                                                # AF exactly 0.5 -> no flip
    above = half.copy(); above[n // 2] = 1.0                          # AF just above 0.5 -> flip
    cols.append(above)
    single = np.zeros(n); single[rng.integers(n)] = 1.0
    cols.append(single)                                              # singleton
    sent = rng.integers(0, 3, n).astype(np.float64); sent[rng.random(n) < 0.1] = -9.0
    cols.append(sent)                                                # <= -1 sentinel counts as missing
    mix = rng.integers(0, 3, n).astype(np.float64); mix[rng.random(n) < 0.3] = np.nan
    cols.append(mix)                                                 # heavy missingness
    return np.asfortranarray(np.stack(cols, axis=1))


# --------------------------------------------------------------------------
# null models
# --------------------------------------------------------------------------
def _single_threaded_blas(fn):
    """The generators' dense products (X'Sigma_iX, Sigma_i (y - X b), ...) go through BLAS, whose summation order depends
    on the thread count — torchrun sets OMP_NUM_THREADS=1, a plain `python bench.py` does not — so the same seed gave
    residuals that differed in the last bit between a 1-GPU and an N-GPU run.  One BLAS thread makes the generated
    null model bit-identical everywhere (the operands are small: n x q)."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        try:
            from threadpoolctl import threadpool_limits
        except ImportError:                                   # pragma: no cover
            return fn(*args, **kwargs)
        with threadpool_limits(limits=1):
            return fn(*args, **kwargs)
    return wrapped


@dataclass
class NullModel:
    n: int
    kind: str                       # "sparse" | "dense" | "multi" | "binary"
    X: np.ndarray
    residuals: np.ndarray
    Sigma_i: sp.csc_matrix | None = None
    Sigma_iX: np.ndarray | None = None
    cov: np.ndarray | None = None
    P: np.ndarray | None = None
    n_pheno: int = 1
    muhat: np.ndarray | None = None
    XW: np.ndarray | None = None
    XXWX_inv: np.ndarray | None = None
    extra: dict = field(default_factory=dict)


def covariates(n: int, rng: np.random.Generator, q: int = 13) -> np.ndarray:
    """intercept, N(50,10^2) 'age', Bernoulli(.5) 'sex', q-3 N(0,1) PCs."""
    X = np.empty((n, q))
    X[:, 0] = 1.0
    if q > 1:
        X[:, 1] = rng.normal(50.0, 10.0, n)
    if q > 2:
        X[:, 2] = rng.integers(0, 2, n)
    if q > 3:
        X[:, 3:] = rng.normal(size=(n, q - 3))
    return np.asfortranarray(X)


def _beta(q: int) -> np.ndarray:
    b = np.zeros(q)
    b[0] = 1.0
    if q > 1:
        b[1] = 0.02
    if q > 2:
        b[2] = 0.3
    if q > 3:
        b[3:] = 0.05
    return b


def family_blocks(n: int, rng: np.random.Generator):
    """Block-diagonal 2*Phi: family sizes {1:70%,2:15%,3:8%,4:5%,6:2%}, entries 0.5 / 0.25.
    Returns a list of (start, size) and the list of 2*Phi blocks."""
    sizes, probs = np.array([1, 2, 3, 4, 6]), np.array([0.70, 0.15, 0.08, 0.05, 0.02])
    starts, blocks, pos = [], [], 0
    draws = rng.choice(sizes, size=n, p=probs)
    for s in draws:
        s = int(min(s, n - pos))
        if s <= 0:
            break
        if s == 6:
            A = 0.5 * np.eye(3) + 0.5 * np.ones((3, 3))
            B = 0.25 * np.ones((3, 3))
            blk = np.block([[A, B], [B, A]])
        else:
            blk = 0.5 * np.eye(s) + 0.5 * np.ones((s, s))
        starts.append((pos, s)); blocks.append(blk); pos += s
        if pos >= n:
            break
    return starts, blocks


@_single_threaded_blas
def nullmodel_unrelated(n: int, seed: int = BASE_SEED, q: int = 13) -> NullModel:
    """cfg1: continuous trait, unrelated samples: Sigma_i = I/sigma^2 as a diagonal sparse matrix."""
    rng = np.random.default_rng(seed)
    X = covariates(n, rng, q)
    y = X @ _beta(q) + rng.normal(0.0, 1.5, n)
    bhat, *_ = np.linalg.lstsq(X, y, rcond=None)
    res = y - X @ bhat
    s2 = float(res @ res) / (n - q)
    Sigma_i = sp.identity(n, format="csc") / s2
    Sigma_iX = np.asfortranarray(X / s2)
    cov = np.linalg.inv(X.T @ Sigma_iX)
    return NullModel(n, "sparse", X, res / s2, Sigma_i=sp.csc_matrix(Sigma_i), Sigma_iX=Sigma_iX,
                     cov=np.asfortranarray(cov))


@_single_threaded_blas
def nullmodel_sparse_grm(n: int, seed: int = BASE_SEED, q: int = 13, tau_g: float = 0.6, tau_e: float = 0.9,
                         shuffle: bool = False) -> NullModel:
    """cfg2: Sigma = tau_g*2Phi + tau_e*I, block-diagonal kinship; Sigma_i is the exact block inverse."""
    rng = np.random.default_rng(seed)
    X = covariates(n, rng, q)
    starts, blocks = family_blocks(n, rng)
    rows, cols, vals = [], [], []
    e = np.empty(n)
    for (s0, s), blk in zip(starts, blocks):
        Sig = tau_g * blk + tau_e * np.eye(s)
        L = np.linalg.cholesky(Sig)
        e[s0:s0 + s] = L @ rng.normal(size=s)
        inv = np.linalg.inv(Sig)
        r, c = np.meshgrid(np.arange(s0, s0 + s), np.arange(s0, s0 + s), indexing="ij")
        rows.append(r.ravel()); cols.append(c.ravel()); vals.append(inv.ravel())
    rows, cols, vals = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
    if shuffle:                      # real cohorts are not family-sorted
        perm = rng.permutation(n)
        rows, cols = perm[rows], perm[cols]
        e_s = np.empty(n); e_s[perm] = e; e = e_s
    Sigma_i = sp.csc_matrix((vals, (rows, cols)), shape=(n, n))
    Sigma_i.sort_indices()
    y = X @ _beta(q) + e
    Sigma_iX = np.asfortranarray(Sigma_i @ X)
    cov = np.linalg.inv(X.T @ Sigma_iX)
    bhat = cov @ (Sigma_iX.T @ y)
    res = Sigma_i @ (y - X @ bhat)
    return NullModel(n, "sparse", X, np.asarray(res).ravel(), Sigma_i=Sigma_i, Sigma_iX=Sigma_iX,
                     cov=np.asfortranarray(cov))


@_single_threaded_blas
def nullmodel_dense_grm(n: int, seed: int = BASE_SEED, q: int = 13, markers: int = 2000,
                        tau_g: float = 0.6, tau_e: float = 0.9) -> NullModel:
    """cfg3: Phi = ZZ'/M + 0.5 I; P = Sigma^-1 - Sigma^-1 X cov X' Sigma^-1 (dense n x n)."""
    rng = np.random.default_rng(seed)
    X = covariates(n, rng, q)
    Z = rng.normal(size=(n, markers))
    Phi = Z @ Z.T / markers + 0.5 * np.eye(n)
    Sig = tau_g * Phi + tau_e * np.eye(n)
    L = np.linalg.cholesky(Sig)
    y = X @ _beta(q) + L @ rng.normal(size=n)
    Si = np.linalg.inv(Sig)
    Si = 0.5 * (Si + Si.T)
    SiX = Si @ X
    cov = np.linalg.inv(X.T @ SiX)
    P = Si - SiX @ cov @ SiX.T
    P = 0.5 * (P + P.T)
    return NullModel(n, "dense", X, P @ y, P=np.asfortranarray(P))


@_single_threaded_blas
def nullmodel_multi(n: int, k: int = 3, seed: int = BASE_SEED, q: int = 13) -> NullModel:
    """cfg4: k correlated traits; Sigma = Sg (x) 2Phi + Se (x) I, trait-major stacking (nk x nk sparse)."""
    rng = np.random.default_rng(seed)
    X = covariates(n, rng, q)
    starts, blocks = family_blocks(n, rng)
    corr = np.array([1.0, 0.5, 0.3])
    Se = np.array([[corr[abs(a - b)] if abs(a - b) < 3 else 0.1 for b in range(k)] for a in range(k)]) * 0.9
    Sg = 0.6 * (0.7 * np.eye(k) + 0.3 * np.ones((k, k)))
    rows, cols, vals = [], [], []
    E = np.empty((n, k))
    for (s0, s), blk in zip(starts, blocks):
        Sig = np.kron(Sg, blk) + np.kron(Se, np.eye(s))          # trait-major within the family
        L = np.linalg.cholesky(Sig)
        E[s0:s0 + s, :] = (L @ rng.normal(size=s * k)).reshape(k, s).T
        inv = np.linalg.inv(Sig)
        idx = (np.arange(k)[:, None] * n + np.arange(s0, s0 + s)[None, :]).ravel()
        r, c = np.meshgrid(idx, idx, indexing="ij")
        rows.append(r.ravel()); cols.append(c.ravel()); vals.append(inv.ravel())
    Sigma_i = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n * k, n * k))
    Sigma_i.sort_indices()
    Xk = np.kron(np.eye(k), X)                                      # nk x kq
    B = np.stack([_beta(q) * (1 + 0.1 * a) for a in range(k)], axis=1)
    Y = X @ B + E
    y = Y.T.ravel()                                                 # as.vector of n x k = trait-major
    Sigma_iX = np.asfortranarray(Sigma_i @ Xk)
    cov = np.linalg.inv(Xk.T @ Sigma_iX)
    bhat = cov @ (Sigma_iX.T @ y)
    res = Sigma_i @ (y - Xk @ bhat)
    return NullModel(n, "multi", X, np.asarray(res).ravel(), Sigma_i=Sigma_i, Sigma_iX=Sigma_iX,
                     cov=np.asfortranarray(cov), n_pheno=k)


@_single_threaded_blas
def nullmodel_binary(n: int, seed: int = BASE_SEED, q: int = 13, intercept: float = -4.7) -> NullModel:
    """cfg5: imbalanced (~1:100) case-control, logistic null fitted by IRLS; SPA operands as in
    R/fit_nullmodel.R:153-162 (XW = t(X*w), XXWX_inv = X (X'WX)^-1)."""
    rng = np.random.default_rng(seed)
    X = covariates(n, rng, q)
    Xs = X.copy()
    if q > 1:
        Xs[:, 1] = (Xs[:, 1] - 50.0) / 10.0
    b = _beta(q) * 0.5
    b[0] = intercept
    eta = Xs @ b
    y = (rng.random(n) < 1.0 / (1.0 + np.exp(-eta))).astype(np.float64)
    bh = np.zeros(q); bh[0] = np.log(max(y.mean(), 1e-6) / (1 - min(y.mean(), 1 - 1e-6)))
    for _ in range(50):
        mu = 1.0 / (1.0 + np.exp(-(Xs @ bh)))
        w = mu * (1 - mu)
        step = np.linalg.solve(Xs.T @ (Xs * w[:, None]), Xs.T @ (y - mu))
        bh += step
        if np.max(np.abs(step)) < 1e-12:
            break
    mu = 1.0 / (1.0 + np.exp(-(Xs @ bh)))
    w = mu * (1 - mu)
    XtWX_inv = np.linalg.inv(Xs.T @ (Xs * w[:, None]))
    nm = NullModel(n, "binary", np.asfortranarray(Xs), y - mu,
                   Sigma_i=sp.diags(w, format="csc"), Sigma_iX=np.asfortranarray(Xs * w[:, None]),
                   cov=np.asfortranarray(XtWX_inv), muhat=mu,
                   XW=np.asfortranarray((Xs * w[:, None]).T), XXWX_inv=np.asfortranarray(Xs @ XtWX_inv))
    nm.extra["y"] = y
    return nm


@_single_threaded_blas
def conditional_design(nm: NullModel, G_known: np.ndarray, method: str = "optimal"):
    """X_adj and refit residuals as R/Individual_Analysis_cond.R:190-197 builds them with lm()."""
    n = nm.n
    if method == "optimal":
        A = np.column_stack([G_known, nm.X])
    else:
        A = np.column_stack([np.ones(n), G_known])
    coef, *_ = np.linalg.lstsq(A, nm.residuals, rcond=None)
    return np.asfortranarray(A), nm.residuals - A @ coef
