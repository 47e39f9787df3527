"""Host-side mirror of the reference's native interface, on top of the C ABI (include/staar_b200.h).

Function names, argument order/meaning, return conventions and error behaviour follow the reference's exported
R functions (/root/reference/R/RcppExports.R:4-54 -> src/RcppExports.cpp:17-291): each score test returns the
named list ``{Score, Score_se, pvalue_log, Est, Est_se}`` (``pvalue_log`` is -ln p), the multi-trait tests
``{Score, pvalue_log}``, the flips ``{Geno, AF, MAF}``, SPA a bare vector of p-values; a condition under which
Armadillo would throw (dimension mismatch, singular ``inv``) raises ``StaarError`` (the Rcpp shim turns the same
status into ``Rcpp::stop``).  Inputs are numpy arrays / scipy.sparse matrices, the Python stand-ins for
``arma::mat`` / ``arma::sp_mat`` (R itself is absent from this image; the real binding is in rshim/).

Everything computes on the GPU through ``libstaar_b200.so``; there is NO CPU fallback — importing works anywhere
(so the symbol-table tests run on CPU), but creating a :class:`Context` without the library or a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("STAAR_B200_LIB") or os.path.join(_HERE, "libstaar_b200.so")   # override: experiment builds
_dp = C.POINTER(C.c_double)
_u64p = C.POINTER(C.c_uint64)
_u8p = C.POINTER(C.c_uint8)

IMPUTE_MEAN, IMPUTE_MINOR = 0, 1


class StaarError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[staar_b200 status {code}] {msg}")
        self.code = code


_lib = None


def load_library() -> C.CDLL:
    """dlopen the CUDA library; fails loudly when it has not been built (no silent fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise StaarError(2, f"{LIB_PATH} is missing: build it with `python -m staarpipeline_b200.build` "
                                "(this package has no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _lib.staar_last_error.restype = C.c_char_p
        _lib.staar_packed_stride.restype = C.c_size_t
        _lib.staar_packed_stride.argtypes = [C.c_int64]
        _lib.staar_ctx_stream.restype = C.c_void_p
        _lib.staar_ctx_launch_count.restype = C.c_uint64
    return _lib


def _f(a) -> np.ndarray:
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _p(a: np.ndarray):
    return a.ctypes.data_as(_dp)


class _Csc:
    """arma::sp_mat stand-in: CSC with uint64 indices (ARMA_64BIT_WORD, reference src/Makevars:2)."""

    def __init__(self, m):
        m = sp.csc_matrix(m, dtype=np.float64)
        m.sort_indices()
        self.shape = m.shape
        self.values = np.ascontiguousarray(m.data, dtype=np.float64)
        self.row_idx = np.ascontiguousarray(m.indices, dtype=np.uint64)
        self.col_ptr = np.ascontiguousarray(m.indptr, dtype=np.uint64)

    def args(self):
        return (_p(self.values), self.row_idx.ctypes.data_as(_u64p), self.col_ptr.ctypes.data_as(_u64p))


class Context:
    """One context = one GPU = one rank."""

    def __init__(self, device: int = 0):
        self._lib = load_library()
        self._h = C.c_void_p()
        rc = self._lib.staar_ctx_create(C.c_int(device), C.byref(self._h))
        if rc != 0:
            raise StaarError(rc, self._lib.staar_last_error().decode())
        self.device = device

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.staar_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _call(self, name: str, *args):
        rc = getattr(self._lib, name)(self._h, *args)
        if rc != 0:
            raise StaarError(rc, self._lib.staar_last_error().decode())

    @property
    def stream(self) -> int:
        return int(self._lib.staar_ctx_stream(self._h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self._lib.staar_ctx_launch_count(self._h))

    def synchronize(self):
        self._call("staar_ctx_synchronize")

    def set_profiling(self, on: bool):
        self._call("staar_ctx_set_profiling", C.c_int(int(on)))

    def last_stage_ms(self):
        ms = (C.c_float * 3)()
        self._call("staar_ctx_last_stage_ms", ms)
        return {"scan": ms[0], "contract": ms[1], "finalize": ms[2]}

    # ------------------------------------------------------------------ flips
    def _flip(self, fn: str, G, want_geno=True):
        G = _f(G)
        if G.ndim != 2:
            raise StaarError(1, "G must be a matrix")
        n, p = G.shape
        AF, MAF = np.zeros(p), np.zeros(p)
        Geno = np.empty((n, p), order="F") if want_geno else None
        self._call(fn, _p(G), C.c_int64(n), C.c_int64(p), _p(Geno) if want_geno else None, _p(AF), _p(MAF))
        return {"Geno": Geno, "AF": AF, "MAF": MAF}

    def matrix_flip_mean(self, G, want_geno=True):
        return self._flip("staar_matrix_flip_mean", G, want_geno)

    def matrix_flip_minor(self, G, want_geno=True):
        return self._flip("staar_matrix_flip_minor", G, want_geno)

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _sigma_args(n, Sigma_i, Sigma_iX, cov, residuals):
        S = _Csc(Sigma_i)
        if S.shape != (n, n):
            raise StaarError(1, f"Sigma_i is {S.shape}, expected ({n}, {n})")
        SX = _f(Sigma_iX)
        if SX.ndim != 2 or SX.shape[0] != n:
            raise StaarError(1, "matrix multiplication: incompatible matrix dimensions (Sigma_iX)")
        q = SX.shape[1]
        cv = _f(cov)
        if cv.shape != (q, q):
            raise StaarError(1, "matrix multiplication: incompatible matrix dimensions (cov)")
        r = _f(residuals).ravel()
        if r.size != n:
            raise StaarError(1, "matrix multiplication: incompatible matrix dimensions (residuals)")
        keep = (S, SX, cv, r)
        return keep, (*S.args(), _p(SX), C.c_int64(q), _p(cv), _p(r))

    @staticmethod
    def _out5(p):
        o = [np.zeros(p) for _ in range(5)]
        return o, [_p(a) for a in o]

    @staticmethod
    def _pack5(o):
        return dict(zip(("Score", "Score_se", "pvalue_log", "Est", "Est_se"), o))

    # ------------------------------------------------------------------ single trait
    def Individual_Score_Test(self, G, Sigma_i, Sigma_iX, cov, residuals):
        G = _f(G); n, p = G.shape
        keep, sargs = self._sigma_args(n, Sigma_i, Sigma_iX, cov, residuals)
        o, po = self._out5(p)
        self._call("staar_individual_score_test", _p(G), C.c_int64(n), C.c_int64(p), *sargs, *po)
        return self._pack5(o)

    def Individual_Score_Test_sp(self, G, Sigma_i, Sigma_iX, cov, residuals):
        Gs = _Csc(G); n, p = Gs.shape
        keep, sargs = self._sigma_args(n, Sigma_i, Sigma_iX, cov, residuals)
        o, po = self._out5(p)
        self._call("staar_individual_score_test_sp", *Gs.args(), C.c_int64(n), C.c_int64(p), *sargs, *po)
        return self._pack5(o)

    @staticmethod
    def _dense_args(n, P, residuals):
        Pm = _f(P)
        if Pm.shape != (n, n):
            raise StaarError(1, "matrix multiplication: incompatible matrix dimensions (P)")
        r = _f(residuals).ravel()
        if r.size != n:
            raise StaarError(1, "matrix multiplication: incompatible matrix dimensions (residuals)")
        return (Pm, r), (_p(Pm), _p(r))

    def Individual_Score_Test_denseGRM(self, G, P, residuals):
        G = _f(G); n, p = G.shape
        keep, dargs = self._dense_args(n, P, residuals)
        o, po = self._out5(p)
        self._call("staar_individual_score_test_denseGRM", _p(G), C.c_int64(n), C.c_int64(p), *dargs, *po)
        return self._pack5(o)

    def Individual_Score_Test_sp_denseGRM(self, G, P, residuals):
        Gs = _Csc(G); n, p = Gs.shape
        keep, dargs = self._dense_args(n, P, residuals)
        o, po = self._out5(p)
        self._call("staar_individual_score_test_sp_denseGRM", *Gs.args(), C.c_int64(n), C.c_int64(p), *dargs, *po)
        return self._pack5(o)

    # ------------------------------------------------------------------ multi trait
    def Individual_Score_Test_multi(self, G, Sigma_i, Sigma_iX, cov, residuals, n_pheno=1):
        G = _f(G); n, p = G.shape
        keep, sargs = self._sigma_args(n, Sigma_i, Sigma_iX, cov, residuals)
        Score, pl = np.zeros(p), np.zeros(p // max(n_pheno, 1))
        self._call("staar_individual_score_test_multi", _p(G), C.c_int64(n), C.c_int64(p), *sargs, C.c_int(n_pheno),
                   _p(Score), _p(pl))
        return {"Score": Score, "pvalue_log": pl}

    def Individual_Score_Test_sp_multi(self, G, Sigma_i, Sigma_iX, cov, residuals, n_pheno=1):
        Gs = _Csc(G); n, p = Gs.shape
        keep, sargs = self._sigma_args(n, Sigma_i, Sigma_iX, cov, residuals)
        Score, pl = np.zeros(p), np.zeros(p // max(n_pheno, 1))
        self._call("staar_individual_score_test_sp_multi", *Gs.args(), C.c_int64(n), C.c_int64(p), *sargs,
                   C.c_int(n_pheno), _p(Score), _p(pl))
        return {"Score": Score, "pvalue_log": pl}

    def Individual_Score_Test_denseGRM_multi(self, G, P, residuals, n_pheno=1):
        G = _f(G); n, p = G.shape
        keep, dargs = self._dense_args(n, P, residuals)
        Score, pl = np.zeros(p), np.zeros(p // max(n_pheno, 1))
        self._call("staar_individual_score_test_denseGRM_multi", _p(G), C.c_int64(n), C.c_int64(p), *dargs,
                   C.c_int(n_pheno), _p(Score), _p(pl))
        return {"Score": Score, "pvalue_log": pl}

    def Individual_Score_Test_sp_denseGRM_multi(self, G, P, residuals, n_pheno=1):
        Gs = _Csc(G); n, p = Gs.shape
        keep, dargs = self._dense_args(n, P, residuals)
        Score, pl = np.zeros(p), np.zeros(p // max(n_pheno, 1))
        self._call("staar_individual_score_test_sp_denseGRM_multi", *Gs.args(), C.c_int64(n), C.c_int64(p), *dargs,
                   C.c_int(n_pheno), _p(Score), _p(pl))
        return {"Score": Score, "pvalue_log": pl}

    # ------------------------------------------------------------------ conditional
    def Individual_Score_Test_cond(self, G, Sigma_i, Sigma_iX, cov, X_adj, residuals):
        G = _f(G); n, p = G.shape
        keep, sargs = self._sigma_args(n, Sigma_i, Sigma_iX, cov, residuals)
        A = _f(X_adj)
        if A.ndim != 2 or A.shape[0] != n:
            raise StaarError(1, "matrix multiplication: incompatible matrix dimensions (X_adj)")
        o, po = self._out5(p)
        self._call("staar_individual_score_test_cond", _p(G), C.c_int64(n), C.c_int64(p), *sargs[:-1], _p(A),
                   C.c_int64(A.shape[1]), sargs[-1], *po)
        return self._pack5(o)

    def Individual_Score_Test_cond_multi(self, G, Sigma_i, Sigma_iX, cov, X_adj, residuals, n_pheno=1):
        G = _f(G); n, p = G.shape
        keep, sargs = self._sigma_args(n, Sigma_i, Sigma_iX, cov, residuals)
        A = _f(X_adj)
        if A.ndim != 2 or A.shape[0] != n:
            raise StaarError(1, "matrix multiplication: incompatible matrix dimensions (X_adj)")
        Score, pl = np.zeros(p), np.zeros(p // max(n_pheno, 1))
        self._call("staar_individual_score_test_cond_multi", _p(G), C.c_int64(n), C.c_int64(p), *sargs[:-1], _p(A),
                   C.c_int64(A.shape[1]), sargs[-1], C.c_int(n_pheno), _p(Score), _p(pl))
        return {"Score": Score, "pvalue_log": pl}

    # ------------------------------------------------------------------ SPA
    def Individual_Score_Test_SPA(self, G, XW, XXWX_inv, residuals, muhat, tol, max_iter, return_trace=False):
        G = _f(G); n, p = G.shape
        XWm, XXi = _f(XW), _f(XXWX_inv)
        q = XWm.shape[0]
        if XWm.shape != (q, n) or XXi.shape != (n, q):
            raise StaarError(1, "matrix multiplication: incompatible matrix dimensions (XW / XXWX_inv)")
        r, mu = _f(residuals).ravel(), _f(muhat).ravel()
        pv = np.zeros(p)
        trace = np.zeros((p, 4), dtype=np.int64) if return_trace else None
        self._call("staar_individual_score_test_SPA", _p(G), C.c_int64(n), C.c_int64(p), _p(XWm), _p(XXi), C.c_int64(q),
                   _p(r), _p(mu), C.c_double(tol), C.c_int(max_iter), _p(pv),
                   trace.ctypes.data_as(C.POINTER(C.c_int64)) if return_trace else None)
        return (pv, trace) if return_trace else pv

    # ------------------------------------------------------------------ packed fast path
    def set_nullmodel_sparse(self, Sigma_i, Sigma_iX, cov, residuals):
        n = sp.csc_matrix(Sigma_i).shape[0]
        keep, sargs = self._sigma_args(n, Sigma_i, Sigma_iX, cov, residuals)
        S, SX, cv, r = keep
        self._call("staar_nullmodel_set_sparse", C.c_int64(n), C.c_int64(SX.shape[1]), *S.args(), _p(SX), _p(cv), _p(r))
        self._n = n

    def packed_score_host(self, packed: np.ndarray, n: int, impute: int = IMPUTE_MEAN):
        """packed: (p, stride) uint8 host array of 2-bit dosage columns -> dict of 7 arrays (host)."""
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        p, stride = packed.shape
        o = [np.zeros(p) for _ in range(7)]
        self._call("staar_packed_score_host", packed.ctypes.data_as(_u8p), C.c_size_t(stride), C.c_int64(n),
                   C.c_int64(p), C.c_int(impute), *[_p(a) for a in o])
        return dict(zip(("AF", "MAF", "Score", "Score_se", "pvalue_log", "Est", "Est_se"), o))

    def individual_analysis_chunk(self, G, impute: int = IMPUTE_MEAN, threads: int = 0, out=None):
        """`$dosage` host block (n x p, F-order) -> dict of the 7 per-variant arrays.  One call per chunk of the R driver
        (flip + score test).  dtype float64 (NaN missing; what the Rcpp boundary delivers), int32 (R integer matrix,
        NA_integer_ missing) or uint8 (SeqArray raw dosage, 0xFF missing)."""
        G = np.asarray(G)
        if G.dtype == np.int32 or G.dtype == np.uint8:
            G = np.asfortranarray(G)
            fn = "staar_individual_analysis_chunk_i32" if G.dtype == np.int32 else "staar_individual_analysis_chunk_u8"
            gp = G.ctypes.data_as(C.c_void_p)
        else:
            G = _f(G)
            fn, gp = "staar_individual_analysis_chunk", _p(G)
        n, p = G.shape
        o = out if out is not None else [np.zeros(p) for _ in range(7)]
        self._call(fn, gp, C.c_int64(n), C.c_int64(p), C.c_int(impute), C.c_int(threads), *[_p(a) for a in o])
        return dict(zip(("AF", "MAF", "Score", "Score_se", "pvalue_log", "Est", "Est_se"), o))

    def chunk_submit(self, G, impute: int = IMPUTE_MEAN, threads: int = 0, slot: int = 0):
        """Pipelined form of individual_analysis_chunk: pack + enqueue, returns at once.  G (F-order, float64 / int32 /
        uint8) must stay alive and unchanged until chunk_wait(slot)."""
        G = np.asarray(G)
        if not G.flags.f_contiguous or G.dtype not in (np.float64, np.int32, np.uint8):
            raise StaarError(1, "chunk_submit needs an F-ordered float64 / int32 / uint8 block (no hidden copy may be made)")
        fn = {np.dtype(np.float64): "staar_chunk_submit", np.dtype(np.int32): "staar_chunk_submit_i32",
              np.dtype(np.uint8): "staar_chunk_submit_u8"}[G.dtype]
        n, p = G.shape
        self._call(fn, G.ctypes.data_as(C.c_void_p), C.c_int64(n), C.c_int64(p), C.c_int(impute), C.c_int(threads), C.c_int(slot))
        self.__dict__.setdefault("_inflight", {})[slot] = (G, p)

    def chunk_wait(self, slot: int = 0, out=None):
        if slot not in self.__dict__.get("_inflight", {}):
            raise StaarError(3, f"chunk_wait: slot {slot} holds no chunk")
        G, p = self._inflight.pop(slot)
        o = out if out is not None else [np.zeros(p) for _ in range(7)]
        self._call("staar_chunk_wait", C.c_int(slot), *[_p(a) for a in o])
        return dict(zip(("AF", "MAF", "Score", "Score_se", "pvalue_log", "Est", "Est_se"), o))

    def packed_debug_dump(self, p: int):
        """intermediates of the last packed score call (test hook)"""
        hi, lo = np.zeros((p, 16), dtype=np.int64), np.zeros((p, 16), dtype=np.int64)
        Sm, quad, mu = np.zeros((p, 16)), np.zeros(p), np.zeros(p)
        flip, scale = np.zeros(p, dtype=np.int32), np.zeros(16)
        i64 = C.POINTER(C.c_longlong)
        self._call("staar_packed_debug_dump", C.c_int64(p), hi.ctypes.data_as(i64), lo.ctypes.data_as(i64), _p(Sm),
                   _p(quad), _p(mu), flip.ctypes.data_as(C.POINTER(C.c_int32)), _p(scale))
        return dict(hi=hi, lo=lo, Sm=Sm, quad=quad, mu=mu, flip=flip, col_scale=scale)

    def set_packed_path(self, path: str):
        """'fused' (default: one pass over G) or 'split' (contraction + block-per-variant scan); staar_b200.h"""
        self._call("staar_ctx_set_packed_path", C.c_int({"fused": 0, "split": 1}[path]))

    @property
    def last_packed_path_was_fused(self) -> bool:
        return bool(self._lib.staar_ctx_last_packed_path_was_fused(self._h))

    def packed_debug_fused(self, p: int):
        """scanner outputs of the last fused packed score call (test hook)"""
        acc = np.zeros(p)
        cnt, cm, guess = (np.zeros(p, dtype=np.int32) for _ in range(3))
        redo = C.c_uint64(0)
        i32 = C.POINTER(C.c_int32)
        self._call("staar_packed_debug_fused", C.c_int64(p), _p(acc), cnt.ctypes.data_as(i32), cm.ctypes.data_as(i32),
                   guess.ctypes.data_as(i32), C.byref(redo))
        return dict(acc=acc, cnt=cnt, cm=cm, guess=guess, redo_count=int(redo.value))

    def packed_score_device(self, d_packed: int, stride: int, n: int, p: int, impute: int, d_out7):
        """device pointers (ints): packed columns and 7 output arrays; asynchronous on the context stream."""
        self._call("staar_packed_score_device", C.c_void_p(d_packed), C.c_size_t(stride), C.c_int64(n), C.c_int64(p),
                   C.c_int(impute), *[C.c_void_p(x) for x in d_out7])

    def packed_analysis_device(self, d_packed: int, stride: int, n: int, p: int, impute: int, mac_cutoff: float,
                               spa_p_cutoff: float, d_index: int, d_out7, d_spa_index: int = 0):
        """resident columns -> compacted table of the tested variants (device pointers); returns (n_tested, n_spa)"""
        nt, ns = C.c_int64(0), C.c_int64(0)
        self._call("staar_packed_analysis_device", C.c_void_p(d_packed), C.c_size_t(stride), C.c_int64(n), C.c_int64(p),
                   C.c_int(impute), C.c_double(mac_cutoff), C.c_double(spa_p_cutoff), C.c_void_p(d_index),
                   *[C.c_void_p(x) for x in d_out7], C.c_void_p(d_spa_index or None), C.byref(nt), C.byref(ns))
        return nt.value, ns.value

    def gpu_pack_share(self) -> float:
        self._lib.staar_ctx_gpu_pack_share.restype = C.c_double
        return float(self._lib.staar_ctx_gpu_pack_share(self._h))

    def set_nullmodel_dense(self, P, residuals):
        Pm = _f(P)
        n = Pm.shape[0]
        if Pm.shape != (n, n):
            raise StaarError(1, "P must be square")
        r = _f(residuals).ravel()
        if r.size != n:
            raise StaarError(1, "matrix multiplication: incompatible matrix dimensions (residuals)")
        self._call("staar_nullmodel_set_dense", C.c_int64(n), _p(Pm), _p(r))

    def packed_score_dense_device(self, d_packed: int, stride: int, n: int, p: int, impute: int, d_out7):
        self._call("staar_packed_score_dense_device", C.c_void_p(d_packed), C.c_size_t(stride), C.c_int64(n), C.c_int64(p),
                   C.c_int(impute), *[C.c_void_p(x) for x in d_out7])

    def packed_score_multi_device(self, d_packed: int, stride: int, n: int, p: int, n_pheno: int, impute: int, d_AF: int,
                                  d_MAF: int, d_Score: int, d_pl: int):
        self._call("staar_packed_score_multi_device", C.c_void_p(d_packed), C.c_size_t(stride), C.c_int64(n), C.c_int64(p),
                   C.c_int(n_pheno), C.c_int(impute), C.c_void_p(d_AF), C.c_void_p(d_MAF), C.c_void_p(d_Score), C.c_void_p(d_pl))

    def packed_cond(self, d_packed: int, stride: int, n: int, d_variant: int, n_sel: int, impute: int, X_adj):
        A = _f(X_adj)
        if A.ndim != 2 or A.shape[0] != n:
            raise StaarError(1, "matrix multiplication: incompatible matrix dimensions (X_adj)")
        o, po = self._out5(n_sel)
        self._call("staar_packed_cond", C.c_void_p(d_packed), C.c_size_t(stride), C.c_int64(n), C.c_void_p(d_variant),
                   C.c_int64(n_sel), C.c_int(impute), _p(A), C.c_int64(A.shape[1]), *po)
        return self._pack5(o)

    def packed_cond_refit(self, d_packed: int, stride: int, n: int, d_tested: int, n_tested: int, d_known: int, n_known: int,
                          impute: int, X=None, method: str = "optimal"):
        """conditional test of a group of resident variants sharing their known loci, least-squares refit on the device"""
        qx = 0
        Xp = None
        if method == "optimal":
            Xf = _f(X); qx = Xf.shape[1]; Xp = _p(Xf)
        o, po = self._out5(n_tested)
        m = n_known + qx if method == "optimal" else 1 + n_known
        coef = np.zeros(m)
        self._call("staar_packed_cond_refit", C.c_void_p(d_packed), C.c_size_t(stride), C.c_int64(n), C.c_void_p(d_tested),
                   C.c_int64(n_tested), C.c_void_p(d_known), C.c_int64(n_known), Xp, C.c_int64(qx),
                   C.c_int(0 if method == "optimal" else 1), C.c_int(impute), *po, _p(coef))
        r = self._pack5(o)
        r["coef"] = coef
        return r

    def set_nullmodel_spa(self, XW, XXWX_inv, residuals, muhat):
        XWm, XXi = _f(XW), _f(XXWX_inv)
        q, n = XWm.shape
        if XXi.shape != (n, q):
            raise StaarError(1, "matrix multiplication: incompatible matrix dimensions (XW / XXWX_inv)")
        r, mu = _f(residuals).ravel(), _f(muhat).ravel()
        self._call("staar_nullmodel_set_spa", C.c_int64(n), C.c_int64(q), _p(XWm), _p(XXi), _p(r), _p(mu))

    def packed_spa_device(self, d_packed: int, stride: int, n: int, d_variant: int, n_sel: int, impute: int, tol: float,
                          max_iter: int, d_pvalue: int, d_trace: int = 0):
        self._call("staar_packed_spa_device", C.c_void_p(d_packed), C.c_size_t(stride), C.c_int64(n), C.c_void_p(d_variant),
                   C.c_int64(n_sel), C.c_int(impute), C.c_double(tol), C.c_int(max_iter), C.c_void_p(d_pvalue),
                   C.c_void_p(d_trace or None))

    def packed_flip_stats_device(self, d_packed: int, stride: int, n: int, p: int, impute: int, d_AF: int, d_MAF: int,
                                 d_flip: int, d_nmiss: int):
        self._call("staar_packed_flip_stats_device", C.c_void_p(d_packed), C.c_size_t(stride), C.c_int64(n),
                   C.c_int64(p), C.c_int(impute), C.c_void_p(d_AF), C.c_void_p(d_MAF), C.c_void_p(d_flip),
                   C.c_void_p(d_nmiss))

    def synth_packed_device(self, d_packed: int, stride: int, n: int, first_variant: int, p: int, seed: int,
                            d_thr_hom: int, d_thr_het: int, d_thr_miss: int, d_store_alt: int, d_sample_order: int = 0):
        self._call("staar_synth_packed_device", C.c_void_p(d_packed), C.c_size_t(stride), C.c_int64(n),
                   C.c_int64(first_variant), C.c_int64(p), C.c_uint64(seed), C.c_void_p(d_thr_hom),
                   C.c_void_p(d_thr_het), C.c_void_p(d_thr_miss), C.c_void_p(d_store_alt),
                   C.c_void_p(d_sample_order or None))

    def nullmodel_sample_order(self) -> np.ndarray:
        """position -> caller's sample index of RESIDENT packed columns (identity for unrelated samples)"""
        order = np.zeros(self._n, dtype=np.int32)
        self._call("staar_nullmodel_sample_order", C.c_int64(self._n), order.ctypes.data_as(C.POINTER(C.c_int32)))
        return order

    def packed_reorder_device(self, d_in: int, d_out: int, stride: int, n: int, p: int):
        self._call("staar_packed_reorder_device", C.c_void_p(d_in), C.c_void_p(d_out), C.c_size_t(stride),
                   C.c_int64(n), C.c_int64(p))


class MultiContext:
    """several GPUs of one box behind one handle (include/staar_b200.h section C); mirrors the Context calls it covers"""

    def __init__(self, devices):
        self._lib = load_library()
        self._h = C.c_void_p()
        dev = (C.c_int * len(devices))(*devices)
        rc = self._lib.staar_multi_create(dev, C.c_int(len(devices)), C.byref(self._h))
        if rc != 0:
            raise StaarError(rc, self._lib.staar_last_error().decode())
        self.devices = list(devices)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.staar_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def _call(self, name, *args):
        rc = getattr(self._lib, name)(self._h, *args)
        if rc != 0:
            raise StaarError(rc, self._lib.staar_last_error().decode())

    def set_nullmodel_sparse(self, Sigma_i, Sigma_iX, cov, residuals):
        n = sp.csc_matrix(Sigma_i).shape[0]
        keep, sargs = Context._sigma_args(n, Sigma_i, Sigma_iX, cov, residuals)
        S, SX, cv, r = keep
        self._call("staar_multi_nullmodel_set_sparse", C.c_int64(n), C.c_int64(SX.shape[1]), *S.args(), _p(SX), _p(cv), _p(r))

    def individual_analysis_chunk(self, G, impute: int = IMPUTE_MEAN, threads: int = 0):
        G = _f(G)
        n, p = G.shape
        o = [np.zeros(p) for _ in range(7)]
        self._call("staar_multi_individual_analysis_chunk", _p(G), C.c_int64(n), C.c_int64(p), C.c_int(impute), C.c_int(threads),
                   *[_p(a) for a in o])
        return dict(zip(("AF", "MAF", "Score", "Score_se", "pvalue_log", "Est", "Est_se"), o))

    def matrix_flip_mean(self, G):
        G = _f(G)
        n, p = G.shape
        out, af, maf = np.empty((n, p), order="F"), np.zeros(p), np.zeros(p)
        self._call("staar_multi_matrix_flip_mean", _p(G), C.c_int64(n), C.c_int64(p), _p(out), _p(af), _p(maf))
        return {"Geno": out, "AF": af, "MAF": maf}

    def Individual_Score_Test(self, G, Sigma_i, Sigma_iX, cov, residuals):
        G = _f(G)
        n, p = G.shape
        keep, sargs = Context._sigma_args(n, Sigma_i, Sigma_iX, cov, residuals)
        o, po = Context._out5(p)
        self._call("staar_multi_individual_score_test", _p(G), C.c_int64(n), C.c_int64(p), *sargs, *po)
        return Context._pack5(o)


def packed_stride(n: int) -> int:
    return int(load_library().staar_packed_stride(C.c_int64(n)))


def pack_dosage_host(G, threads: int = 0) -> np.ndarray:
    """`$dosage` matrix (n x p; float64 with NaN / <= -1 missing, int32 with negative missing, or uint8 with 0xFF
    missing) -> (p, stride) uint8 packed columns (host, threaded, AVX-512 where available)."""
    lib = load_library()
    G = np.asarray(G)
    if G.dtype == np.int32 or G.dtype == np.uint8:
        G = np.asfortranarray(G)
        fn = lib.staar_pack_dosage_host_i32 if G.dtype == np.int32 else lib.staar_pack_dosage_host_u8
        gp = G.ctypes.data_as(C.c_void_p)
    else:
        G = _f(G)
        fn, gp = lib.staar_pack_dosage_host, _p(G)
    n, p = G.shape
    stride = packed_stride(n)
    out = np.full((p, stride), 0xCC, dtype=np.uint8)        # poison: the packer must write every byte of the stride
    rc = fn(gp, C.c_int64(n), C.c_int64(p), out.ctypes.data_as(_u8p), C.c_size_t(stride), C.c_int(threads))
    if rc != 0:
        raise StaarError(rc, lib.staar_last_error().decode())
    return out
