"""ctypes front-end to the CPU oracle — TEST INFRASTRUCTURE, NOT product code.

Loads ``oracle/liboracle.so`` (the plain-C restatement, ``staar_oracle.c``) and,
when it has been built, ``oracle/_ref/libstaar_ref.so`` (the reference's own
``src/*.cpp`` compiled against the stand-in headers in ``oracle/ref_stub/``).
Both expose the same flat signatures, so ``Oracle("port")`` and
``Oracle("reference")`` are interchangeable in the tests.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / reference
arm may import this module.  Function names mirror the reference's exported R
functions (/root/reference/R/RcppExports.R:4-54); inputs are numpy arrays
(column-major semantics: an ``(n, p)`` array means the n x p matrix) and
``scipy.sparse`` matrices.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_u64p = C.POINTER(C.c_uint64)


class _Csc(C.Structure):
    _fields_ = [("n_rows", C.c_int64), ("n_cols", C.c_int64), ("values", _dp),
                ("row_idx", _u64p), ("col_ptr", _u64p)]


def build(with_ref: bool = True) -> None:
    """Compile liboracle.so (always) and _ref/libstaar_ref.so (only where /root/reference exists)."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    if with_ref and os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=True)


def have_reference() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libstaar_ref.so"))


def _f(a) -> np.ndarray:
    """FP64, Fortran(column-major)-contiguous copy/view."""
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_dp)


class _CscHolder:
    def __init__(self, m):
        m = sp.csc_matrix(m, dtype=np.float64)
        m.sort_indices()
        self.values = np.ascontiguousarray(m.data, dtype=np.float64)
        self.row_idx = np.ascontiguousarray(m.indices, dtype=np.uint64)
        self.col_ptr = np.ascontiguousarray(m.indptr, dtype=np.uint64)
        self.shape = m.shape
        self.c = _Csc(m.shape[0], m.shape[1], _ptr(self.values),
                      self.row_idx.ctypes.data_as(_u64p), self.col_ptr.ctypes.data_as(_u64p))

    def ref(self):
        return C.byref(self.c)


class Oracle:
    """kind = "port" (C restatement) or "reference" (reference sources + stand-in headers)."""

    def __init__(self, kind: str = "port"):
        self.kind = kind
        port = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(port):
            build(with_ref=False)
        self._port = C.CDLL(port, mode=C.RTLD_GLOBAL)
        self._port.orc_pchisq_upper_log.restype = C.c_double
        self._port.orc_pchisq_upper_log.argtypes = [C.c_double, C.c_double]
        self._port.orc_pnorm.restype = C.c_double
        self._port.orc_pnorm.argtypes = [C.c_double, C.c_int]
        self._port.orc_det.restype = C.c_double
        if kind == "port":
            self._lib, self._pfx = self._port, "orc_"
        elif kind == "reference":
            path = os.path.join(_HERE, "_ref", "libstaar_ref.so")
            if not os.path.exists(path):
                raise FileNotFoundError(path + " (build with `make -C oracle ref` where /root/reference exists)")
            self._lib, self._pfx = C.CDLL(path), "ref_"
            self._lib.ref_last_error.restype = C.c_char_p
        else:
            raise ValueError(kind)

    # -- misc ------------------------------------------------------------
    def set_blas(self, path: str | None = None) -> bool:
        """reference kind only: route the stand-in's dense GEMM to a real multi-threaded BLAS."""
        if self.kind != "reference":
            return False
        if path is None:
            import numpy
            cands = glob.glob(os.path.join(os.path.dirname(numpy.__file__), "..", "numpy.libs", "libscipy_openblas*.so"))
            if not cands:
                return False
            path = cands[0]
        return self._lib.ref_set_blas(path.encode()) == 0

    def _call(self, name, *args):
        rc = getattr(self._lib, self._pfx + name)(*args)
        if rc != 0:
            msg = self._lib.ref_last_error().decode() if self.kind == "reference" else "singular matrix in inv()"
            raise RuntimeError(f"{name}: {msg}")

    def pchisq_upper_log(self, x, df):
        return self._port.orc_pchisq_upper_log(float(x), float(df))

    def pnorm(self, x, lower=True):
        return self._port.orc_pnorm(float(x), int(bool(lower)))

    # -- flips -------------------------------------------------------------
    def _flip(self, G, minor):
        G = _f(G).copy(order="F")
        n, p = G.shape
        AF = np.zeros(p); MAF = np.zeros(p)
        if self.kind == "port":
            fn = self._lib.orc_matrix_flip_minor if minor else self._lib.orc_matrix_flip_mean
            fn(_ptr(G), C.c_int64(n), C.c_int64(p), _ptr(AF), _ptr(MAF))
        else:
            self._call("matrix_flip", C.c_int(int(minor)), _ptr(G), C.c_int64(n), C.c_int64(p), _ptr(AF), _ptr(MAF))
        return {"Geno": G, "AF": AF, "MAF": MAF}

    def matrix_flip_mean(self, G):
        return self._flip(G, False)

    def matrix_flip_minor(self, G):
        return self._flip(G, True)

    # -- single-trait ---------------------------------------------------------
    @staticmethod
    def _out5(p):
        return [np.zeros(p) for _ in range(5)]

    @staticmethod
    def _pack5(o):
        return dict(zip(("Score", "Score_se", "pvalue_log", "Est", "Est_se"), o))

    def Individual_Score_Test(self, G, Sigma_i, Sigma_iX, cov, residuals):
        G = _f(G); n, p = G.shape
        S = _CscHolder(Sigma_i); SX = _f(Sigma_iX).reshape(n, -1, order="F"); q = SX.shape[1]
        cv = _f(cov); r = _f(residuals).ravel(); o = self._out5(p)
        self._call("score_test", _ptr(G), C.c_int64(n), C.c_int64(p), S.ref(), _ptr(SX), C.c_int64(q),
                   _ptr(cv), _ptr(r), *map(_ptr, o))
        return self._pack5(o)

    def Individual_Score_Test_sp(self, G, Sigma_i, Sigma_iX, cov, residuals):
        Gs = _CscHolder(G); n, p = Gs.shape
        S = _CscHolder(Sigma_i); SX = _f(Sigma_iX).reshape(n, -1, order="F"); q = SX.shape[1]
        cv = _f(cov); r = _f(residuals).ravel(); o = self._out5(p)
        self._call("score_test_sp", Gs.ref(), S.ref(), _ptr(SX), C.c_int64(q), _ptr(cv), _ptr(r), *map(_ptr, o))
        return self._pack5(o)

    def Individual_Score_Test_denseGRM(self, G, P, residuals):
        G = _f(G); n, p = G.shape
        Pm = _f(P); r = _f(residuals).ravel(); o = self._out5(p)
        self._call("score_test_denseGRM", _ptr(G), C.c_int64(n), C.c_int64(p), _ptr(Pm), _ptr(r), *map(_ptr, o))
        return self._pack5(o)

    def Individual_Score_Test_sp_denseGRM(self, G, P, residuals):
        Gs = _CscHolder(G); n, p = Gs.shape
        Pm = _f(P); r = _f(residuals).ravel(); o = self._out5(p)
        self._call("score_test_sp_denseGRM", Gs.ref(), _ptr(Pm), _ptr(r), *map(_ptr, o))
        return self._pack5(o)

    # -- multi-trait -------------------------------------------------------
    def Individual_Score_Test_multi(self, G, Sigma_i, Sigma_iX, cov, residuals, n_pheno=1):
        G = _f(G); n, p = G.shape
        S = _CscHolder(Sigma_i); SX = _f(Sigma_iX).reshape(n, -1, order="F"); q = SX.shape[1]
        cv = _f(cov); r = _f(residuals).ravel()
        Score = np.zeros(p); pl = np.zeros(p // n_pheno)
        self._call("score_test_multi", _ptr(G), C.c_int64(n), C.c_int64(p), S.ref(), _ptr(SX), C.c_int64(q),
                   _ptr(cv), _ptr(r), C.c_int(n_pheno), _ptr(Score), _ptr(pl))
        return {"Score": Score, "pvalue_log": pl}

    def Individual_Score_Test_sp_multi(self, G, Sigma_i, Sigma_iX, cov, residuals, n_pheno=1):
        Gs = _CscHolder(G); n, p = Gs.shape
        S = _CscHolder(Sigma_i); SX = _f(Sigma_iX).reshape(n, -1, order="F"); q = SX.shape[1]
        cv = _f(cov); r = _f(residuals).ravel()
        Score = np.zeros(p); pl = np.zeros(p // n_pheno)
        self._call("score_test_sp_multi", Gs.ref(), S.ref(), _ptr(SX), C.c_int64(q), _ptr(cv), _ptr(r),
                   C.c_int(n_pheno), _ptr(Score), _ptr(pl))
        return {"Score": Score, "pvalue_log": pl}

    def Individual_Score_Test_denseGRM_multi(self, G, P, residuals, n_pheno=1):
        G = _f(G); n, p = G.shape
        Pm = _f(P); r = _f(residuals).ravel()
        Score = np.zeros(p); pl = np.zeros(p // n_pheno)
        self._call("score_test_denseGRM_multi", _ptr(G), C.c_int64(n), C.c_int64(p), _ptr(Pm), _ptr(r),
                   C.c_int(n_pheno), _ptr(Score), _ptr(pl))
        return {"Score": Score, "pvalue_log": pl}

    def Individual_Score_Test_sp_denseGRM_multi(self, G, P, residuals, n_pheno=1):
        Gs = _CscHolder(G); n, p = Gs.shape
        Pm = _f(P); r = _f(residuals).ravel()
        Score = np.zeros(p); pl = np.zeros(p // n_pheno)
        self._call("score_test_sp_denseGRM_multi", Gs.ref(), _ptr(Pm), _ptr(r), C.c_int(n_pheno), _ptr(Score), _ptr(pl))
        return {"Score": Score, "pvalue_log": pl}

    # -- conditional -----------------------------------------------------
    def Individual_Score_Test_cond(self, G, Sigma_i, Sigma_iX, cov, X_adj, residuals):
        G = _f(G); n, p = G.shape
        S = _CscHolder(Sigma_i); SX = _f(Sigma_iX).reshape(n, -1, order="F"); q = SX.shape[1]
        A = _f(X_adj).reshape(n, -1, order="F"); m = A.shape[1]
        cv = _f(cov); r = _f(residuals).ravel(); o = self._out5(p)
        self._call("score_test_cond", _ptr(G), C.c_int64(n), C.c_int64(p), S.ref(), _ptr(SX), C.c_int64(q),
                   _ptr(cv), _ptr(A), C.c_int64(m), _ptr(r), *map(_ptr, o))
        return self._pack5(o)

    def Individual_Score_Test_cond_multi(self, G, Sigma_i, Sigma_iX, cov, X_adj, residuals, n_pheno=1):
        G = _f(G); n, p = G.shape
        S = _CscHolder(Sigma_i); SX = _f(Sigma_iX).reshape(n, -1, order="F"); q = SX.shape[1]
        A = _f(X_adj).reshape(n, -1, order="F"); m = A.shape[1]
        cv = _f(cov); r = _f(residuals).ravel()
        Score = np.zeros(p); pl = np.zeros(p // n_pheno)
        self._call("score_test_cond_multi", _ptr(G), C.c_int64(n), C.c_int64(p), S.ref(), _ptr(SX), C.c_int64(q),
                   _ptr(cv), _ptr(A), C.c_int64(m), _ptr(r), C.c_int(n_pheno), _ptr(Score), _ptr(pl))
        return {"Score": Score, "pvalue_log": pl}

    # -- SPA -------------------------------------------------------------
    def Individual_Score_Test_SPA(self, G, XW, XXWX_inv, residuals, muhat, tol, max_iter, return_trace=False):
        G = _f(G); n, p = G.shape
        XWm = _f(XW); q = XWm.shape[0]; XXi = _f(XXWX_inv)
        r = _f(residuals).ravel(); mu = _f(muhat).ravel(); pv = np.zeros(p)
        if self.kind == "port":
            trace = np.zeros((p, 4), dtype=np.int64)
            self._call("score_test_SPA", _ptr(G), C.c_int64(n), C.c_int64(p), _ptr(XWm), _ptr(XXi), C.c_int64(q),
                       _ptr(r), _ptr(mu), C.c_double(tol), C.c_int(max_iter), _ptr(pv),
                       trace.ctypes.data_as(C.POINTER(C.c_int64)))
            return (pv, trace) if return_trace else pv
        self._call("score_test_SPA", _ptr(G), C.c_int64(n), C.c_int64(p), _ptr(XWm), _ptr(XXi), C.c_int64(q),
                   _ptr(r), _ptr(mu), C.c_double(tol), C.c_int(max_iter), _ptr(pv))
        return (pv, None) if return_trace else pv
