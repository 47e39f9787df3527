/*
 * ref_stub/RcppArmadillo.h — TEST INFRASTRUCTURE (oracle), NOT product code.
 *
 * A minimal stand-in for the slice of the RcppArmadillo / Rcpp / Rmath API
 * that the 13 hot-path sources of the reference use, so that those sources
 * can be compiled UNMODIFIED, where they lie under /root/reference/src, into
 * oracle/_ref/libstaar_ref.so (recipe: oracle/Makefile target `ref`).
 *
 * What this pins: the reference's own formulas, association order and
 * control flow (e.g. the 5-term conditional covariance, every SPA branch).
 * What it does NOT pin: Armadillo's / BLAS's summation order and R's nmath —
 * those are stand-ins written here (eager evaluation, in-order loops or an
 * optional cblas_dgemm hook; Rmath functions come from liboracle.so).
 * Nothing in this header is copied from Armadillo or Rcpp.
 */
#ifndef REF_STUB_RCPPARMADILLO_H
#define REF_STUB_RCPPARMADILLO_H

#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include <algorithm>

extern "C" double orc_pchisq_upper_log(double x, double df);
extern "C" double orc_pnorm(double x, int lower);

namespace arma {

typedef uint64_t uword;   /* ARMA_64BIT_WORD=1 (reference src/Makevars:2) */

/* optional BLAS hook (set by ref_glue.cpp when an OpenBLAS/MKL cblas is found):
 * C(m x n) = A(m x k, lda) * B(k x n, ldb), column-major, transA/transB flags */
typedef void (*dgemm_hook_t)(int transA, int transB, int64_t m, int64_t n, int64_t k,
                             const double *A, int64_t lda, const double *B, int64_t ldb,
                             double *C, int64_t ldc);
inline dgemm_hook_t &dgemm_hook() { static dgemm_hook_t h = nullptr; return h; }

struct uvec {
    std::vector<uword> mem;
    void zeros(uword n) { mem.assign(n, 0); }
    uword &operator()(uword i) { return mem[i]; }
    uword operator()(uword i) const { return mem[i]; }
    uword size() const { return mem.size(); }
};

struct vec;

struct mat {
    uword n_rows = 0, n_cols = 0;
    std::vector<double> mem;
    mat() {}
    mat(uword r, uword c) : n_rows(r), n_cols(c), mem(r * c, 0.0) {}
    void zeros(uword r, uword c) { n_rows = r; n_cols = c; mem.assign(r * c, 0.0); }
    double &operator()(uword i, uword j) { return mem[i + j * n_rows]; }
    double operator()(uword i, uword j) const { return mem[i + j * n_rows]; }
    double *memptr() { return mem.data(); }
    const double *memptr() const { return mem.data(); }
    uword size() const { return mem.size(); }
    inline vec col(uword j) const;
    mat operator()(const uvec &r, const uvec &c) const {
        mat out(r.size(), c.size());
        for (uword j = 0; j < c.size(); j++)
            for (uword i = 0; i < r.size(); i++) out(i, j) = (*this)(r(i), c(j));
        return out;
    }
};

struct vec : mat {
    vec() {}
    vec(const mat &m) : mat(m) {
        if (m.n_cols != 1 && m.n_rows == 1) { n_rows = m.n_cols; n_cols = 1; }
    }
    void zeros(uword n) { mat::zeros(n, 1); }
    void ones(uword n) { n_rows = n; n_cols = 1; mem.assign(n, 1.0); }
    double &operator()(uword i) { return mem[i]; }
    double operator()(uword i) const { return mem[i]; }
};

struct rowvec : mat {
    rowvec() {}
    rowvec(const mat &m) : mat(m) {
        if (m.n_rows != 1 && m.n_cols == 1) { n_cols = m.n_rows; n_rows = 1; }
    }
    double &operator()(uword i) { return mem[i]; }
    double operator()(uword i) const { return mem[i]; }
    /* Armadillo: X(uvec) is a subview_elem1, which behaves as a COLUMN vector */
    vec operator()(const uvec &id) const {
        vec out; out.zeros(id.size());
        for (uword i = 0; i < id.size(); i++) out(i) = mem[id(i)];
        return out;
    }
};

inline vec mat::col(uword j) const {
    vec out; out.zeros(n_rows);
    std::memcpy(out.mem.data(), mem.data() + j * n_rows, sizeof(double) * n_rows);
    return out;
}

/* CSC sparse matrix */
struct sp_mat {
    uword n_rows = 0, n_cols = 0;
    std::vector<double> values;
    std::vector<uword> row_indices;
    std::vector<uword> col_ptrs;
    sp_mat() {}
    sp_mat(uword r, uword c) : n_rows(r), n_cols(c), col_ptrs(c + 1, 0) {}
};

/* ---- transposes ---- */
inline mat trans(const mat &A) {
    mat T(A.n_cols, A.n_rows);
    for (uword j = 0; j < A.n_cols; j++)
        for (uword i = 0; i < A.n_rows; i++) T(j, i) = A(i, j);
    return T;
}
inline sp_mat trans(const sp_mat &A) {
    sp_mat T(A.n_cols, A.n_rows);
    uword nnz = A.values.size();
    T.values.resize(nnz); T.row_indices.resize(nnz);
    std::vector<uword> cnt(A.n_rows + 1, 0);
    for (uword e = 0; e < nnz; e++) cnt[A.row_indices[e] + 1]++;
    for (uword r = 0; r < A.n_rows; r++) cnt[r + 1] += cnt[r];
    T.col_ptrs = cnt;
    std::vector<uword> pos(cnt.begin(), cnt.end() - 1);
    for (uword c = 0; c < A.n_cols; c++)
        for (uword e = A.col_ptrs[c]; e < A.col_ptrs[c + 1]; e++) {
            uword d = pos[A.row_indices[e]]++;
            T.values[d] = A.values[e]; T.row_indices[d] = c;
        }
    return T;
}

inline void check_mul(uword a_cols, uword b_rows) {
    if (a_cols != b_rows) throw std::logic_error("matrix multiplication: incompatible matrix dimensions");
}

/* ---- dense * dense ---- */
inline mat operator*(const mat &A, const mat &B) {
    check_mul(A.n_cols, B.n_rows);
    mat C(A.n_rows, B.n_cols);
    if (dgemm_hook() && A.n_rows * B.n_cols * A.n_cols > 4096) {
        dgemm_hook()(0, 0, A.n_rows, B.n_cols, A.n_cols, A.memptr(), A.n_rows, B.memptr(), B.n_rows,
                     C.memptr(), C.n_rows);
        return C;
    }
    for (uword j = 0; j < B.n_cols; j++)
        for (uword k = 0; k < A.n_cols; k++) {
            double b = B(k, j);
            if (b == 0.0) continue;
            const double *a = A.memptr() + k * A.n_rows;
            double *c = C.memptr() + j * C.n_rows;
            for (uword i = 0; i < A.n_rows; i++) c[i] += a[i] * b;
        }
    return C;
}
/* ---- dense * sparse -> dense ---- */
inline mat operator*(const mat &A, const sp_mat &B) {
    check_mul(A.n_cols, B.n_rows);
    mat C(A.n_rows, B.n_cols);
    for (uword j = 0; j < B.n_cols; j++) {
        double *c = C.memptr() + j * C.n_rows;
        for (uword e = B.col_ptrs[j]; e < B.col_ptrs[j + 1]; e++) {
            double b = B.values[e];
            const double *a = A.memptr() + B.row_indices[e] * A.n_rows;
            for (uword i = 0; i < A.n_rows; i++) c[i] += a[i] * b;
        }
    }
    return C;
}
/* ---- sparse * dense -> dense ---- */
inline mat operator*(const sp_mat &A, const mat &B) {
    check_mul(A.n_cols, B.n_rows);
    mat C(A.n_rows, B.n_cols);
    for (uword j = 0; j < B.n_cols; j++) {
        double *c = C.memptr() + j * C.n_rows;
        for (uword k = 0; k < A.n_cols; k++) {
            double b = B(k, j);
            if (b == 0.0) continue;
            for (uword e = A.col_ptrs[k]; e < A.col_ptrs[k + 1]; e++) c[A.row_indices[e]] += A.values[e] * b;
        }
    }
    return C;
}
/* ---- sparse * sparse -> sparse (Gustavson, sorted rows) ---- */
inline sp_mat operator*(const sp_mat &A, const sp_mat &B) {
    check_mul(A.n_cols, B.n_rows);
    sp_mat C(A.n_rows, B.n_cols);
    std::vector<double> acc(A.n_rows, 0.0);
    std::vector<char> seen(A.n_rows, 0);
    std::vector<uword> touched;
    for (uword j = 0; j < B.n_cols; j++) {
        touched.clear();
        for (uword e = B.col_ptrs[j]; e < B.col_ptrs[j + 1]; e++) {
            uword k = B.row_indices[e]; double b = B.values[e];
            for (uword f = A.col_ptrs[k]; f < A.col_ptrs[k + 1]; f++) {
                uword r = A.row_indices[f];
                if (!seen[r]) { seen[r] = 1; touched.push_back(r); }
                acc[r] += A.values[f] * b;
            }
        }
        std::sort(touched.begin(), touched.end());
        for (uword r : touched) {
            if (acc[r] != 0.0) { C.values.push_back(acc[r]); C.row_indices.push_back(r); }
            acc[r] = 0.0; seen[r] = 0;
        }
        C.col_ptrs[j + 1] = C.values.size();
    }
    return C;
}

inline mat operator-(const mat &A, const mat &B) {
    if (A.n_rows != B.n_rows || A.n_cols != B.n_cols) throw std::logic_error("subtraction: incompatible matrix dimensions");
    mat C(A.n_rows, A.n_cols);
    for (uword i = 0; i < A.mem.size(); i++) C.mem[i] = A.mem[i] - B.mem[i];
    return C;
}
inline mat operator+(const mat &A, const mat &B) {
    if (A.n_rows != B.n_rows || A.n_cols != B.n_cols) throw std::logic_error("addition: incompatible matrix dimensions");
    mat C(A.n_rows, A.n_cols);
    for (uword i = 0; i < A.mem.size(); i++) C.mem[i] = A.mem[i] + B.mem[i];
    return C;
}
inline mat operator-(const sp_mat &A, const mat &B) {
    if (A.n_rows != B.n_rows || A.n_cols != B.n_cols) throw std::logic_error("subtraction: incompatible matrix dimensions");
    mat C(B.n_rows, B.n_cols);
    for (uword i = 0; i < B.mem.size(); i++) C.mem[i] = -B.mem[i];
    for (uword j = 0; j < A.n_cols; j++)
        for (uword e = A.col_ptrs[j]; e < A.col_ptrs[j + 1]; e++)
            C(A.row_indices[e], j) = A.values[e] - B(A.row_indices[e], j);
    return C;
}

/* ---- LU (partial pivoting) det / inv, standing in for LAPACK getrf/getri ---- */
inline int lu_inplace(mat &A, std::vector<uword> &piv) {
    uword k = A.n_rows; int sign = 1; piv.resize(k);
    for (uword c = 0; c < k; c++) {
        uword pr = c; double best = std::fabs(A(c, c));
        for (uword r = c + 1; r < k; r++) if (std::fabs(A(r, c)) > best) { best = std::fabs(A(r, c)); pr = r; }
        piv[c] = pr;
        if (pr != c) { for (uword cc = 0; cc < k; cc++) std::swap(A(c, cc), A(pr, cc)); sign = -sign; }
        double d = A(c, c);
        if (d == 0.0) return 0;
        for (uword r = c + 1; r < k; r++) {
            double f = A(r, c) / d; A(r, c) = f;
            for (uword cc = c + 1; cc < k; cc++) A(r, cc) -= f * A(c, cc);
        }
    }
    return sign;
}
inline double det(const mat &A0) {
    if (A0.n_rows != A0.n_cols) throw std::logic_error("det(): given matrix must be square sized");
    mat A = A0; std::vector<uword> piv;
    int sign = lu_inplace(A, piv);
    double d = (double)sign;
    if (sign != 0) for (uword c = 0; c < A.n_rows; c++) d *= A(c, c);
    return d;
}
inline mat inv(const mat &A0) {
    if (A0.n_rows != A0.n_cols) throw std::logic_error("inv(): given matrix must be square sized");
    mat A = A0; std::vector<uword> piv; uword k = A.n_rows;
    if (lu_inplace(A, piv) == 0) throw std::runtime_error("inv(): matrix is singular");
    mat X(k, k);
    for (uword c = 0; c < k; c++) {
        double *x = X.memptr() + c * k;
        x[c] = 1.0;
        for (uword r = 0; r < k; r++) if (piv[r] != r) std::swap(x[r], x[piv[r]]);
        for (uword r = 0; r < k; r++) for (uword cc = 0; cc < r; cc++) x[r] -= A(r, cc) * x[cc];
        for (uword r = k; r-- > 0;) {
            for (uword cc = r + 1; cc < k; cc++) x[r] -= A(r, cc) * x[cc];
            x[r] /= A(r, r);
        }
    }
    return X;
}

} /* namespace arma */

/* ---- Rcpp / R API slice ---- */
#define REALSXP 14
inline int R_finite(double x) { return std::isfinite(x) ? 1 : 0; }

namespace R {
inline double pchisq(double x, double df, int lower, int logp) {
    if (lower || !logp) throw std::logic_error("ref_stub: only pchisq(lower=false, log=true) is provided");
    return orc_pchisq_upper_log(x, df);
}
inline double pnorm(double x, double mu, double sd, int lower, int logp) {
    if (logp) throw std::logic_error("ref_stub: only pnorm(log=false) is provided");
    return orc_pnorm((x - mu) / sd, lower);
}
}

namespace Rcpp {
namespace traits {
template <int RTYPE> inline bool is_na(double x) { return std::isnan(x); }
}
struct NamedValue { std::string name; arma::mat value; };
struct Named {
    std::string name;
    explicit Named(const char *n) : name(n) {}
    NamedValue operator=(const arma::mat &v) const { return NamedValue{name, v}; }
};
struct List {
    std::vector<NamedValue> items;
    template <typename... Args> static List create(Args &&...args) {
        List l; (l.items.push_back(std::forward<Args>(args)), ...); return l;
    }
    const arma::mat &operator[](const char *n) const {
        for (auto &it : items) if (it.name == n) return it.value;
        throw std::out_of_range(std::string("List: no element named ") + n);
    }
};
}

#endif
