/*
 * ref_glue.cpp — TEST INFRASTRUCTURE (oracle), NOT product code.
 *
 * extern "C" entry points around the reference's OWN functions
 * (/root/reference/src/*.cpp, compiled unmodified against ref_stub/), with
 * the same flat signatures as the C oracle (staar_oracle.h) so that tests
 * can run oracle and reference side by side.  Built by `make -C oracle ref`
 * into oracle/_ref/libstaar_ref.so (git-ignored, travels with gpurun).
 */
#include <RcppArmadillo.h>
#include <dlfcn.h>
#include <cstdio>
#include "staar_oracle.h"

using Rcpp::List;

/* the reference's exported functions (src/RcppExports.cpp:18,33,50,66,83,96,110,126,141,154,168,184,218) */
List Individual_Score_Test(arma::mat G, arma::sp_mat Sigma_i, arma::mat Sigma_iX, arma::mat cov, arma::vec residuals);
arma::vec Individual_Score_Test_SPA(arma::mat G, arma::mat XW, arma::mat XXWX_inv, arma::vec residuals, arma::vec muhat, double tol, int max_iter);
List Individual_Score_Test_cond(arma::mat G, arma::sp_mat Sigma_i, arma::mat Sigma_iX, arma::mat cov, arma::mat X_adj, arma::vec residuals);
List Individual_Score_Test_cond_multi(arma::mat G, arma::sp_mat Sigma_i, arma::mat Sigma_iX, arma::mat cov, arma::mat X_adj, arma::vec residuals, int n_pheno);
List Individual_Score_Test_denseGRM(arma::mat G, arma::mat P, arma::vec residuals);
List Individual_Score_Test_denseGRM_multi(arma::mat G, arma::mat P, arma::vec residuals, int n_pheno);
List Individual_Score_Test_multi(arma::mat G, arma::sp_mat Sigma_i, arma::mat Sigma_iX, arma::mat cov, arma::vec residuals, int n_pheno);
List Individual_Score_Test_sp(arma::sp_mat G, arma::sp_mat Sigma_i, arma::mat Sigma_iX, arma::mat cov, arma::vec residuals);
List Individual_Score_Test_sp_denseGRM(arma::sp_mat G, arma::mat P, arma::vec residuals);
List Individual_Score_Test_sp_denseGRM_multi(arma::sp_mat G, arma::mat P, arma::vec residuals, int n_pheno);
List Individual_Score_Test_sp_multi(arma::sp_mat G, arma::sp_mat Sigma_i, arma::mat Sigma_iX, arma::mat cov, arma::vec residuals, int n_pheno);
List matrix_flip_mean(arma::mat G);
List matrix_flip_minor(arma::mat G);

namespace {

arma::mat to_mat(const double *p, int64_t r, int64_t c) {
    arma::mat m(r, c);
    if (r * c) std::memcpy(m.memptr(), p, sizeof(double) * r * c);
    return m;
}
arma::vec to_vec(const double *p, int64_t n) { return arma::vec(to_mat(p, n, 1)); }
arma::sp_mat to_sp(const orc_csc *s) {
    arma::sp_mat m(s->n_rows, s->n_cols);
    uint64_t nnz = s->col_ptr[s->n_cols];
    m.values.assign(s->values, s->values + nnz);
    m.row_indices.assign(s->row_idx, s->row_idx + nnz);
    m.col_ptrs.assign(s->col_ptr, s->col_ptr + s->n_cols + 1);
    return m;
}
void out(const List &l, const char *name, double *dst) {
    if (!dst) return;
    const arma::mat &m = l[name];
    std::memcpy(dst, m.memptr(), sizeof(double) * m.size());
}
void out5(const List &l, double *Score, double *Score_se, double *pvalue_log, double *Est, double *Est_se) {
    out(l, "Score", Score); out(l, "Score_se", Score_se); out(l, "pvalue_log", pvalue_log);
    out(l, "Est", Est); out(l, "Est_se", Est_se);
}
char g_err[512];
template <typename F> int guarded(F &&f) {
    try { f(); return 0; }
    catch (const std::exception &e) { std::snprintf(g_err, sizeof g_err, "%s", e.what()); return 1; }
}

/* cblas hook */
typedef void (*cblas_dgemm64_t)(int, int, int, int64_t, int64_t, int64_t, double, const double *, int64_t,
                                const double *, int64_t, double, double *, int64_t);
typedef void (*cblas_dgemm32_t)(int, int, int, int, int, int, double, const double *, int,
                                const double *, int, double, double *, int);
cblas_dgemm64_t g_dgemm64 = nullptr;
cblas_dgemm32_t g_dgemm32 = nullptr;
void hook(int tA, int tB, int64_t m, int64_t n, int64_t k, const double *A, int64_t lda,
          const double *B, int64_t ldb, double *C, int64_t ldc) {
    if (g_dgemm64) g_dgemm64(102, tA ? 112 : 111, tB ? 112 : 111, m, n, k, 1.0, A, lda, B, ldb, 0.0, C, ldc);
    else g_dgemm32(102, tA ? 112 : 111, tB ? 112 : 111, (int)m, (int)n, (int)k, 1.0, A, (int)lda, B, (int)ldb, 0.0, C, (int)ldc);
}

} // namespace

extern "C" {

const char *ref_last_error(void) { return g_err; }

/* Point the stand-in's dense GEMM at a real BLAS (the reference links R's BLAS, src/Makevars:1). */
int ref_set_blas(const char *path) {
    void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) { std::snprintf(g_err, sizeof g_err, "%s", dlerror()); return 1; }
    if (void *s = dlsym(h, "scipy_cblas_dgemm64_")) g_dgemm64 = (cblas_dgemm64_t)s;
    else if (void *s2 = dlsym(h, "cblas_dgemm64_")) g_dgemm64 = (cblas_dgemm64_t)s2;
    else if (void *s3 = dlsym(h, "scipy_cblas_dgemm")) g_dgemm32 = (cblas_dgemm32_t)s3;
    else if (void *s4 = dlsym(h, "cblas_dgemm")) g_dgemm32 = (cblas_dgemm32_t)s4;
    else { std::snprintf(g_err, sizeof g_err, "no cblas_dgemm symbol in %s", path); return 1; }
    arma::dgemm_hook() = hook;
    return 0;
}

int ref_matrix_flip(int minor, double *G, int64_t n, int64_t p, double *AF, double *MAF) {
    return guarded([&] {
        List l = minor ? matrix_flip_minor(to_mat(G, n, p)) : matrix_flip_mean(to_mat(G, n, p));
        out(l, "Geno", G); out(l, "AF", AF); out(l, "MAF", MAF);
    });
}
int ref_score_test(const double *G, int64_t n, int64_t p, const orc_csc *Si, const double *SiX, int64_t q,
                   const double *cov, const double *res, double *Score, double *Score_se, double *pvalue_log,
                   double *Est, double *Est_se) {
    return guarded([&] {
        out5(Individual_Score_Test(to_mat(G, n, p), to_sp(Si), to_mat(SiX, n, q), to_mat(cov, q, q), to_vec(res, n)),
             Score, Score_se, pvalue_log, Est, Est_se);
    });
}
int ref_score_test_sp(const orc_csc *G, const orc_csc *Si, const double *SiX, int64_t q, const double *cov,
                      const double *res, double *Score, double *Score_se, double *pvalue_log, double *Est, double *Est_se) {
    int64_t n = G->n_rows;
    return guarded([&] {
        out5(Individual_Score_Test_sp(to_sp(G), to_sp(Si), to_mat(SiX, n, q), to_mat(cov, q, q), to_vec(res, n)),
             Score, Score_se, pvalue_log, Est, Est_se);
    });
}
int ref_score_test_denseGRM(const double *G, int64_t n, int64_t p, const double *P, const double *res,
                            double *Score, double *Score_se, double *pvalue_log, double *Est, double *Est_se) {
    return guarded([&] {
        out5(Individual_Score_Test_denseGRM(to_mat(G, n, p), to_mat(P, n, n), to_vec(res, n)),
             Score, Score_se, pvalue_log, Est, Est_se);
    });
}
int ref_score_test_sp_denseGRM(const orc_csc *G, const double *P, const double *res,
                               double *Score, double *Score_se, double *pvalue_log, double *Est, double *Est_se) {
    int64_t n = G->n_rows;
    return guarded([&] {
        out5(Individual_Score_Test_sp_denseGRM(to_sp(G), to_mat(P, n, n), to_vec(res, n)),
             Score, Score_se, pvalue_log, Est, Est_se);
    });
}
int ref_score_test_multi(const double *G, int64_t n, int64_t p, const orc_csc *Si, const double *SiX, int64_t q,
                         const double *cov, const double *res, int k, double *Score, double *pvalue_log) {
    return guarded([&] {
        List l = Individual_Score_Test_multi(to_mat(G, n, p), to_sp(Si), to_mat(SiX, n, q), to_mat(cov, q, q), to_vec(res, n), k);
        out(l, "Score", Score); out(l, "pvalue_log", pvalue_log);
    });
}
int ref_score_test_sp_multi(const orc_csc *G, const orc_csc *Si, const double *SiX, int64_t q,
                            const double *cov, const double *res, int k, double *Score, double *pvalue_log) {
    int64_t n = G->n_rows;
    return guarded([&] {
        List l = Individual_Score_Test_sp_multi(to_sp(G), to_sp(Si), to_mat(SiX, n, q), to_mat(cov, q, q), to_vec(res, n), k);
        out(l, "Score", Score); out(l, "pvalue_log", pvalue_log);
    });
}
int ref_score_test_denseGRM_multi(const double *G, int64_t n, int64_t p, const double *P, const double *res,
                                  int k, double *Score, double *pvalue_log) {
    return guarded([&] {
        List l = Individual_Score_Test_denseGRM_multi(to_mat(G, n, p), to_mat(P, n, n), to_vec(res, n), k);
        out(l, "Score", Score); out(l, "pvalue_log", pvalue_log);
    });
}
int ref_score_test_sp_denseGRM_multi(const orc_csc *G, const double *P, const double *res,
                                     int k, double *Score, double *pvalue_log) {
    int64_t n = G->n_rows;
    return guarded([&] {
        List l = Individual_Score_Test_sp_denseGRM_multi(to_sp(G), to_mat(P, n, n), to_vec(res, n), k);
        out(l, "Score", Score); out(l, "pvalue_log", pvalue_log);
    });
}
int ref_score_test_cond(const double *G, int64_t n, int64_t p, const orc_csc *Si, const double *SiX, int64_t q,
                        const double *cov, const double *A, int64_t m, const double *res,
                        double *Score, double *Score_se, double *pvalue_log, double *Est, double *Est_se) {
    return guarded([&] {
        out5(Individual_Score_Test_cond(to_mat(G, n, p), to_sp(Si), to_mat(SiX, n, q), to_mat(cov, q, q),
                                        to_mat(A, n, m), to_vec(res, n)),
             Score, Score_se, pvalue_log, Est, Est_se);
    });
}
int ref_score_test_cond_multi(const double *G, int64_t n, int64_t p, const orc_csc *Si, const double *SiX, int64_t q,
                              const double *cov, const double *A, int64_t m, const double *res, int k,
                              double *Score, double *pvalue_log) {
    return guarded([&] {
        List l = Individual_Score_Test_cond_multi(to_mat(G, n, p), to_sp(Si), to_mat(SiX, n, q), to_mat(cov, q, q),
                                                  to_mat(A, n, m), to_vec(res, n), k);
        out(l, "Score", Score); out(l, "pvalue_log", pvalue_log);
    });
}
int ref_score_test_SPA(const double *G, int64_t n, int64_t p, const double *XW, const double *XXWX_inv, int64_t q,
                       const double *res, const double *muhat, double tol, int max_iter, double *pvalue) {
    return guarded([&] {
        arma::vec r = Individual_Score_Test_SPA(to_mat(G, n, p), to_mat(XW, q, n), to_mat(XXWX_inv, n, q),
                                                to_vec(res, n), to_vec(muhat, n), tol, max_iter);
        std::memcpy(pvalue, r.memptr(), sizeof(double) * p);
    });
}

} // extern "C"
