/*
 * staar_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the 13 native functions of xihaoli/STAARpipeline
 * (reference tree /root/reference, v0.9.8) that make up the per-variant score
 * test hot path.  Every function cites the reference file:line it follows.
 *
 * PARITY STATUS: "parity unpinned" by the reference — the reference ships no
 * tests, examples or golden vectors and cannot be built
 * with its real dependencies here (R, Rcpp, RcppArmadillo, libRmath absent).
 * The oracle is pinned instead by (1) oracle/_ref: the reference's OWN .cpp
 * sources compiled where they lie against a minimal stand-in for the
 * RcppArmadillo/Rmath API (oracle/ref_stub/), (2) an independent numpy route
 * (oracle/np_oracle.py), (3) mpmath for the chi-square / normal tails.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
 * leg may load this library.  The product (staarpipeline_b200/) never does.
 *
 * Conventions: all dense matrices are column-major FP64 (Armadillo layout);
 * sparse matrices are CSC with 64-bit indices (ARMA_64BIT_WORD=1,
 * reference src/Makevars:2).  Return value 0 = ok, non-zero = the condition
 * under which the reference would raise an R error (e.g. singular inv()).
 */
#ifndef STAAR_ORACLE_H
#define STAAR_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- Rmath substitutes (reference calls R::pchisq / R::pnorm) ---------- */
/* log of the upper tail of chi-square(df) at x: R::pchisq(x, df, false, true) */
double orc_pchisq_upper_log(double x, double df);
/* R::pnorm(x, 0, 1, lower, false) */
double orc_pnorm(double x, int lower);

/* ---- genotype flip / impute (src/matrix_flip_mean.cpp:8-74, _minor.cpp:8-99) */
void orc_matrix_flip_mean(double *G, int64_t n, int64_t p, double *AF, double *MAF);
void orc_matrix_flip_minor(double *G, int64_t n, int64_t p, double *AF, double *MAF);

/* ---- score tests ------------------------------------------------------- */
/* sparse matrix = (values, row_idx, col_ptr) CSC, n_rows x n_cols */
typedef struct {
    int64_t n_rows, n_cols;
    const double *values;
    const uint64_t *row_idx;
    const uint64_t *col_ptr;
} orc_csc;

/* src/Individual_Score_Test.cpp:9-67 (dense G, sparse Sigma_i) */
int orc_score_test(const double *G, int64_t n, int64_t p, const orc_csc *Sigma_i,
                   const double *Sigma_iX, int64_t q, const double *cov, const double *residuals,
                   double *Score, double *Score_se, double *pvalue_log, double *Est, double *Est_se);
/* src/Individual_Score_Test_sp.cpp:9-67 (sparse G) */
int orc_score_test_sp(const orc_csc *G, const orc_csc *Sigma_i,
                      const double *Sigma_iX, int64_t q, const double *cov, const double *residuals,
                      double *Score, double *Score_se, double *pvalue_log, double *Est, double *Est_se);
/* src/Individual_Score_Test_denseGRM.cpp:9-61 / _sp_denseGRM.cpp:9-61 */
int orc_score_test_denseGRM(const double *G, int64_t n, int64_t p, const double *P, const double *residuals,
                            double *Score, double *Score_se, double *pvalue_log, double *Est, double *Est_se);
int orc_score_test_sp_denseGRM(const orc_csc *G, const double *P, const double *residuals,
                               double *Score, double *Score_se, double *pvalue_log, double *Est, double *Est_se);
/* src/Individual_Score_Test_multi.cpp:9-68 / _sp_multi.cpp:9-68; G is (n x p), p = n_pheno*pp */
int orc_score_test_multi(const double *G, int64_t n, int64_t p, const orc_csc *Sigma_i,
                         const double *Sigma_iX, int64_t q, const double *cov, const double *residuals,
                         int n_pheno, double *Score, double *pvalue_log);
int orc_score_test_sp_multi(const orc_csc *G, const orc_csc *Sigma_i,
                            const double *Sigma_iX, int64_t q, const double *cov, const double *residuals,
                            int n_pheno, double *Score, double *pvalue_log);
/* src/Individual_Score_Test_denseGRM_multi.cpp:9-62 / _sp_denseGRM_multi.cpp:9-62 */
int orc_score_test_denseGRM_multi(const double *G, int64_t n, int64_t p, const double *P, const double *residuals,
                                  int n_pheno, double *Score, double *pvalue_log);
int orc_score_test_sp_denseGRM_multi(const orc_csc *G, const double *P, const double *residuals,
                                     int n_pheno, double *Score, double *pvalue_log);
/* src/Individual_Score_Test_cond.cpp:10-75 */
int orc_score_test_cond(const double *G, int64_t n, int64_t p, const orc_csc *Sigma_i,
                        const double *Sigma_iX, int64_t q, const double *cov,
                        const double *X_adj, int64_t m, const double *residuals,
                        double *Score, double *Score_se, double *pvalue_log, double *Est, double *Est_se);
/* src/Individual_Score_Test_cond_multi.cpp:10-74 */
int orc_score_test_cond_multi(const double *G, int64_t n, int64_t p, const orc_csc *Sigma_i,
                              const double *Sigma_iX, int64_t q, const double *cov,
                              const double *X_adj, int64_t m, const double *residuals,
                              int n_pheno, double *Score, double *pvalue_log);
/* src/Individual_Score_Test_SPA.cpp:38-533.  XW is q x n, XXWX_inv is n x q.
 * trace (optional, may be NULL): per variant 4 int64 counters
 * {#K evaluations, #K1, #K2, branch bits} used by the iteration-trace tests. */
int orc_score_test_SPA(const double *G, int64_t n, int64_t p, const double *XW, const double *XXWX_inv,
                       int64_t q, const double *residuals, const double *muhat, double tol, int max_iter,
                       double *pvalue, int64_t *trace);

/* ---- helpers shared with the tests ------------------------------------- */
/* LU (partial pivoting) determinant / inverse of a k x k column-major matrix;
 * stands in for arma::det / arma::inv (LAPACK getrf/getri).  inv returns
 * non-zero when a pivot is exactly zero (Armadillo would throw). */
double orc_det(const double *A, int k);
int orc_inv(const double *A, int k, double *Ainv);

#ifdef __cplusplus
}
#endif
#endif
