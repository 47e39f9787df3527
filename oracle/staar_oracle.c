/*
 * staar_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 * See staar_oracle.h for scope, parity status ("parity unpinned" by the
 * reference; pinned here against oracle/_ref, numpy and mpmath) and rules.
 *
 * Plain C99, in-order loops, no BLAS: summation order is the textbook
 * left-to-right order of the reference's own hand loops where it has them
 * (flip, SPA) and the natural order of the Armadillo expressions elsewhere.
 * Citations are file:line under /root/reference/.
 */
#include "staar_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================= */
/* Rmath substitutes                                                        */
/* ======================================================================= */

/* log Q(a, y), Q = regularised upper incomplete gamma.  Stands in for
 * R::pchisq(x, df, lower=false, log=true) = log Q(df/2, x/2), used at
 * src/Individual_Score_Test.cpp:57 and every sibling.  Series for P below
 * y < a+1 (then log1p(-P)), modified-Lentz continued fraction above; both
 * are evaluated in the log domain so the result stays finite for
 * log p ~ -1e4 (scipy.stats.chi2.logsf underflows to -inf).
 * Checked against mpmath (50 digits) in tests/test_oracle_rmath.py. */
static double log_gamma_q(double a, double y)
{
    if (isnan(a) || isnan(y)) return NAN;
    if (y <= 0.0) return 0.0;
    if (isinf(y)) return -INFINITY;
    if (y < a + 1.0) {
        double ap = a, term = 1.0 / a, sum = term;
        for (int i = 0; i < 100000; i++) {
            ap += 1.0;
            term *= y / ap;
            sum += term;
            if (fabs(term) < fabs(sum) * 1e-17) break;
        }
        double logP = a * log(y) - y - lgamma(a) + log(sum);
        return log1p(-exp(logP));
    } else {
        const double tiny = 1e-300;
        double b = y + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
        for (int i = 1; i < 100000; i++) {
            double an = -(double)i * ((double)i - a);
            b += 2.0;
            d = an * d + b; if (fabs(d) < tiny) d = tiny;
            c = b + an / c; if (fabs(c) < tiny) c = tiny;
            d = 1.0 / d;
            double del = d * c;
            h *= del;
            if (fabs(del - 1.0) < 1e-16) break;
        }
        return a * log(y) - y - lgamma(a) + log(h);
    }
}

double orc_pchisq_upper_log(double x, double df)
{
    return log_gamma_q(0.5 * df, 0.5 * x);
}

/* R::pnorm(x,0,1,lower,false) (src/Individual_Score_Test_SPA.cpp:151,282).
 * R's pnorm_both returns exactly 0 / 1 beyond these cut-offs (nmath/pnorm.c);
 * inside them 0.5*erfc() agrees with Cody's algorithm to ~1 ulp. */
double orc_pnorm(double x, int lower)
{
    if (isnan(x)) return NAN;
    if (lower) {
        if (x <= -37.5193) return 0.0;
        if (x >= 8.2924) return 1.0;
        return 0.5 * erfc(-x * M_SQRT1_2);
    } else {
        if (x >= 37.5193) return 0.0;
        if (x <= -8.2924) return 1.0;
        return 0.5 * erfc(x * M_SQRT1_2);
    }
}

/* ======================================================================= */
/* matrix_flip_mean / matrix_flip_minor                                     */
/* ======================================================================= */

/* src/matrix_flip_mean.cpp:23-41 — AF over entries with G > -1 (NaN and any
 * value <= -1 are "missing"), AF = sum/2/num. */
static void af_nonmissing(const double *G, int64_t n, int64_t p, double *AF)
{
    for (int64_t i = 0; i < p; i++) {
        const double *g = G + i * n;
        double num = 0, af = 0;
        for (int64_t j = 0; j < n; j++) {
            if (g[j] > -1) { af = af + g[j]; num = num + 1; }
        }
        if (num > 0) af = af / 2 / num;
        AF[i] = af;
    }
}

/* src/matrix_flip_mean.cpp:56-70 / matrix_flip_minor.cpp:81-95 */
static void flip_to_minor(double *G, int64_t n, int64_t p, const double *AF, double *MAF)
{
    for (int64_t i = 0; i < p; i++) {
        if (AF[i] <= 0.5) {
            MAF[i] = AF[i];
        } else {
            MAF[i] = 1 - AF[i];
            double *g = G + i * n;
            for (int64_t j = 0; j < n; j++) g[j] = 2 - g[j];
        }
    }
}

void orc_matrix_flip_mean(double *G, int64_t n, int64_t p, double *AF, double *MAF)
{
    af_nonmissing(G, n, p, AF);                       /* :23-41 */
    for (int64_t i = 0; i < p; i++) {                 /* :44-53 */
        double *g = G + i * n;
        for (int64_t j = 0; j < n; j++)
            if (!(g[j] > -1)) g[j] = AF[i] * 2;
    }
    flip_to_minor(G, n, p, AF, MAF);                  /* :56-70 */
}

void orc_matrix_flip_minor(double *G, int64_t n, int64_t p, double *AF, double *MAF)
{
    af_nonmissing(G, n, p, AF);                       /* :23-41 */
    for (int64_t i = 0; i < p; i++) {                 /* :44-65 */
        double *g = G + i * n;
        double fill = (AF[i] <= 0.5) ? 0.0 : 2.0;
        for (int64_t j = 0; j < n; j++)
            if (!(g[j] > -1)) g[j] = fill;
    }
    for (int64_t i = 0; i < p; i++) {                 /* :68-78 */
        const double *g = G + i * n;
        double af = 0;
        for (int64_t j = 0; j < n; j++) af = af + g[j];
        AF[i] = af / 2 / (double)n;
    }
    flip_to_minor(G, n, p, AF, MAF);                  /* :81-95 */
}

/* ======================================================================= */
/* small dense linear algebra (stand-in for arma::det / arma::inv)          */
/* ======================================================================= */

/* LU with partial pivoting in place; returns sign (0 if an exact-zero pivot). */
static int lu_decomp(double *A, int k, int *piv)
{
    int sign = 1;
    for (int c = 0; c < k; c++) {
        int pr = c; double best = fabs(A[c + c * k]);
        for (int r = c + 1; r < k; r++) {
            double v = fabs(A[r + c * k]);
            if (v > best) { best = v; pr = r; }
        }
        piv[c] = pr;
        if (pr != c) {
            for (int cc = 0; cc < k; cc++) {
                double t = A[c + cc * k]; A[c + cc * k] = A[pr + cc * k]; A[pr + cc * k] = t;
            }
            sign = -sign;
        }
        double d = A[c + c * k];
        if (d == 0.0) return 0;
        for (int r = c + 1; r < k; r++) {
            double f = A[r + c * k] / d;
            A[r + c * k] = f;
            for (int cc = c + 1; cc < k; cc++) A[r + cc * k] -= f * A[c + cc * k];
        }
    }
    return sign;
}

double orc_det(const double *A, int k)
{
    double *W = (double *)malloc(sizeof(double) * k * k);
    int *piv = (int *)malloc(sizeof(int) * k);
    memcpy(W, A, sizeof(double) * k * k);
    int sign = lu_decomp(W, k, piv);
    double det = (double)sign;
    if (sign != 0) for (int c = 0; c < k; c++) det *= W[c + c * k];
    free(W); free(piv);
    return det;
}

int orc_inv(const double *A, int k, double *Ainv)
{
    double *W = (double *)malloc(sizeof(double) * k * k);
    int *piv = (int *)malloc(sizeof(int) * k);
    memcpy(W, A, sizeof(double) * k * k);
    int sign = lu_decomp(W, k, piv);
    if (sign == 0) { free(W); free(piv); return 1; }
    for (int c = 0; c < k; c++) {
        double *x = Ainv + (size_t)c * k;
        for (int r = 0; r < k; r++) x[r] = (r == c) ? 1.0 : 0.0;
        for (int r = 0; r < k; r++) {           /* apply row swaps */
            int pr = piv[r];
            if (pr != r) { double t = x[r]; x[r] = x[pr]; x[pr] = t; }
        }
        for (int r = 0; r < k; r++)             /* L y = b */
            for (int cc = 0; cc < r; cc++) x[r] -= W[r + cc * k] * x[cc];
        for (int r = k - 1; r >= 0; r--) {      /* U x = y */
            for (int cc = r + 1; cc < k; cc++) x[r] -= W[r + cc * k] * x[cc];
            x[r] /= W[r + r * k];
        }
    }
    free(W); free(piv);
    return 0;
}

/* ======================================================================= */
/* sparse helpers                                                           */
/* ======================================================================= */

/* y = x^T S  (length n_cols) — the `trans(G)*Sigma_i` factor */
static void vT_csc(const orc_csc *S, const double *x, double *y)
{
    for (int64_t c = 0; c < S->n_cols; c++) {
        double acc = 0;
        for (uint64_t e = S->col_ptr[c]; e < S->col_ptr[c + 1]; e++)
            acc += x[S->row_idx[e]] * S->values[e];
        y[c] = acc;
    }
}

/* y = S x  (length n_rows) — the `Sigma_i*G` factor */
static void csc_v(const orc_csc *S, const double *x, double *y)
{
    for (int64_t r = 0; r < S->n_rows; r++) y[r] = 0;
    for (int64_t c = 0; c < S->n_cols; c++) {
        double xc = x[c];
        if (xc == 0.0) continue;
        for (uint64_t e = S->col_ptr[c]; e < S->col_ptr[c + 1]; e++)
            y[S->row_idx[e]] += S->values[e] * xc;
    }
}

static void csc_col_to_dense(const orc_csc *G, int64_t i, double *g)
{
    for (int64_t r = 0; r < G->n_rows; r++) g[r] = 0;
    for (uint64_t e = G->col_ptr[i]; e < G->col_ptr[i + 1]; e++) g[G->row_idx[e]] = G->values[e];
}

static double dot(const double *a, const double *b, int64_t n)
{
    double s = 0;
    for (int64_t j = 0; j < n; j++) s += a[j] * b[j];
    return s;
}

/* t = M^T g, M is n x q column-major */
static void matT_v(const double *M, int64_t n, int64_t q, const double *g, double *t)
{
    for (int64_t c = 0; c < q; c++) t[c] = dot(M + c * n, g, n);
}

/* x^T C y for a q x q column-major C, associated as (x^T C) y */
static double bilinear(const double *x, const double *C, const double *y, int64_t q)
{
    double s = 0;
    for (int64_t b = 0; b < q; b++) {
        double xb = 0;
        for (int64_t a = 0; a < q; a++) xb += x[a] * C[a + b * q];
        s += xb * y[b];
    }
    return s;
}

/* ======================================================================= */
/* per-variant statistics loop shared by all single-trait tests             */
/* src/Individual_Score_Test.cpp:45-64 (identical in _sp, _denseGRM, _cond) */
/* ======================================================================= */
static void single_trait_stats(double U, double C, double *Score_se, double *pvalue_log,
                               double *Est, double *Est_se)
{
    if (C == 0) {
        *pvalue_log = 0; *Score_se = 0; *Est = 0; *Est_se = 0;
    } else {
        double test_stat = pow(U, 2) / C;
        *pvalue_log = -orc_pchisq_upper_log(test_stat, 1);
        *Score_se = sqrt(C);
        *Est = U / C;
        *Est_se = 1 / *Score_se;
    }
}

/* k-trait block test: src/Individual_Score_Test_multi.cpp:45-65 */
static void multi_trait_stat(const double *Ublk, const double *Cblk, int k, double *pvalue_log)
{
    if (orc_det(Cblk, k) == 0) { *pvalue_log = 0; return; }
    double *Ci = (double *)malloc(sizeof(double) * k * k);
    if (orc_inv(Cblk, k, Ci) != 0) { *pvalue_log = 0; free(Ci); return; }
    double stat = bilinear(Ublk, Ci, Ublk, k);
    *pvalue_log = -orc_pchisq_upper_log(stat, (double)k);
    free(Ci);
}

/* ======================================================================= */
/* column access abstraction: dense or CSC genotype matrix                  */
/* ======================================================================= */
typedef struct {
    const double *dense; const orc_csc *sp; int64_t n, p;
} gmat;

static const double *gcol(const gmat *G, int64_t i, double *scratch)
{
    if (G->dense) return G->dense + i * G->n;
    csc_col_to_dense(G->sp, i, scratch);
    return scratch;
}

enum { ASSOC_GT_S_G = 0,   /* (trans(G)*Sigma_i)*G          Individual_Score_Test.cpp:43        */
       ASSOC_T_SG_G_T = 1, /* trans(trans(Sigma_i*G)*G)      Individual_Score_Test_sp.cpp:43     */
       ASSOC_SG_T_G = 2 }; /* trans(Sigma_i*G)*G             Individual_Score_Test_cond.cpp:49   */

/* sparse-Sigma family, single trait */
static int score_sparse_sigma(const gmat *G, int assoc, const orc_csc *Sigma_i,
                              const double *Sigma_iX, int64_t q, const double *cov, const double *residuals,
                              double *Score, double *Score_se, double *pvalue_log, double *Est, double *Est_se)
{
    int64_t n = G->n, p = G->p;
    double *scratch = (double *)malloc(sizeof(double) * n);
    double *w = (double *)malloc(sizeof(double) * n);
    double *t = (double *)malloc(sizeof(double) * (q > 0 ? q : 1));
    for (int64_t i = 0; i < p; i++) {
        const double *g = gcol(G, i, scratch);
        double U = dot(residuals, g, n);                       /* :17 */
        matT_v(Sigma_iX, n, q, g, t);                          /* :42 */
        if (assoc == ASSOC_GT_S_G) vT_csc(Sigma_i, g, w); else csc_v(Sigma_i, g, w);
        double C = dot(w, g, n) - bilinear(t, cov, t, q);      /* :43 (diagonal entry) */
        Score[i] = U;
        single_trait_stats(U, C, &Score_se[i], &pvalue_log[i], &Est[i], &Est_se[i]);
    }
    free(scratch); free(w); free(t);
    return 0;
}

int orc_score_test(const double *Gd, int64_t n, int64_t p, const orc_csc *Sigma_i,
                   const double *Sigma_iX, int64_t q, const double *cov, const double *residuals,
                   double *Score, double *Score_se, double *pvalue_log, double *Est, double *Est_se)
{
    gmat G = { Gd, NULL, n, p };
    return score_sparse_sigma(&G, ASSOC_GT_S_G, Sigma_i, Sigma_iX, q, cov, residuals,
                              Score, Score_se, pvalue_log, Est, Est_se);
}

int orc_score_test_sp(const orc_csc *Gs, const orc_csc *Sigma_i,
                      const double *Sigma_iX, int64_t q, const double *cov, const double *residuals,
                      double *Score, double *Score_se, double *pvalue_log, double *Est, double *Est_se)
{
    gmat G = { NULL, Gs, Gs->n_rows, Gs->n_cols };
    return score_sparse_sigma(&G, ASSOC_T_SG_G_T, Sigma_i, Sigma_iX, q, cov, residuals,
                              Score, Score_se, pvalue_log, Est, Est_se);
}

/* dense-P family: Cov = trans(P*G)*G  (src/Individual_Score_Test_denseGRM.cpp:37) */
static void dense_P_times(const double *P, int64_t n, const double *g, double *w)
{
    for (int64_t r = 0; r < n; r++) w[r] = 0;
    for (int64_t c = 0; c < n; c++) {
        double gc = g[c];
        if (gc == 0.0) continue;
        const double *Pc = P + c * n;
        for (int64_t r = 0; r < n; r++) w[r] += Pc[r] * gc;
    }
}

static int score_dense_P(const gmat *G, const double *P, const double *residuals,
                         double *Score, double *Score_se, double *pvalue_log, double *Est, double *Est_se)
{
    int64_t n = G->n, p = G->p;
    double *scratch = (double *)malloc(sizeof(double) * n);
    double *w = (double *)malloc(sizeof(double) * n);
    for (int64_t i = 0; i < p; i++) {
        const double *g = gcol(G, i, scratch);
        double U = dot(residuals, g, n);                       /* :17 */
        dense_P_times(P, n, g, w);
        double C = dot(w, g, n);                               /* :37 (diagonal entry) */
        Score[i] = U;
        single_trait_stats(U, C, &Score_se[i], &pvalue_log[i], &Est[i], &Est_se[i]);
    }
    free(scratch); free(w);
    return 0;
}

int orc_score_test_denseGRM(const double *Gd, int64_t n, int64_t p, const double *P, const double *residuals,
                            double *Score, double *Score_se, double *pvalue_log, double *Est, double *Est_se)
{
    gmat G = { Gd, NULL, n, p };
    return score_dense_P(&G, P, residuals, Score, Score_se, pvalue_log, Est, Est_se);
}

int orc_score_test_sp_denseGRM(const orc_csc *Gs, const double *P, const double *residuals,
                               double *Score, double *Score_se, double *pvalue_log, double *Est, double *Est_se)
{
    gmat G = { NULL, Gs, Gs->n_rows, Gs->n_cols };
    return score_dense_P(&G, P, residuals, Score, Score_se, pvalue_log, Est, Est_se);
}

/* ---- multi-trait ------------------------------------------------------- */
/* src/Individual_Score_Test_multi.cpp:18-65.  For variant i the k x k block
 * Cov(id,id), id(k) = k*pp+i (:47-52), needs the cross terms between the k
 * columns {i, pp+i, ...}; computed here for general G (the R caller passes
 * I_k (x) G0, R/Individual_Analysis.R:267). */
static int score_multi(const gmat *G, int assoc, const orc_csc *Sigma_i, const double *P,
                       const double *Sigma_iX, int64_t q, const double *cov, const double *residuals,
                       int k, double *Score, double *pvalue_log)
{
    int64_t n = G->n, p = G->p, pp = p / k;
    double *cols = (double *)malloc(sizeof(double) * n * k);
    double *ws = (double *)malloc(sizeof(double) * n * k);
    double *ts = (double *)malloc(sizeof(double) * (q > 0 ? q : 1) * k);
    double *Cb = (double *)malloc(sizeof(double) * k * k);
    double *Ub = (double *)malloc(sizeof(double) * k);
    double *scratch = (double *)malloc(sizeof(double) * n);
    for (int64_t i = 0; i < p; i++) {                          /* :18  Uscore = trans(residuals)*G */
        const double *g = gcol(G, i, scratch);
        Score[i] = dot(residuals, g, n);
    }
    for (int64_t i = 0; i < pp; i++) {
        for (int a = 0; a < k; a++) {
            const double *g = gcol(G, a * pp + i, scratch);
            memcpy(cols + (size_t)a * n, g, sizeof(double) * n);
            Ub[a] = Score[a * pp + i];
            if (P) dense_P_times(P, n, g, ws + (size_t)a * n);
            else if (assoc == ASSOC_GT_S_G) vT_csc(Sigma_i, g, ws + (size_t)a * n);
            else csc_v(Sigma_i, g, ws + (size_t)a * n);
            if (!P) matT_v(Sigma_iX, n, q, g, ts + (size_t)a * q);
        }
        for (int a = 0; a < k; a++)
            for (int b = 0; b < k; b++) {
                double c1;
                if (assoc == ASSOC_T_SG_G_T) c1 = dot(ws + (size_t)b * n, cols + (size_t)a * n, n);
                else c1 = dot(ws + (size_t)a * n, cols + (size_t)b * n, n);
                double c2 = P ? 0.0 : bilinear(ts + (size_t)a * q, cov, ts + (size_t)b * q, q);
                Cb[a + b * k] = P ? c1 : c1 - c2;
            }
        multi_trait_stat(Ub, Cb, k, &pvalue_log[i]);
    }
    free(cols); free(ws); free(ts); free(Cb); free(Ub); free(scratch);
    return 0;
}

int orc_score_test_multi(const double *Gd, int64_t n, int64_t p, const orc_csc *Sigma_i,
                         const double *Sigma_iX, int64_t q, const double *cov, const double *residuals,
                         int n_pheno, double *Score, double *pvalue_log)
{
    gmat G = { Gd, NULL, n, p };
    return score_multi(&G, ASSOC_GT_S_G, Sigma_i, NULL, Sigma_iX, q, cov, residuals, n_pheno, Score, pvalue_log);
}

int orc_score_test_sp_multi(const orc_csc *Gs, const orc_csc *Sigma_i,
                            const double *Sigma_iX, int64_t q, const double *cov, const double *residuals,
                            int n_pheno, double *Score, double *pvalue_log)
{
    gmat G = { NULL, Gs, Gs->n_rows, Gs->n_cols };
    return score_multi(&G, ASSOC_T_SG_G_T, Sigma_i, NULL, Sigma_iX, q, cov, residuals, n_pheno, Score, pvalue_log);
}

int orc_score_test_denseGRM_multi(const double *Gd, int64_t n, int64_t p, const double *P, const double *residuals,
                                  int n_pheno, double *Score, double *pvalue_log)
{
    gmat G = { Gd, NULL, n, p };
    return score_multi(&G, ASSOC_SG_T_G, NULL, P, NULL, 0, NULL, residuals, n_pheno, Score, pvalue_log);
}

int orc_score_test_sp_denseGRM_multi(const orc_csc *Gs, const double *P, const double *residuals,
                                     int n_pheno, double *Score, double *pvalue_log)
{
    gmat G = { NULL, Gs, Gs->n_rows, Gs->n_cols };
    return score_multi(&G, ASSOC_SG_T_G, NULL, P, NULL, 0, NULL, residuals, n_pheno, Score, pvalue_log);
}

/* ---- conditional ------------------------------------------------------- */
/* src/Individual_Score_Test_cond.cpp:48-49 / _cond_multi.cpp:45-46.
 * k == 0 selects the single-trait loop (:51-72), k >= 1 the block test. */
static int score_cond(const double *Gd, int64_t n, int64_t p, const orc_csc *Sigma_i,
                      const double *Sigma_iX, int64_t q, const double *cov,
                      const double *A, int64_t m, const double *residuals, int k,
                      double *Score, double *Score_se, double *pvalue_log, double *Est, double *Est_se)
{
    int rc = 0;
    double *SiA = (double *)malloc(sizeof(double) * n * m);       /* Sigma_i*X_adj */
    double *SXtA = (double *)malloc(sizeof(double) * q * m);      /* trans(Sigma_iX)*X_adj */
    double *T1 = (double *)malloc(sizeof(double) * q * m);        /* cov*(...) */
    double *PA = (double *)malloc(sizeof(double) * n * m);
    double *AtA = (double *)malloc(sizeof(double) * m * m);
    double *M = (double *)malloc(sizeof(double) * m * m);
    double *PAtA = (double *)malloc(sizeof(double) * m * m);
    double *MQM = (double *)malloc(sizeof(double) * m * m);
    double *tmp = (double *)malloc(sizeof(double) * m * m);
    for (int64_t c = 0; c < m; c++) {
        csc_v(Sigma_i, A + c * n, SiA + c * n);
        matT_v(Sigma_iX, n, q, A + c * n, SXtA + c * q);
    }
    for (int64_t c = 0; c < m; c++)
        for (int64_t a = 0; a < q; a++) {
            double s = 0;
            for (int64_t b = 0; b < q; b++) s += cov[a + b * q] * SXtA[b + c * q];
            T1[a + c * q] = s;
        }
    for (int64_t c = 0; c < m; c++)                              /* :48 PX_adj */
        for (int64_t j = 0; j < n; j++) {
            double s = 0;
            for (int64_t a = 0; a < q; a++) s += Sigma_iX[j + a * n] * T1[a + c * q];
            PA[j + c * n] = SiA[j + c * n] - s;
        }
    for (int64_t a = 0; a < m; a++)
        for (int64_t b = 0; b < m; b++) {
            AtA[a + b * m] = dot(A + a * n, A + b * n, n);
            PAtA[a + b * m] = dot(PA + a * n, A + b * n, n);      /* trans(PX_adj)*X_adj */
        }
    if (orc_inv(AtA, (int)m, M) != 0) { rc = 1; goto done; }      /* arma::inv throws -> R error */
    /* MQM = M * PAtA * M */
    for (int64_t a = 0; a < m; a++)
        for (int64_t b = 0; b < m; b++) {
            double s = 0;
            for (int64_t c = 0; c < m; c++) s += M[a + c * m] * PAtA[c + b * m];
            tmp[a + b * m] = s;
        }
    for (int64_t a = 0; a < m; a++)
        for (int64_t b = 0; b < m; b++) {
            double s = 0;
            for (int64_t c = 0; c < m; c++) s += tmp[a + c * m] * M[c + b * m];
            MQM[a + b * m] = s;
        }
    {
        int kk = k > 0 ? k : 1;
        int64_t pp = p / kk;
        double *ws = (double *)malloc(sizeof(double) * n * kk);
        double *ts = (double *)malloc(sizeof(double) * q * kk);
        double *as = (double *)malloc(sizeof(double) * m * kk);
        double *bs = (double *)malloc(sizeof(double) * m * kk);
        double *Cb = (double *)malloc(sizeof(double) * kk * kk);
        double *Ub = (double *)malloc(sizeof(double) * kk);
        for (int64_t i = 0; i < p; i++) Score[i] = dot(residuals, Gd + i * n, n);
        for (int64_t i = 0; i < pp; i++) {
            for (int a = 0; a < kk; a++) {
                const double *g = Gd + (a * pp + i) * n;
                csc_v(Sigma_i, g, ws + (size_t)a * n);
                matT_v(Sigma_iX, n, q, g, ts + (size_t)a * q);
                matT_v(A, n, m, g, as + (size_t)a * m);
                matT_v(PA, n, m, g, bs + (size_t)a * m);
                Ub[a] = Score[a * pp + i];
            }
            for (int a = 0; a < kk; a++)
                for (int b = 0; b < kk; b++) {
                    const double *gb = Gd + (b * pp + i) * n;
                    double c = dot(ws + (size_t)a * n, gb, n)
                             - bilinear(ts + (size_t)a * q, cov, ts + (size_t)b * q, q)
                             - bilinear(as + (size_t)a * m, M, bs + (size_t)b * m, m)
                             - bilinear(bs + (size_t)a * m, M, as + (size_t)b * m, m)
                             + bilinear(as + (size_t)a * m, MQM, as + (size_t)b * m, m);
                    Cb[a + b * kk] = c;
                }
            if (k == 0) single_trait_stats(Ub[0], Cb[0], &Score_se[i], &pvalue_log[i], &Est[i], &Est_se[i]);
            else multi_trait_stat(Ub, Cb, kk, &pvalue_log[i]);
        }
        free(ws); free(ts); free(as); free(bs); free(Cb); free(Ub);
    }
done:
    free(SiA); free(SXtA); free(T1); free(PA); free(AtA); free(M); free(PAtA); free(MQM); free(tmp);
    return rc;
}

int orc_score_test_cond(const double *G, int64_t n, int64_t p, const orc_csc *Sigma_i,
                        const double *Sigma_iX, int64_t q, const double *cov,
                        const double *X_adj, int64_t m, const double *residuals,
                        double *Score, double *Score_se, double *pvalue_log, double *Est, double *Est_se)
{
    return score_cond(G, n, p, Sigma_i, Sigma_iX, q, cov, X_adj, m, residuals, 0,
                      Score, Score_se, pvalue_log, Est, Est_se);
}

int orc_score_test_cond_multi(const double *G, int64_t n, int64_t p, const orc_csc *Sigma_i,
                              const double *Sigma_iX, int64_t q, const double *cov,
                              const double *X_adj, int64_t m, const double *residuals,
                              int n_pheno, double *Score, double *pvalue_log)
{
    return score_cond(G, n, p, Sigma_i, Sigma_iX, q, cov, X_adj, m, residuals, n_pheno,
                      Score, NULL, pvalue_log, NULL, NULL);
}

/* ======================================================================= */
/* SPA — src/Individual_Score_Test_SPA.cpp:38-533, literal control flow     */
/* ======================================================================= */
typedef struct {
    const double *mu; const double *g; int64_t n;
    int64_t nK, nK1, nK2, nr_iter;
} spa_ctx;

static int is_bad(double v) { return !isfinite(v); }   /* (R_finite(v)==0)||check_is_na(v) */

static double K_spa(spa_ctx *c, double x)               /* :421-433 */
{
    double res = 0.0; c->nK++;
    for (int64_t i = 0; i < c->n; i++) {
        res = res - x * c->mu[i] * c->g[i];
        res = res + log(1 - c->mu[i] + c->mu[i] * exp(x * c->g[i]));
    }
    return res;
}
static double K_spa_alt(spa_ctx *c, double x)           /* :437-450 */
{
    double res = 0.0; c->nK++;
    for (int64_t i = 0; i < c->n; i++)
        res = res + log((1 - c->mu[i]) * exp(-x * c->g[i]) + c->mu[i]);
    return res;
}
static double K1_spa(spa_ctx *c, double x, double q)    /* :453-467 */
{
    double res = 0.0; c->nK1++;
    for (int64_t i = 0; i < c->n; i++) {
        res = res - c->mu[i] * c->g[i];
        res = res + c->mu[i] * c->g[i] / (c->mu[i] + (1 - c->mu[i]) * exp(-x * c->g[i]));
    }
    res = res - q;
    return res;
}
static double K1_spa_alt(spa_ctx *c, double x, double q) /* :471-485 */
{
    double res = 0.0; c->nK1++;
    for (int64_t i = 0; i < c->n; i++) {
        res = res - c->mu[i] * c->g[i];
        res = res + c->mu[i] * c->g[i] * exp(x * c->g[i]) / (c->mu[i] * exp(x * c->g[i]) + (1 - c->mu[i]));
    }
    res = res - q;
    return res;
}
static double K2_spa(spa_ctx *c, double x)              /* :488-505 */
{
    double res = 0.0; c->nK2++;
    for (int64_t i = 0; i < c->n; i++) {
        double temp1 = c->mu[i] * (1 - c->mu[i]) * pow(c->g[i], 2.0) * exp(-x * c->g[i]);
        double temp2 = c->mu[i] + (1 - c->mu[i]) * exp(-x * c->g[i]);
        res = res + temp1 / pow(temp2, 2.0);
    }
    return res;
}
static double K2_spa_alt(spa_ctx *c, double x)          /* :508-525 */
{
    double res = 0.0; c->nK2++;
    for (int64_t i = 0; i < c->n; i++) {
        double temp1 = c->mu[i] * (1 - c->mu[i]) * pow(c->g[i], 2.0);
        double temp2 = c->mu[i] * exp(x * c->g[i]) + (1 - c->mu[i]);
        res = res + temp1 / pow(temp2, 2.0);
    }
    return res;
}
static double K1_fb(spa_ctx *c, double x, double q)      /* K1 with the "_alt if not finite" fallback */
{
    double v = K1_spa(c, x, q);
    if (is_bad(v)) v = K1_spa_alt(c, x, q);
    return v;
}
static int same_sign(double a, double b) { return (a >= 0) == (b >= 0); }   /* :531-533 */

static double NR_spa(spa_ctx *c, double q, double init, double tol, int max_iter)   /* :158-233 */
{
    double xi = init, xi_update = init;
    if (fabs(K1_spa(c, xi, q)) > tol)
        xi_update = xi - K1_spa(c, xi, q) / K2_spa(c, xi);
    int no_iter = 0;
    double numerator, denominator;
    while (!isnan(xi_update) && isfinite(xi_update) && (fabs(xi_update - xi) > tol)
           && (fabs(K1_spa(c, xi_update, q)) > tol) && (no_iter < max_iter)) {
        no_iter = no_iter + 1;
        xi = xi_update;
        numerator = K1_spa(c, xi, q);
        if (is_bad(numerator)) numerator = K1_spa_alt(c, xi, q);
        denominator = K2_spa(c, xi);
        if (is_bad(denominator)) denominator = K2_spa_alt(c, xi);
        xi_update = xi - numerator / denominator;
    }
    c->nr_iter += no_iter;
    if (is_bad(xi_update)) xi_update = xi;
    double w = 0.0, xhat = 0.0, ki = 0.0;
    if (no_iter == max_iter) {                                   /* :211-230 */
        w = sqrt(2 * (xhat * q - K_spa(c, xhat)));                /* NB: xhat is still 0.0 here */
        xhat = xi_update;
        if (xhat < 0) w = -w;
        ki = xhat * sqrt(K2_spa(c, xhat));
        if (is_bad(ki)) ki = xhat * sqrt(K2_spa_alt(c, xhat));
        if (fabs(w + log(ki / w) / w) < 38) xi_update = 0.0;
    }
    return xi_update;
}

static double tail_from_xhat(spa_ctx *c, double xhat, double q, int lower)  /* :129-152 and :260-283 */
{
    double w = sqrt(2 * (xhat * q - K_spa(c, xhat)));
    if (is_bad(w)) w = sqrt(2 * (xhat * q - K_spa_alt(c, xhat)));
    if (xhat < 0) w = -w;
    double ki = xhat * sqrt(K2_spa(c, xhat));
    if (is_bad(ki)) ki = xhat * sqrt(K2_spa_alt(c, xhat));
    if (fabs(xhat) < 1e-04) return 1.0;
    return orc_pnorm(w + log(ki / w) / w, lower);
}

static double saddle_NR(spa_ctx *c, double q, double tol, int max_iter, int lower)   /* :117-155 */
{
    double xhat = NR_spa(c, q, 0.0, tol, max_iter);
    return tail_from_xhat(c, xhat, q, lower);
}

static void golden_search(spa_ctx *c, double a, double b, double q, double tol, int max_iter, double x[2]) /* :360-417 */
{
    int iter = 0;
    x[0] = a; x[1] = b;
    double phi = (1 + sqrt(5)) / 2;
    double K1x0 = K1_fb(c, x[0], q);
    double K1x1 = K1_fb(c, x[1], q);
    while ((iter < max_iter) && (fabs(x[1] - x[0]) > tol) && same_sign(K1x0, K1x1)) {
        x[1] = b - (b - a) / phi;
        x[0] = a + (b - a) / phi;
        K1x0 = K1_fb(c, x[0], q);
        K1x1 = K1_fb(c, x[1], q);
        if (K1x1 < K1x0) b = x[0]; else a = x[1];
        iter++;
    }
}

static double bisection(spa_ctx *c, double q, double xmin, double xmax, double tol)  /* :290-358 */
{
    double xupper = xmax, xlower = xmin;
    double x0 = 0.0, K1x0 = 1.0;
    double K1left = K1_fb(c, xlower, q);
    double K1right = K1_fb(c, xupper, q);
    if (!is_bad(K1right) && !is_bad(K1left)) {
        if (!same_sign(K1left, K1right)) {
            while ((fabs(xupper - xlower) > tol) && (fabs(K1x0) > tol)) {
                x0 = (xupper + xlower) / 2.0;
                K1x0 = K1_fb(c, x0, q);
                if (K1x0 == 0) break;
                if (same_sign(K1left, K1x0)) xlower = x0; else xupper = x0;
                K1left = K1_fb(c, xlower, q);
                K1right = K1_fb(c, xupper, q);
            }
        }
    }
    return x0;
}

static double saddle_bisect(spa_ctx *c, double q, double tol, int max_iter, double xmin, double xmax, int lower) /* :235-286 */
{
    double xl[2];
    golden_search(c, xmin, xmax, q, tol, max_iter, xl);
    if (xl[0] < xl[1]) { xmin = xl[0]; xmax = xl[1]; } else { xmin = xl[1]; xmax = xl[0]; }
    double xhat = bisection(c, q, xmin, xmax, tol);
    return tail_from_xhat(c, xhat, q, lower);
}

int orc_score_test_SPA(const double *G, int64_t n, int64_t p, const double *XW, const double *XXWX_inv,
                       int64_t q, const double *residuals, const double *muhat, double tol, int max_iter,
                       double *pvalue, int64_t *trace)
{
    double *gt = (double *)malloc(sizeof(double) * n);
    double *b = (double *)malloc(sizeof(double) * (q > 0 ? q : 1));
    const double xmin = -100.0, xmax = 100.0;                      /* :68-69 */
    for (int64_t i = 0; i < p; i++) {
        const double *g = G + i * n;
        /* :61  G_tilde = G - XXWX_inv*(XW*G);  XW is q x n */
        for (int64_t c = 0; c < q; c++) {
            double s = 0;
            for (int64_t j = 0; j < n; j++) s += XW[c + j * q] * g[j];
            b[c] = s;
        }
        for (int64_t j = 0; j < n; j++) {
            double s = 0;
            for (int64_t c = 0; c < q; c++) s += XXWX_inv[j + c * n] * b[c];
            gt[j] = g[j] - s;
        }
        double sum0 = dot(residuals, g, n);                       /* :64 (raw G) */
        spa_ctx c = { muhat, gt, n, 0, 0, 0, 0 };
        int64_t bits = 0;
        double res = 1.0, respart1, respart2;
        respart1 = saddle_NR(&c, fabs(sum0), tol, max_iter, 0);   /* :78 */
        if (isnan(respart1) || respart1 == 1.0) {                 /* :80-83 */
            bits |= 1;
            respart1 = saddle_bisect(&c, fabs(sum0), tol, max_iter, xmin, xmax, 0);
        }
        if (isnan(respart1) || respart1 == 1.0) {                 /* :85-88 */
            bits |= 2;
            res = 1;
        } else {
            respart2 = saddle_NR(&c, -fabs(sum0), tol, max_iter, 1);   /* :90 */
            if (isnan(respart2) || respart2 == 1.0) {
                bits |= 4;
                respart2 = saddle_bisect(&c, -fabs(sum0), tol, max_iter, xmin, xmax, 1);
            }
            if (isnan(respart2) || respart2 == 1.0) { bits |= 8; res = 1; }
            else res = respart1 + respart2;
        }
        if (res > 1) { bits |= 16; res = 1; }                     /* :107-110 */
        pvalue[i] = res;
        if (trace) {
            trace[4 * i + 0] = c.nK; trace[4 * i + 1] = c.nK1; trace[4 * i + 2] = c.nK2;
            trace[4 * i + 3] = bits | (c.nr_iter << 8);
        }
    }
    free(gt); free(b);
    return 0;
}
