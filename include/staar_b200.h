/*
 * staar_b200.h — C ABI of the B200-native score-test hot path.
 *
 * Drop-in boundary for the 13 native functions of xihaoli/STAARpipeline
 * (reference tree: /root/reference, v0.9.8).  Each "frozen-signature" entry
 * point below takes exactly the operands the corresponding Rcpp export
 * receives, as plain host pointers + sizes, and fills caller-allocated output
 * arrays; the reference-side binding (the Rcpp shim that replaces
 * src/*.cpp behind the unchanged RcppExports.cpp) is in
 * staarpipeline_b200/rshim/ and described in INTEGRATION.md.
 *
 * Conventions
 *  - dense matrices: column-major FP64 (Armadillo layout), leading dim = rows;
 *  - sparse matrices: CSC, FP64 values, uint64 row indices / column pointers
 *    (ARMA_64BIT_WORD=1, reference src/Makevars:2);
 *  - every function returns 0 on success, non-zero on error;
 *    staar_last_error() returns the message (the shim turns it into
 *    Rcpp::stop, mirroring BEGIN_RCPP/END_RCPP, src/RcppExports.cpp:20,30);
 *  - calls are synchronous for the caller; one staar_ctx per host thread;
 *  - there is NO CPU fallback: without a CUDA device every compute entry
 *    point fails with STAAR_ERR_CUDA.
 */
#ifndef STAAR_B200_H
#define STAAR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STAAR_OK 0
#define STAAR_ERR_ARG 1      /* bad argument / dimension mismatch (Armadillo would throw)   */
#define STAAR_ERR_CUDA 2     /* CUDA runtime failure, or no device                          */
#define STAAR_ERR_SINGULAR 3 /* inv() of a singular matrix (reachable in _cond, cond.cpp:49) */
#define STAAR_ERR_STATE 4    /* required null-model operands not set                        */

typedef struct staar_ctx staar_ctx;

/* ---- context ---------------------------------------------------------- */
int staar_abi_version(void);
const char *staar_last_error(void);
/* device: CUDA ordinal (one context = one GPU = one rank) */
int staar_ctx_create(int device, staar_ctx **out);
void staar_ctx_destroy(staar_ctx *ctx);
/* stream the context launches on (cudaStream_t as void*), for CUDA-event timing by the caller */
void *staar_ctx_stream(staar_ctx *ctx);
int staar_ctx_synchronize(staar_ctx *ctx);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
uint64_t staar_ctx_launch_count(staar_ctx *ctx);

/* ======================================================================= */
/* A. Frozen-signature entry points (host buffers in, host buffers out).    */
/*    One per Rcpp export; citations are the reference definitions.         */
/* ======================================================================= */

/* matrix_flip_mean(arma::mat G) -> List{Geno, AF, MAF}       src/matrix_flip_mean.cpp:8-74
 * matrix_flip_minor(arma::mat G)                              src/matrix_flip_minor.cpp:8-99
 * G: n x p, NaN or any value <= -1 is missing.  Geno_out (n x p) may alias G or be NULL
 * (AF/MAF only). */
int staar_matrix_flip_mean(staar_ctx *ctx, const double *G, int64_t n, int64_t p,
                           double *Geno_out, double *AF, double *MAF);
int staar_matrix_flip_minor(staar_ctx *ctx, const double *G, int64_t n, int64_t p,
                            double *Geno_out, double *AF, double *MAF);

/* Individual_Score_Test(G, Sigma_i, Sigma_iX, cov, residuals)  src/Individual_Score_Test.cpp:9-67
 * outputs: 5 arrays of length p. */
int staar_individual_score_test(staar_ctx *ctx, const double *G, int64_t n, int64_t p,
                                const double *Si_values, const uint64_t *Si_row_idx, const uint64_t *Si_col_ptr,
                                const double *Sigma_iX, int64_t q, const double *cov, const double *residuals,
                                double *Score, double *Score_se, double *pvalue_log, double *Est, double *Est_se);

/* Individual_Score_Test_sp(sp_mat G, ...)                      src/Individual_Score_Test_sp.cpp:9-67 */
int staar_individual_score_test_sp(staar_ctx *ctx, const double *G_values, const uint64_t *G_row_idx,
                                   const uint64_t *G_col_ptr, int64_t n, int64_t p,
                                   const double *Si_values, const uint64_t *Si_row_idx, const uint64_t *Si_col_ptr,
                                   const double *Sigma_iX, int64_t q, const double *cov, const double *residuals,
                                   double *Score, double *Score_se, double *pvalue_log, double *Est, double *Est_se);

/* Individual_Score_Test_denseGRM(G, P, residuals)              src/Individual_Score_Test_denseGRM.cpp:9-61 */
int staar_individual_score_test_denseGRM(staar_ctx *ctx, const double *G, int64_t n, int64_t p,
                                         const double *P, const double *residuals,
                                         double *Score, double *Score_se, double *pvalue_log, double *Est, double *Est_se);
/* Individual_Score_Test_sp_denseGRM(sp_mat G, P, residuals)    src/Individual_Score_Test_sp_denseGRM.cpp:9-61 */
int staar_individual_score_test_sp_denseGRM(staar_ctx *ctx, const double *G_values, const uint64_t *G_row_idx,
                                            const uint64_t *G_col_ptr, int64_t n, int64_t p,
                                            const double *P, const double *residuals,
                                            double *Score, double *Score_se, double *pvalue_log, double *Est, double *Est_se);

/* Individual_Score_Test_multi(G, Sigma_i, Sigma_iX, cov, residuals, n_pheno)
 *                                                              src/Individual_Score_Test_multi.cpp:9-68
 * G is n x p with p = n_pheno*pp (the R caller passes I_k (x) G0 with n = k*n0).
 * outputs: Score (length p), pvalue_log (length p/n_pheno). */
int staar_individual_score_test_multi(staar_ctx *ctx, const double *G, int64_t n, int64_t p,
                                      const double *Si_values, const uint64_t *Si_row_idx, const uint64_t *Si_col_ptr,
                                      const double *Sigma_iX, int64_t q, const double *cov, const double *residuals,
                                      int n_pheno, double *Score, double *pvalue_log);
/* Individual_Score_Test_sp_multi(sp_mat G, ...)                src/Individual_Score_Test_sp_multi.cpp:9-68 */
int staar_individual_score_test_sp_multi(staar_ctx *ctx, const double *G_values, const uint64_t *G_row_idx,
                                         const uint64_t *G_col_ptr, int64_t n, int64_t p,
                                         const double *Si_values, const uint64_t *Si_row_idx, const uint64_t *Si_col_ptr,
                                         const double *Sigma_iX, int64_t q, const double *cov, const double *residuals,
                                         int n_pheno, double *Score, double *pvalue_log);
/* Individual_Score_Test_denseGRM_multi(G, P, residuals, n_pheno)   src/Individual_Score_Test_denseGRM_multi.cpp:9-62 */
int staar_individual_score_test_denseGRM_multi(staar_ctx *ctx, const double *G, int64_t n, int64_t p,
                                               const double *P, const double *residuals, int n_pheno,
                                               double *Score, double *pvalue_log);
/* Individual_Score_Test_sp_denseGRM_multi(sp_mat G, P, residuals, n_pheno)
 *                                                              src/Individual_Score_Test_sp_denseGRM_multi.cpp:9-62 */
int staar_individual_score_test_sp_denseGRM_multi(staar_ctx *ctx, const double *G_values, const uint64_t *G_row_idx,
                                                  const uint64_t *G_col_ptr, int64_t n, int64_t p,
                                                  const double *P, const double *residuals, int n_pheno,
                                                  double *Score, double *pvalue_log);

/* Individual_Score_Test_cond(G, Sigma_i, Sigma_iX, cov, X_adj, residuals)
 *                                                              src/Individual_Score_Test_cond.cpp:10-75
 * X_adj: n x m.  Returns STAAR_ERR_SINGULAR where arma::inv(X_adj'X_adj) would throw. */
int staar_individual_score_test_cond(staar_ctx *ctx, const double *G, int64_t n, int64_t p,
                                     const double *Si_values, const uint64_t *Si_row_idx, const uint64_t *Si_col_ptr,
                                     const double *Sigma_iX, int64_t q, const double *cov,
                                     const double *X_adj, int64_t m, const double *residuals,
                                     double *Score, double *Score_se, double *pvalue_log, double *Est, double *Est_se);
/* Individual_Score_Test_cond_multi(..., n_pheno)               src/Individual_Score_Test_cond_multi.cpp:10-74 */
int staar_individual_score_test_cond_multi(staar_ctx *ctx, const double *G, int64_t n, int64_t p,
                                           const double *Si_values, const uint64_t *Si_row_idx, const uint64_t *Si_col_ptr,
                                           const double *Sigma_iX, int64_t q, const double *cov,
                                           const double *X_adj, int64_t m, const double *residuals, int n_pheno,
                                           double *Score, double *pvalue_log);

/* Individual_Score_Test_SPA(G, XW, XXWX_inv, residuals, muhat, tol, max_iter) -> p-values (NOT log)
 *                                                              src/Individual_Score_Test_SPA.cpp:38-533
 * XW: q x n, XXWX_inv: n x q.  trace (optional, may be NULL): 4 int64 per variant
 * {passes over n, NR iterations, 0, branch bits} — see oracle/staar_oracle.h. */
int staar_individual_score_test_SPA(staar_ctx *ctx, const double *G, int64_t n, int64_t p,
                                    const double *XW, const double *XXWX_inv, int64_t q,
                                    const double *residuals, const double *muhat, double tol, int max_iter,
                                    double *pvalue, int64_t *trace);

/* ======================================================================= */
/* B. Resident fast path: packed dosage columns + null model kept in HBM.   */
/*    (What bench.py's `value` measures; the Rcpp shim reaches it through   */
/*    the operand cache inside the A-entry points.)                         */
/* ======================================================================= */

/* Packed dosage block: variant-major, 2 bits per genotype, the sample at POSITION j of a variant in byte j/4,
 * bits 2*(j%4)..+1; codes 0/1/2 = dosage, 3 = missing.  stride_bytes >= ceil(n/4), multiple of 16;
 * bits beyond position n-1 must be 0.
 *
 * Sample order.  Host-side packed columns (staar_pack_dosage_host, staar_packed_score_host,
 * staar_individual_analysis_chunk) are in the CALLER's sample order (position j = sample j).  Columns RESIDENT in
 * HBM for the *_device entry points are in the "null-model sample order": staar_nullmodel_set_sparse* re-orders the
 * samples once so that every family of Sigma_i (connected component of its off-diagonal pattern) of up to 16
 * members lies inside one 16-sample word and related samples come first; staar_nullmodel_sample_order returns the
 * order (position -> caller's sample index; the identity for unrelated samples) and staar_packed_reorder_device
 * converts caller-order columns already on the device. */
size_t staar_packed_stride(int64_t n);

/* host FP64 dosage (n x p, NaN / <= -1 missing) -> packed host buffer (p x stride).  Returns
 * STAAR_ERR_ARG if a non-missing value is not exactly 0, 1 or 2 (such matrices take the FP64 path). */
int staar_pack_dosage_host(const double *G, int64_t n, int64_t p, uint8_t *packed, size_t stride_bytes, int threads);
/* same for an R integer matrix (NA_integer_ = INT_MIN, or any value <= -1, is missing) and for SeqArray's raw
 * dosage bytes (0xFF missing): what seqGetData("$dosage") holds BEFORE Rcpp widens it to FP64
 * (R/Individual_Analysis.R:186; the conversion happens at src/RcppExports.cpp:23). */
int staar_pack_dosage_host_i32(const int32_t *G, int64_t n, int64_t p, uint8_t *packed, size_t stride_bytes, int threads);
int staar_pack_dosage_host_u8(const uint8_t *G, int64_t n, int64_t p, uint8_t *packed, size_t stride_bytes, int threads);

/* Null model, uploaded once (host pointers) ... */
int staar_nullmodel_set_sparse(staar_ctx *ctx, int64_t n, int64_t q,
                               const double *Si_values, const uint64_t *Si_row_idx, const uint64_t *Si_col_ptr,
                               const double *Sigma_iX, const double *cov, const double *residuals);
/* ... or adopted from device memory the caller owns (e.g. torch tensors filled by an NCCL broadcast):
 * CSR of Sigma_i with int32 columns / int64 row pointers, all pointers device. */
int staar_nullmodel_set_sparse_device(staar_ctx *ctx, int64_t n, int64_t q,
                                      const int64_t *d_row_ptr, const int32_t *d_col, const double *d_val,
                                      const double *d_Sigma_iX, const double *d_cov, const double *d_residuals);

/* order[j] = caller's sample index stored at position j of a resident packed column (length n). */
int staar_nullmodel_sample_order(staar_ctx *ctx, int64_t n, int32_t *order);
/* caller-order packed columns -> null-model order (device pointers, d_out != d_in, same stride). */
int staar_packed_reorder_device(staar_ctx *ctx, const uint8_t *d_in, uint8_t *d_out, size_t stride_bytes, int64_t n,
                                int64_t p);

#define STAAR_IMPUTE_MEAN 0
#define STAAR_IMPUTE_MINOR 1

/* K1 alone: AF / MAF / flip flag / missing count per variant from packed columns (device pointers).
 * Bit-exact with matrix_flip_mean / matrix_flip_minor on the same dosages. */
int staar_packed_flip_stats_device(staar_ctx *ctx, const uint8_t *d_packed, size_t stride_bytes, int64_t n, int64_t p,
                                   int impute, double *d_AF, double *d_MAF, int32_t *d_flip, int32_t *d_nmiss);

/* The fused per-variant path on resident packed columns (null-model sample order): flip/impute +
 * Individual_Score_Test.  All pointers are DEVICE pointers; outputs are 7 arrays of length p.
 * Requires staar_nullmodel_set_sparse*. */
int staar_packed_score_device(staar_ctx *ctx, const uint8_t *d_packed, size_t stride_bytes, int64_t n, int64_t p,
                              int impute, double *d_AF, double *d_MAF, double *d_Score, double *d_Score_se,
                              double *d_pvalue_log, double *d_Est, double *d_Est_se);

/* One region of the single-variant driver on resident columns (R/Individual_Analysis.R:164-450 without the GDS I/O):
 * flip/impute + score test for all p variants, then the driver's filter on the device — a variant is kept when
 * MAF > (mac_cutoff - 0.5)/n/2 (:242, :324; mac_cutoff = 20 by default, :45) — and the kept rows are compacted in
 * variant order into the caller's DEVICE arrays (capacity p each): d_index (variant number within the block) and the
 * seven statistics.  If d_spa_index is not NULL it receives, for binary traits analysed with SPA, the rows OF THE
 * COMPACTED TABLE with exp(-pvalue_log) < spa_p_cutoff (:290-297; 0.05 in the reference).  Synchronous; the counts
 * come back in *n_tested / *n_spa. */
int staar_packed_analysis_device(staar_ctx *ctx, const uint8_t *d_packed, size_t stride_bytes, int64_t n, int64_t p,
                                 int impute, double mac_cutoff, double spa_p_cutoff, int64_t *d_index, double *d_AF,
                                 double *d_MAF, double *d_Score, double *d_Score_se, double *d_pvalue_log, double *d_Est,
                                 double *d_Est_se, int64_t *d_spa_index, int64_t *n_tested, int64_t *n_spa);

/* SPA stage on resident columns (configuration 5: imbalanced binary trait).  staar_nullmodel_set_spa uploads the SPA
 * operands once (XW q x n, XXWX_inv n x q, residuals, muhat: R/fit_nullmodel.R:153-162).  staar_packed_spa_device
 * then runs Individual_Score_Test_SPA (src/Individual_Score_Test_SPA.cpp:38-533) on the n_sel variants listed in
 * d_variant (device int64: column numbers of d_packed, e.g. d_index gathered at d_spa_index of
 * staar_packed_analysis_device): each column is flipped / imputed exactly as matrix_flip_mean|minor and decoded to the
 * FP64 column the reference would pass, inside the saddle-point kernel (no FP64 matrix is materialised).  d_pvalue: n_sel
 * p-values (device).  Synchronous. */
int staar_nullmodel_set_spa(staar_ctx *ctx, int64_t n, int64_t q, const double *XW, const double *XXWX_inv,
                            const double *residuals, const double *muhat);
int staar_packed_spa_device(staar_ctx *ctx, const uint8_t *d_packed, size_t stride_bytes, int64_t n,
                            const int64_t *d_variant, int64_t n_sel, int impute, double tol, int max_iter,
                            double *d_pvalue, int64_t *d_trace /* 4 per variant as staar_individual_score_test_SPA, or NULL */);

/* Dense-GRM score test on resident packed columns (configuration 3).  staar_nullmodel_set_dense uploads P (n x n,
 * column-major: Sigma^-1 - Sigma^-1 X cov X' Sigma^-1) and the residuals once.  staar_packed_score_dense_device then
 * computes, for p resident 2-bit columns in the CALLER's sample order (P's order), AF / MAF as matrix_flip_mean|minor
 * and the five statistics of Individual_Score_Test_denseGRM (src/Individual_Score_Test_denseGRM.cpp:9-61) into device
 * arrays of length p.  Variants with at most n/8 non-zero genotypes after flip and imputation take the carrier form of
 * g'Pg, the others are decoded in batches and go through P*G.  Synchronous. */
int staar_nullmodel_set_dense(staar_ctx *ctx, int64_t n, const double *P, const double *residuals);
int staar_packed_score_dense_device(staar_ctx *ctx, const uint8_t *d_packed, size_t stride_bytes, int64_t n, int64_t p,
                                    int impute, double *d_AF, double *d_MAF, double *d_Score, double *d_Score_se,
                                    double *d_pvalue_log, double *d_Est, double *d_Est_se);

/* Multi-trait score test on resident packed columns (configuration 4).  The null model is the (n k) x (n k) one of
 * staar_nullmodel_set_sparse (trait-major stacking, as GMMAT / R/Individual_Analysis.R:57-68 deliver it).  For p
 * resident 2-bit columns of G0 (n samples, caller's order) this computes AF / MAF as matrix_flip_mean|minor and what
 * Individual_Score_Test_sp_multi (src/Individual_Score_Test_sp_multi.cpp:9-68) returns for Diagonal(k) (x) G0:
 * d_Score (k p values: trait-major, index t p + i) and d_pvalue_log (p values).  Synchronous. */
int staar_packed_score_multi_device(staar_ctx *ctx, const uint8_t *d_packed, size_t stride_bytes, int64_t n, int64_t p,
                                    int n_pheno, int impute, double *d_AF, double *d_MAF, double *d_Score,
                                    double *d_pvalue_log);

/* Conditional test on resident columns (configuration 5, R/Individual_Analysis_cond.R:183-208).  The null model set by
 * staar_nullmodel_set_sparse carries the REFIT residuals (residuals of lm(residuals ~ X_adj), :190-197); X_adj is the
 * host n x m adjustment matrix (its derived operands stay on the device between calls with the same X_adj).  The n_sel
 * columns listed in d_variant (device int64: column numbers of the resident block) are flipped / imputed as
 * matrix_flip_mean|minor, decoded on the device and go through Individual_Score_Test_cond
 * (src/Individual_Score_Test_cond.cpp:10-75).  Outputs: host arrays of length n_sel.  Synchronous. */
int staar_packed_cond(staar_ctx *ctx, const uint8_t *d_packed, size_t stride_bytes, int64_t n, const int64_t *d_variant,
                      int64_t n_sel, int impute, const double *X_adj, int64_t m, double *Score, double *Score_se,
                      double *pvalue_log, double *Est, double *Est_se);

/* Conditional analysis of ONE GROUP of tested variants that share their known loci, refit included (SURVEY §8f f3;
 * R/Individual_Analysis_cond.R:183-212 does, per tested variant, lm(scaled.residuals ~ genotype_adj + X - 1)
 * ["optimal", method = 0] or lm(scaled.residuals ~ genotype_adj) [intercept, method = 1], takes the fit's residuals and
 * X_adj = model.matrix, and calls Individual_Score_Test_cond with p = 1).  Here the known variants (d_known, indices of
 * resident columns) are decoded, flipped and imputed on the device into X_adj, the least-squares refit runs on the
 * device (Gram contraction, m x m Cholesky of the column-scaled normal equations on the host, two rounds of iterative
 * refinement), and all n_tested variants (d_tested) are tested in one call.  The null model set by
 * staar_nullmodel_set_sparse* must hold the ORIGINAL scaled residuals.  X: host n x qx covariates (method 0; ignored
 * for method 1).  coef_out (optional): the m = n_known + qx (or 1 + n_known) fitted coefficients.  The caller groups
 * tested variants by identical known-loci sets (the +-1 Mb windows of R/Individual_Analysis_cond.R:185). */
int staar_packed_cond_refit(staar_ctx *ctx, const uint8_t *d_packed, size_t stride_bytes, int64_t n, const int64_t *d_tested,
                            int64_t n_tested, const int64_t *d_known, int64_t n_known, const double *X, int64_t qx, int method,
                            int impute, double *Score, double *Score_se, double *pvalue_log, double *Est, double *Est_se,
                            double *coef_out);

/* Per-kernel device times of the LAST staar_packed_score_device call, measured with CUDA events on the context
 * stream when profiling is on: ms[0] = scan (K1+K3), ms[1] = contraction (K2), ms[2] = finalize (K5).
 * Used by bench.py for the live roofline figure; off by default (no events recorded). */
int staar_ctx_set_profiling(staar_ctx *ctx, int on);
int staar_ctx_last_stage_ms(staar_ctx *ctx, float ms[3]);

/* Route of staar_packed_score_device (and everything built on it: the chunk driver, staar_packed_analysis_device):
 *   STAAR_PACKED_PATH_FUSED — one pass over the genotypes: tcgen05 contraction and the sparse terms from the
 *     same shared-memory tiles (fused_tc.cu), warp-per-variant finaliser, block-per-variant redo of the few variants
 *     whose streaming orientation guess was wrong or whose missing list overflowed.  Null models with a family of
 *     more than 16 members fall back to the split route.  With profiling on, staar_ctx_last_stage_ms then reports
 *     ms[0] = redo, ms[1] = fused kernel, ms[2] = finaliser.
 *   STAAR_PACKED_PATH_SPLIT (default) — contraction kernel, then block-per-variant scan, then finaliser (two passes
 *     over G; on B200 the faster of the two).  The contraction kernel also lists the missing samples of
 *     every variant, so the scan knows the flip up front and runs without a detection pass (per-call scratch: 3.9 KB per
 *     variant at n = 1e5; variants with more missing samples than the lists hold, and calls whose lists do not fit on
 *     the device, go through the scan's own detection pass).  With profiling on: ms[0] = scan stage (list expansion +
 *     both scan launches), ms[1] = contraction, ms[2] = finaliser.
 * Both routes compute the same statistics (src/Individual_Score_Test.cpp:42-64); environment overrides at context
 * creation: STAAR_B200_PACKED_PATH=fused|split; STAAR_B200_MISSING_LISTS=0 (split route without the lists);
 * STAAR_B200_CONTRACT=pair|mma (CTA-pair tcgen05 kernel / legacy mma.sync kernel instead of the default tcgen05 one). */
#define STAAR_PACKED_PATH_FUSED 0
#define STAAR_PACKED_PATH_SPLIT 1
int staar_ctx_set_packed_path(staar_ctx *ctx, int path);
/* 1 when the last staar_packed_score_device call took the fused route, 0 otherwise */
int staar_ctx_last_packed_path_was_fused(staar_ctx *ctx);
/* Test hook, fused route: per-variant scanner outputs of the LAST call (any pointer may be NULL): acc (2 x gathered
 * hom-minor diagonal + kinship tables), cnt (hom-minor carriers in the uniform-diagonal region), cm (missing
 * samples), guess (orientation assumed while streaming); redo_count = variants sent to the block-per-variant scan. */
int staar_packed_debug_fused(staar_ctx *ctx, int64_t p, double *acc, int32_t *cnt, int32_t *cm, int32_t *guess,
                             uint64_t *redo_count);
/* Test hook, fused route with STAAR_B200_FUSED_DBG & 128: clocks per phase, 32 warps x 8 counters per CTA. */
int staar_packed_debug_timers(staar_ctx *ctx, int64_t ctas, uint32_t *out);

/* Test hook: intermediates of the LAST staar_packed_score_device call, copied to host arrays (any may be NULL):
 * hi/lo (p x 16 int64: exact integer contraction of the raw codes, X = hi*2^24 + lo in units of col_scale[c]),
 * Sm (p x 16 sums over missing samples; slot 15 = sum of Sigma_i[j,j] over hom-minor carriers), quad (kinship off-diagonal part), mu, flip,
 * col_scale (16). */
int staar_packed_debug_dump(staar_ctx *ctx, int64_t p, long long *hi, long long *lo, double *Sm, double *quad,
                            double *mu, int32_t *flip, double *col_scale);

/* Same, host buffers (packed host columns in, 7 host arrays out); H2D/D2H inside. */
int staar_packed_score_host(staar_ctx *ctx, const uint8_t *packed, size_t stride_bytes, int64_t n, int64_t p,
                            int impute, double *AF, double *MAF, double *Score, double *Score_se,
                            double *pvalue_log, double *Est, double *Est_se);

/* One chunk of the single-variant driver (R/Individual_Analysis.R:186-196, 264): an FP64 `$dosage` block in HOST
 * memory (n x p, NaN / <= -1 missing) -> flip/impute + Individual_Score_Test statistics (7 host arrays of length p)
 * in one call, without materialising Geno on the host.  0/1/2/NA blocks are packed to 2 bits by `threads` host
 * threads (0 = all cores) so that 1/32 of the bytes cross PCIe; real-valued blocks take the FP64 kernels.
 * Requires staar_nullmodel_set_sparse*. */
int staar_individual_analysis_chunk(staar_ctx *ctx, const double *G, int64_t n, int64_t p, int impute, int threads,
                                    double *AF, double *MAF, double *Score, double *Score_se, double *pvalue_log,
                                    double *Est, double *Est_se);
/* When the FP64 block lies in pinned (device-visible) host memory, this share of its variants is packed by the GPU,
 * which reads the host matrix over PCIe, while the host threads pack the rest (the two bandwidths add up). */
/* Pipelined form of the three calls above.  staar_chunk_submit* packs the block (host threads, and the GPU for its
 * share of a pinned FP64 block), enqueues the H2D copy and the kernels on the context's stream and returns;
 * staar_chunk_wait blocks until that chunk is done and copies its seven result arrays out (any may be NULL).  Two
 * chunks can be in flight (slot 0 and 1): submit(0); submit(1); wait(0); submit(0); wait(1); ... — while the GPU copies
 * and scores one chunk the host already packs the next.  G must stay valid and unchanged until the slot's wait()
 * returns.  Results are bit-identical to the synchronous calls; a block with real-valued genotypes is detected at
 * wait() and re-run through the FP64 kernels there. */
int staar_chunk_submit(staar_ctx *ctx, const double *G, int64_t n, int64_t p, int impute, int threads, int slot);
int staar_chunk_submit_i32(staar_ctx *ctx, const int32_t *G, int64_t n, int64_t p, int impute, int threads, int slot);
int staar_chunk_submit_u8(staar_ctx *ctx, const uint8_t *G, int64_t n, int64_t p, int impute, int threads, int slot);
int staar_chunk_wait(staar_ctx *ctx, int slot, double *AF, double *MAF, double *Score, double *Score_se,
                     double *pvalue_log, double *Est, double *Est_se);

double staar_chunk_gpu_pack_share(void);
/* the share the next FP64 chunk call on this context will use (re-balanced after every chunk unless fixed by the
 * environment variable) */
double staar_ctx_gpu_pack_share(staar_ctx *ctx);
/* the same chunk from the R integer matrix / raw bytes SeqArray returns (4 or 1 bytes per genotype on the host) */
int staar_individual_analysis_chunk_i32(staar_ctx *ctx, const int32_t *G, int64_t n, int64_t p, int impute, int threads,
                                        double *AF, double *MAF, double *Score, double *Score_se, double *pvalue_log,
                                        double *Est, double *Est_se);
int staar_individual_analysis_chunk_u8(staar_ctx *ctx, const uint8_t *G, int64_t n, int64_t p, int impute, int threads,
                                       double *AF, double *MAF, double *Score, double *Score_se, double *pvalue_log,
                                       double *Est, double *Est_se);

/* Synthetic packed dosage columns generated on the device (bench / parity subsets); per-variant
 * thresholds as in staarpipeline_b200/synth.py (device pointers, uint64 / uint8).  d_sample_order (device int32[n],
 * may be NULL = identity): the sample whose genotype is written at each position, so a shard can be generated
 * directly in the null-model order. */
int staar_synth_packed_device(staar_ctx *ctx, uint8_t *d_packed, size_t stride_bytes, int64_t n,
                              int64_t first_variant, int64_t p, uint64_t seed,
                              const uint64_t *d_thr_hom, const uint64_t *d_thr_het, const uint64_t *d_thr_miss,
                              const uint8_t *d_store_alt, const int32_t *d_sample_order);

/* ======================================================================= */
/* C. Several GPUs of one box behind one handle.    */
/* ======================================================================= */
/* One process, one staar_ctx per listed device, one host thread per device inside every call.  The variants (columns)
 * of a call are sharded contiguously across the devices; there is no data-path collective, every variant is
 * independent given the null model (R/Individual_Analysis.R:164-450 loops over variant chunks).  The sparse null
 * model is built ONCE — the host pass (families, sample order, kinship tables, digit slicing) runs for the first
 * device — and replicated to the others by device-to-device copies.  Results are bit-identical to the one-GPU calls:
 * a variant's arithmetic does not depend on which device or which other variants share the call.
 * The Rcpp shim opens such a handle when the environment variable STAAR_B200_DEVICES lists more than one device
 * (e.g. "0,1,2,3"), and routes matrix_flip_* and the dense-G Individual_Score_Test through it. */
typedef struct staar_multi staar_multi;
int staar_multi_create(const int *devices, int n_devices, staar_multi **out);
void staar_multi_destroy(staar_multi *m);
int staar_multi_size(const staar_multi *m);
staar_ctx *staar_multi_ctx(staar_multi *m, int i);            /* the per-device context (resident *_device entry points) */
int staar_multi_nullmodel_set_sparse(staar_multi *m, int64_t n, int64_t q, const double *Si_values,
                                     const uint64_t *Si_row_idx, const uint64_t *Si_col_ptr, const double *Sigma_iX,
                                     const double *cov, const double *residuals);
/* staar_individual_analysis_chunk over all devices (host FP64 n x p dosage block, seven arrays of length p out) */
int staar_multi_individual_analysis_chunk(staar_multi *m, const double *G, int64_t n, int64_t p, int impute, int threads,
                                          double *AF, double *MAF, double *Score, double *Score_se, double *pvalue_log,
                                          double *Est, double *Est_se);
/* frozen signatures, columns sharded: src/matrix_flip_mean.cpp:8-74, matrix_flip_minor.cpp:8-99,
 * Individual_Score_Test.cpp:9-67 (each device moves its own columns over its own PCIe link) */
int staar_multi_matrix_flip_mean(staar_multi *m, const double *G, int64_t n, int64_t p, double *Geno_out, double *AF, double *MAF);
int staar_multi_matrix_flip_minor(staar_multi *m, const double *G, int64_t n, int64_t p, double *Geno_out, double *AF, double *MAF);
int staar_multi_individual_score_test(staar_multi *m, const double *G, int64_t n, int64_t p,
                                      const double *Si_values, const uint64_t *Si_row_idx, const uint64_t *Si_col_ptr,
                                      const double *Sigma_iX, int64_t q, const double *cov, const double *residuals,
                                      double *Score, double *Score_se, double *pvalue_log, double *Est, double *Est_se);

#ifdef __cplusplus
}
#endif
#endif /* STAAR_B200_H */
