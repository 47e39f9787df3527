// common.cuh — shared declarations for the sm_100a score-test kernels and the C ABI behind them.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cmath>
#include <vector>
#include <string>
#include "../../include/staar_b200.h"

namespace staar {

// ------------------------------------------------------------------ errors
void set_error(const char *fmt, ...);
#define STAAR_CUDA(x)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (x);                                                                 \
        if (e__ != cudaSuccess) {                                                              \
            staar::set_error("%s:%d: %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e__)); \
            return STAAR_ERR_CUDA;                                                             \
        }                                                                                      \
    } while (0)
#define STAAR_TRY(x)                   \
    do {                               \
        int rc__ = (x);                \
        if (rc__ != STAAR_OK) return rc__; \
    } while (0)

// ------------------------------------------------------------------ device buffers
// Grow-only device scratch owned by the context: chunk calls reuse it instead of cudaMalloc per call.
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
    template <typename T> T *as() { return reinterpret_cast<T *>(p); }
};

// Sliced (exact INT8) operand of the packed contraction: see contract_packed.cu.
constexpr int kSlices = 7;          // 7 balanced base-256 digits = 54-bit fixed point per column
constexpr int kColsPerSlice = 16;   // 1 + q <= 16 columns per slice group (r | Sigma_iX), padded
constexpr int kNT = kSlices * kColsPerSlice / 8;   // n-tiles of 8 columns (14)
constexpr int kChunkSamples = 256;  // samples per k-chunk (8 IMMA k32 steps)
constexpr int kStepBytesB = kNT * 32 * 8;             // mma.sync version, per IMMA k32 step: 14 column tiles (3584)
constexpr int kChunkBytesB = 8 * kStepBytesB;         // bytes of fragment-major B per k-chunk (28672)

struct NullSparse {
    bool set = false, owned = true;
    int64_t n = 0, q = 0;
    // CSR of Sigma_i (symmetric), full rows including the diagonal
    int64_t nnz = 0, nnz_offdiag = 0;
    int64_t *row_ptr = nullptr;   // n+1
    int32_t *col = nullptr;       // nnz
    double *val = nullptr;        // nnz
    double *diag = nullptr;       // n   (Sigma_i[j,j])
    struct alignas(16) SampleInfo { double diag; uint32_t off_start, off_count; };
    struct alignas(16) OffEntry { double val; int32_t col; int32_t pad; };
    // ---- packed path: "null-model sample order".  Samples are re-ordered once, at set-up, so that every family
    // (connected component of Sigma_i's off-diagonal graph) of 2..16 members sits inside ONE 16-sample word of the
    // packed column, families of <= 4 members inside one aligned 4-sample group; related samples come first,
    // unrelated ones after.  The kinship part of g' Sigma_i g then needs no genotype gathers: a lane takes the byte
    // of a group (four 2-bit codes) from the word it already holds and looks the group's contribution up in that
    // group's table (<= 256 entries); families of 5..16 members add member x later-chunk terms ("cross" words).
    // Families of more than 16 members take the "generic" record path below.
    bool permuted = false;            // false: perm is the identity
    std::vector<int32_t> h_perm;      // position -> original sample (empty when identity)
    int32_t *perm = nullptr;          // device copy (nullptr when identity)
    int32_t *inv_perm = nullptr;      // original sample -> position (nullptr when identity)
    double *Yp = nullptr;             // Y rows in null-model order (== Y when !permuted; then not owned)
    int64_t kin_words = 0;            // 16-sample words [0, kin_words) hold every sample with kinship off-diagonals
    int64_t tab_words = 0;            // words [0, tab_words): table words; [tab_words, kin_words): generic families
    int64_t x_words = 0;              // words [0, x_words) additionally carry cross terms (families of 5..16 members)
    uint32_t *kmask = nullptr;        // per table word (padded to a multiple of 4 words): both bits of every related slot
    // cross term: value = code(word >> jshift2) * lut[off + ((word >> shift2) & mask)]
    struct KinTerm { uint32_t off; uint8_t shift2, mask, jshift2, pad; };
    struct WordDesc { uint32_t start, count; };
    WordDesc *wdesc = nullptr;        // x_words descriptors
    KinTerm *xterms = nullptr;
    double *lut = nullptr;            // group tables at stride 256: lut[((4 word + group) << 8) + codes]; then cross-term tables
    int64_t lut_entries = 0;
    SampleInfo *sinfo_p = nullptr;    // full kinship rows in null-model positions (missing related samples)
    double *diag_p = nullptr;         // Sigma_i[j,j] in null-model positions (n + 64 entries, zero padded)
    double *lutf = nullptr;           // fused path: same offsets as lut; group tables indexed by all four codes, hom-minor diagonal folded in
    int64_t diag2_off = 0;            // lutf[diag2_off + pos] = 2 * Sigma_i[j,j] (n + 64 entries): one array serves both gathers of the fused path
    int64_t diag_uniform_vec = -1;  // from this 64-sample vector on every sample has the same Sigma_i[j,j] (-1: none)
    double diag_uniform_val = 0.0;
    OffEntry *off_ent_p = nullptr;
    // generic path (families of more than 16 members, or every family when the column does not fit in shared memory):
    // 2-bit mask (11 = sample has kinship off-diagonals with a LARGER position), one 32-byte record per sample with
    // its first two such entries inline, and the remaining ones in `up_ent`.  Positions are null-model positions.
    struct alignas(32) CarrierRec { double val[2]; int32_t col[2]; uint32_t start, count; };
    uint8_t *relmask = nullptr;   // packed_stride(n) bytes
    CarrierRec *rel = nullptr;    // n (only when some family is generic)
    OffEntry *up_ent = nullptr;   // upper-triangle off-diagonal entries, rows concatenated
    int64_t n_generic = 0;        // samples on the generic path
    double *SX = nullptr;         // Sigma_iX, n x q column-major (kept for the conditional test)
    // Y = [residuals | Sigma_iX | diag(Sigma_i)], row-major n x ldY (ldY = 8-padded 2+q), FP64.  The last column
    // rides along in the packed contraction (sum_j g_j s_jj); the FP64 kernels use the first 1+q columns only.
    int ldY = 0;
    double *Y = nullptr;
    double *cov = nullptr;        // q x q column-major
    // exact INT8 slices of Y in IMMA-fragment-major order, and their exact column sums
    bool sliced = false;
    int64_t n_chunks = 0;
    int8_t *Yfrag = nullptr;      // n_chunks * kChunkBytesB (mma.sync fragment-major order)
    int8_t *Ytc = nullptr;        // n_chunks * kChunkBytesB (tcgen05 canonical K-major core-matrix order)
    int8_t *Ypair = nullptr;      // the same digits split into two 56-row halves per k-block (CTA-pair contraction)
    double *col_scale = nullptr;  // kColsPerSlice: 2^(e_c - 54)
    long long *col_total = nullptr;   // 2 x kColsPerSlice: exact column sums of the fixed-point operand, hi | lo (X = hi 2^24 + lo)
    uint64_t fingerprint = 0;     // content hash of (Sigma_i, Sigma_iX, cov): cache key of the frozen-signature entry points
    uint64_t res_fingerprint = 0; // content hash of the residual vector currently in column 0 of Y
    // every device array above, as (offset of the pointer member in this struct, bytes): what a replica on another GPU
    // of the same process has to copy (multi.cu: the host-side build of the sample order and tables runs once)
    std::vector<std::pair<size_t, size_t>> arrays;
    void reg(void *member, size_t bytes)
    {
        const size_t off = (size_t)(static_cast<char *>(member) - reinterpret_cast<char *>(this));
        for (auto &a : arrays) if (a.first == off) { a.second = bytes; return; }
        arrays.emplace_back(off, bytes);
    }
    void release();
};

// SPA operands of a binary-trait null model (R/fit_nullmodel.R:153-162), resident for the packed path
struct NullSpa {
    bool set = false;
    int64_t n = 0, q = 0;
    double *XW = nullptr;         // q x n column-major
    double *XXWX_inv = nullptr;   // n x q column-major
    double *residuals = nullptr, *muhat = nullptr;
    void release();
};

struct NullDense {
    bool set = false, res_set = false;   // res_set: residuals supplied by staar_nullmodel_set_dense (resident path)
    int64_t n = 0;
    double *P = nullptr;          // n x n column-major
    double *residuals = nullptr;
    uint64_t fingerprint = 0;
    void release();
};

// per-variant output of the fused kernel's scanners (fused_tc.cu)
struct alignas(8) FusedScanOut {
    double acc;      // 2 * sum of Sigma_i[j,j] over gathered hom-minor carriers + kinship table look-ups (guessed orientation)
    int cnt;         // hom-minor carriers in the uniform-diagonal region
    int cm;          // missing samples
    int guess;       // orientation the scanners assumed (1 = flip)
    int pad;
};

}  // namespace staar

struct staar_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    uint64_t launches = 0;
    staar::NullSparse ns;
    staar::NullDense nd;
    staar::NullSpa nspa;
    // scratch
    staar::DevBuf G, L, aux, out, tmp, tmp2, pk, pk2, ctr;
    // operands of the conditional test derived from X_adj (PA, M, M Q M, Ycond), kept while X_adj and the null model
    // stay the same: the reference calls Individual_Score_Test_cond once per tested variant with the same X_adj
    staar::DevBuf cond;
    uint64_t cond_key = 0;
    int cond_m = 0, cond_ld = 0;
    bool use_mma_sync = false;       // STAAR_B200_CONTRACT=mma: legacy mma.sync contraction instead of tcgen05
    bool use_missing_lists = true;   // STAAR_B200_MISSING_LISTS=0: the scan detects missing samples itself (previous flow)
    bool use_tc_pair = false;        // STAAR_B200_CONTRACT=pair: tcgen05 cta_group::2 contraction (CTA pairs share the digit stream)
    bool attr_contract_pair = false;
    // packed score path: 1 = contraction + block-per-variant scan (two passes over G; default), 0 = fused single-pass
    // kernel where the null model allows it.  STAAR_B200_PACKED_PATH=split|fused, staar_ctx_set_packed_path.
    // fused_tiles: accumulator tiles per CTA of the fused kernel (2; 1 for experiments).
    int packed_path = 1, fused_tiles = 2;
    staar::DevBuf fscan, mlist, msubcnt, redo, plist, pcnt;
    bool last_fused = false;         // the last staar_packed_score_device call took the fused path
    // per-device launch attributes already set (cudaFuncSetAttribute is per device, the context owns one device)
    bool attr_contract_tc = false, attr_contract_packed = false, attr_permute = false;
    int spa_best_width = 0, spa_resident[17] = {0};     // SPA cluster width / resident clusters per width (spa.cu)
    void *pinned = nullptr;
    size_t pinned_cap = 0;
    // hybrid host/GPU packing of pinned FP64 chunks: share of the variants the GPU packs, re-balanced after every chunk
    // from the two measured packing rates (negative: not initialised; fixed when STAAR_B200_GPU_PACK_SHARE is set)
    double pack_share = -1.0;
    cudaEvent_t ev_pack[2] = {nullptr, nullptr};
    // pipelined chunk driver (staar_chunk_submit / staar_chunk_wait): per-slot host staging and result buffers
    struct ChunkSlot {
        void *pinned = nullptr; size_t pinned_cap = 0;
        staar::DevBuf res;                        // 7 p doubles + the GPU packer's bad-value flag
        cudaEvent_t done = nullptr, ev0 = nullptr, ev1 = nullptr;
        bool busy = false, host_ok = true;
        const void *G = nullptr; int elem = 0, impute = 0, threads = 0;
        int64_t n = 0, p = 0, p_cpu = 0, p_gpu = 0;
        double t_host = 0.0;
    } slot[2];
    cudaStream_t stream2 = nullptr;               // result read-back of a finished chunk while the next one is queued
    // staging ring for large transfers from / to pageable host memory (abi.cu:h2d / d2h)
    static constexpr int kStageBufs = 4;
    void *stage[kStageBufs] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t stage_ev[kStageBufs] = {nullptr, nullptr, nullptr, nullptr};
    void *stage_out[kStageBufs] = {nullptr, nullptr, nullptr, nullptr};       // second ring: device -> host while the first feeds the device
    cudaEvent_t stage_out_ev[kStageBufs] = {nullptr, nullptr, nullptr, nullptr}, stage_k_ev[kStageBufs] = {nullptr, nullptr, nullptr, nullptr};
    // optional per-kernel timing of the packed path (bench.py roofline)
    bool profiling = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
};

namespace staar {

int pinned_reserve(staar_ctx *ctx, size_t bytes);

// ------------------------------------------------------------------ kernel launch wrappers (one per .cu)
// flip.cu — K1
int launch_flip_dense(staar_ctx *ctx, double *dG, int64_t n, int64_t p, bool minor, bool write_geno,
                      double *dAF, double *dMAF);
int launch_flip_stats_packed(staar_ctx *ctx, const uint8_t *d_packed, size_t stride, int64_t n, int64_t p,
                             int impute, double *dAF, double *dMAF, int32_t *dflip, int32_t *dnmiss);
// contract_dense.cu — K2 (FP64 G): L[i, c] = sum_j G[j,i] * Y[j,c]   (L: p x ldL row-major)
int launch_contract_dense(staar_ctx *ctx, const double *dG, int64_t n, int64_t p, const double *dY, int ldY, int ncols,
                          double *dL, int ldL);
// quadform.cu — K3/K4
// out[pair] = g_a' * S * g_b for column pairs of a dense FP64 G and a CSR S (full rows)
int launch_quad_pairs_csr(staar_ctx *ctx, const double *dG, int64_t n, const int64_t *row_ptr, const int32_t *col,
                          const double *val, const int32_t *d_colA, const int32_t *d_colB, int64_t npairs, double *d_out);
// W = P * G (n x p), then out[pair] = g_a' * W[:, b]
int launch_dense_P_times_G(staar_ctx *ctx, const double *dP, const double *dG, int64_t n, int64_t p, double *dW);
int launch_quad_pairs_carriers_P(staar_ctx *ctx, const double *d_val, const int64_t *d_row, const int64_t *d_colptr, int64_t n,
                                 const double *dP, const int32_t *d_colA, const int32_t *d_colB, int64_t npairs, double *d_out);
int launch_contract_carriers(staar_ctx *ctx, const double *d_val, const int64_t *d_row, const int64_t *d_colptr, int64_t p,
                            const double *dY, int ldY, int ncols, double *dL, int ldL);
int launch_quad_pairs_carriers_csr(staar_ctx *ctx, const double *d_val, const int64_t *d_row, const int64_t *d_colptr,
                                   const int64_t *S_row_ptr, const int32_t *S_col, const double *S_val, const int32_t *d_colA,
                                   const int32_t *d_colB, int64_t npairs, double *d_out);
int launch_score_carriers(staar_ctx *ctx, const double *d_val, const int64_t *d_row, const int64_t *d_colptr, int64_t p,
                          const double *d_res, double *dL, int ldL);
int launch_dot_pairs(staar_ctx *ctx, const double *dG, const double *dW, int64_t n, const int32_t *d_colA,
                     const int32_t *d_colB, int64_t npairs, double *d_out);
// epilogue.cu — K5/K7.  k == 0: single-trait outputs; k >= 1: k-trait block test (pl has p_total/k entries).
// m > 0 adds the conditional terms (columns colA0.. = X_adj'G, colPA0.. = PX_adj'G of L).
int launch_epilogue_general(staar_ctx *ctx, int64_t p_total, int k, const double *dL, int ldL, int q,
                            const double *d_cov, const double *d_quad, int m, const double *dM, const double *dMQM,
                            int colA0, int colPA0, double *dScore, double *dScore_se, double *dpl, double *dEst,
                            double *dEst_se);
// spa.cu — K6
int launch_spa(staar_ctx *ctx, const double *dG, int64_t n, int64_t p, const double *dXW, const double *dXXWX_inv,
               int q, const double *d_res, const double *d_mu, double tol, int max_iter, double *d_gt,
               double *d_pvalue, long long *d_trace);
int launch_spa_packed(staar_ctx *ctx, const uint8_t *d_packed, size_t stride, const long long *d_variant,
                      const int32_t *d_inv_perm, int impute, int64_t n, int64_t p, const double *dXW,
                      const double *dXXWX_inv, int q, const double *d_res, const double *d_mu, double tol, int max_iter,
                      double *d_gt, double *d_pvalue, long long *d_trace);
// packed path — scan_packed.cu / contract_packed.cu / finalize
int launch_scan_packed(staar_ctx *ctx, const uint8_t *d_packed, size_t stride, int64_t n, int64_t p, int impute,
                       const long long *d_lo /* contraction output: slot 15 = sum of codes */, double *dAF, double *dMAF,
                       int32_t *dflip, double *d_mu, double *d_quad, double *d_Sm,
                       const long long *d_vlist = nullptr, const unsigned long long *d_vcount = nullptr,
                       const uint32_t *d_plist = nullptr, const uint32_t *d_pcnt = nullptr, int pcap = 0);
// segment records of the contraction (d_mlist / d_mcnt) -> positions d_plist[i * pcap ..] and counts d_pcnt[i] the scan's
// list mode reads; variants whose lists did not fit -> d_vlist / *d_vcount (work list of the detection path)
int launch_expand_lists(staar_ctx *ctx, const uint2 *d_mlist, const uint32_t *d_mcnt, int mcap, int64_t p, uint32_t *d_plist,
                        uint32_t *d_pcnt, int pcap, long long *d_vlist, unsigned long long *d_vcount);
int launch_contract_packed(staar_ctx *ctx, const uint8_t *d_packed, size_t stride, int64_t n, int64_t p,
                           long long *d_hi, long long *d_lo);
int launch_finalize_packed(staar_ctx *ctx, int64_t p, const int32_t *d_flip, const double *d_mu, const double *d_quad,
                           const double *d_Sm, const long long *d_hi, const long long *d_lo,
                           double *dScore, double *dScore_se, double *dpl, double *dEst, double *dEst_se,
                           const long long *d_vlist = nullptr, const unsigned long long *d_vcount = nullptr);
// fused_tc.cu — single pass: contraction + scanners, then the warp-per-variant finaliser
bool fused_path_available(const staar_ctx *ctx);
int fused_missing_sub(int64_t n);
int launch_fused_tc(staar_ctx *ctx, const uint8_t *d_packed, size_t stride, int64_t n, int64_t p, long long *d_hi,
                    long long *d_lo, FusedScanOut *d_scan, uint2 *d_mlist, int msub, uint16_t *d_msub_cnt);
int launch_fused_finalize(staar_ctx *ctx, int64_t n, int64_t p, int impute, const long long *d_hi, const long long *d_lo,
                          const FusedScanOut *d_scan, const uint2 *d_mlist, int msub, const uint16_t *d_msub_cnt,
                          double *dAF, double *dMAF,
                          double *dScore, double *dScore_se, double *dpl, double *dEst, double *dEst_se, long long *d_redo,
                          unsigned long long *d_redo_count);
int build_sliced_operand(staar_ctx *ctx, const double *hY /* n x ldY row-major host, natural sample order */);
int build_tc_operand(staar_ctx *ctx, const int8_t *digits, int ncol);
// d_mlist / d_mcnt (optional): per variant two sub-lists of mcap records {32-sample segment, mask of its missing samples}
// and per sub-list the number of records (above mcap = list truncated); expand_lists_kernel turns them into positions
int launch_contract_tc(staar_ctx *ctx, const uint8_t *d_packed, size_t stride, int64_t n, int64_t p, long long *d_hi,
                       long long *d_lo, uint2 *d_mlist = nullptr, uint32_t *d_mcnt = nullptr, int mcap = 0);
int refresh_permuted_Y(staar_ctx *ctx, const double *hY);
// ctx.cu / abi.cu — shared with multi.cu
int nullmodel_sparse_clone(staar_ctx *dst, const staar_ctx *src);
int ensure_sparse_shared(staar_ctx *ctx, int64_t n, int64_t q, const double *Sv, const uint64_t *Sr, const uint64_t *Sc,
                         const double *SX, const double *cov, const double *res);
// packed_dense.cu — dense-GRM score test on resident packed columns
int packed_score_dense(staar_ctx *ctx, const uint8_t *d_packed, size_t stride, int64_t n, int64_t p, int impute,
                       double *d_AF, double *d_MAF, double *d_Score, double *d_Score_se, double *d_pl, double *d_Est,
                       double *d_Est_se);
int launch_decode_selected(staar_ctx *ctx, const uint8_t *d_packed, size_t stride, int64_t n, const long long *d_variant,
                           int64_t count, int impute, const int32_t *d_perm, double *dG, double *dAF, double *dMAF);
int packed_score_multi(staar_ctx *ctx, const uint8_t *d_packed, size_t stride, int64_t n0, int64_t pp, int k, int impute,
                       double *d_AF, double *d_MAF, double *d_Score, double *d_pl);
// driver.cu — MAF/MAC filter + SPA pre-filter: stable compaction of the per-variant arrays
int launch_compact_results(staar_ctx *ctx, int64_t p, const double *const d_in[7], double maf_cut, double p_cut,
                           long long *d_index, double *const d_out[7], long long *d_sel_index, long long *d_totals);
// synth.cu
int launch_synth_packed(staar_ctx *ctx, uint8_t *d_packed, size_t stride, int64_t n, int64_t first_variant, int64_t p,
                        uint64_t seed, const uint64_t *thr_hom, const uint64_t *thr_het, const uint64_t *thr_miss,
                        const uint8_t *store_alt, const int32_t *sample_order);
// misc.cu
int launch_csc_to_dense(staar_ctx *ctx, const double *d_val, const int64_t *d_row, const int64_t *d_colptr,
                        int64_t n, int64_t p, double *dG);
int launch_build_Y(staar_ctx *ctx, int64_t n, int ldY, const double *d_res, const double *d_SX, int q, const double *dA,
                   const double *dPA, int m, double *dY);
int launch_set_Y_col0(staar_ctx *ctx, int64_t n, int ldY, const double *d_res, double *dY);
int launch_residual_update(staar_ctx *ctx, int64_t n, const double *dA, int m, const double *d_coef, const double *d_r, double *d_out);
int launch_px_adj(staar_ctx *ctx, int64_t n, const int64_t *row_ptr, const int32_t *col, const double *val,
                  const double *dA, int m, const double *dSX, int q, const double *dT1, double *dOut);
int launch_fill_pairs(staar_ctx *ctx, int64_t pp, int k, int32_t *d_colA, int32_t *d_colB);
int launch_pack_f64(staar_ctx *ctx, const double *G_dev_visible, int64_t n, int64_t pcount, uint8_t *d_packed, size_t stride,
                    int *d_bad);
int launch_permute_packed(staar_ctx *ctx, const uint8_t *d_in, uint8_t *d_out, size_t stride, int64_t n, int64_t p,
                          const int32_t *d_perm);
int spa_scratch_slots(staar_ctx *ctx, int64_t n, int64_t p);

// ------------------------------------------------------------------ device math shared by kernels
#ifdef __CUDACC__

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum with a fixed tree (deterministic: no atomics).  `red` = shared double[32].
__device__ __forceinline__ double block_sum(double v, double *red)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double t = (lane < nw) ? red[lane] : 0.0;
    t = warp_sum(t);
    return t;   // every thread holds the total
}

// log Q(a, y): regularised upper incomplete gamma in the log domain.  Same algorithm as the oracle's
// stand-in for R::pchisq(x, df, lower=false, log=true)  (reference src/Individual_Score_Test.cpp:57).
__device__ inline double log_gamma_q(double a, double y)
{
    if (isnan(a) || isnan(y)) return nan("");
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
    }
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

// R::pnorm(x, 0, 1, lower, false) with nmath's exact-0/1 cut-offs (reference _SPA.cpp:151,282)
__device__ inline double pnorm_std(double x, bool lower)
{
    if (isnan(x)) return nan("");
    if (lower) {
        if (x <= -37.5193) return 0.0;
        if (x >= 8.2924) return 1.0;
        return 0.5 * erfc(-x * 0.70710678118654752440);
    }
    if (x >= 37.5193) return 0.0;
    if (x <= -8.2924) return 1.0;
    return 0.5 * erfc(x * 0.70710678118654752440);
}

// The per-variant branch every single-trait test ends with (reference Individual_Score_Test.cpp:45-64)
__device__ inline void single_trait_stats(double U, double C, double &se, double &pl, double &est, double &est_se)
{
    if (C == 0.0) {
        se = 0.0; pl = 0.0; est = 0.0; est_se = 0.0;
    } else {
        double stat = (U * U) / C;
        pl = -log_gamma_q(0.5, 0.5 * stat);
        se = sqrt(C);
        est = U / C;
        est_se = 1.0 / se;
    }
}
#endif  // __CUDACC__

}  // namespace staar
