// scan_packed.cu — K1 + K3 for packed dosages, and the per-variant finalisation (K5) of the packed path.
//
// scan_packed_kernel (one warp per variant, two sweeps over the 2-bit column):
//   sweep 1: popcount the codes -> AF, MAF, flip decision, imputation value — bit-exact with
//            matrix_flip_mean / matrix_flip_minor (src/matrix_flip_mean.cpp:23-70, src/matrix_flip_minor.cpp:23-95);
//   sweep 2 (the column is L1/L2-hot): flip the codes in the bit domain and visit only the NON-ZERO genotypes
//            (carriers and imputed entries, a few % of a rare variant's column) to accumulate the sparse quadratic
//            form g' Sigma_i g  — diagonal term s_jj g_j^2 plus the off-diagonal kinship rows of related samples
//            (the (trans(G)*Sigma_i)*G factor of src/Individual_Score_Test.cpp:43, diagonal entry only) — and the
//            column sums over missing entries that the integer contraction (contract_packed.cu) leaves out.
// finalize_packed_kernel (one thread per variant): integer contraction -> FP64, imputation correction mu * Sm,
//   Cov_ii = quad - t' cov t, and the statistics loop of src/Individual_Score_Test.cpp:45-64.
//
// Integer/byte work, HBM/L2-bound: 128-bit coalesced loads, popc/ffs bit tricks, warp-shuffle reductions with a
// fixed tree (no atomics => 1/2/4/8-GPU runs are bit-identical).
#include "common.cuh"
#include "packed.cuh"

namespace staar {

namespace {

struct ScanParams {
    const uint8_t *packed; size_t stride; int64_t n, p; int impute;
    const double *diag;          // Sigma_i[j,j]
    const int64_t *row_ptr; const int32_t *col; const double *val;   // CSR (full rows)
    const uint32_t *has_off;     // bit j set if row j has off-diagonal entries (nullptr: Sigma_i diagonal)
    const double *Y; int ldY, ncol;
    double *AF, *MAF; int32_t *flip; double *mu, *quad, *Sm;
};

__device__ __forceinline__ double code_value(uint32_t code, double mu)
{
    return code == 3u ? mu : (double)code;
}

// decoded (post-flip) genotype of sample k in a packed column
__device__ __forceinline__ double lookup_value(const uint8_t *colp, int64_t k, bool flip, double mu)
{
    uint32_t code = (colp[k >> 2] >> (2 * (k & 3))) & 3u;
    if (flip && code != 3u && code != 1u) code ^= 2u;
    return code_value(code, mu);
}

__global__ void __launch_bounds__(256) scan_packed_kernel(ScanParams P)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int nvec = (int)(P.stride >> 4);
    for (int64_t i = warp; i < P.p; i += nwarps) {
        const uint8_t *colp = P.packed + (size_t)i * P.stride;
        // ---- sweep 1: counts -> flip decision
        const PackedCounts pc = count_packed_column(colp, P.stride, lane);
        const VariantFlip vf = flip_from_counts(pc, P.n, P.impute);
        const bool flip = vf.flip != 0;
        const double mu = vf.mu;
        // ---- sweep 2: non-zero genotypes only
        double quad = 0.0;
        double sm[kColsPerSlice];
#pragma unroll
        for (int c = 0; c < kColsPerSlice; c++) sm[c] = 0.0;
        const uint4 *vecs = reinterpret_cast<const uint4 *>(colp);
        for (int k = lane; k < nvec; k += 32) {
            const uint4 q4 = __ldg(vecs + k);
            const uint32_t words[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
            for (int wi = 0; wi < 4; wi++) {
                uint32_t w = words[wi];
                const int64_t base = ((int64_t)k * 4 + wi) * 16;     // first sample of this word
                if (flip) {
                    w = flip_word(w);
                    const int64_t valid = P.n - base;                // padding codes would flip to 2
                    if (valid < 16) w &= (valid <= 0) ? 0u : ((1u << (2 * valid)) - 1u);
                }
                while (w) {
                    const int f = (__ffs(w) - 1) >> 1;
                    const uint32_t code = (w >> (2 * f)) & 3u;
                    w &= ~(3u << (2 * f));
                    const int64_t j = base + f;
                    const double v = code_value(code, mu);
                    if (code == 3u) {
                        const double *yr = P.Y + j * P.ldY;
#pragma unroll
                        for (int c = 0; c < kColsPerSlice; c++) if (c < P.ncol) sm[c] += yr[c];
                        if (v == 0.0) continue;                       // imputed to 0: contributes nothing more
                    }
                    quad += P.diag[j] * v * v;
                    if (P.has_off && (P.has_off[j >> 5] >> (j & 31)) & 1u) {
                        double s = 0.0;
                        const int64_t e1 = P.row_ptr[j + 1];
                        for (int64_t e = P.row_ptr[j]; e < e1; e++) {
                            const int32_t kcol = P.col[e];
                            if (kcol != j) s += P.val[e] * lookup_value(colp, kcol, flip, mu);
                        }
                        quad += v * s;
                    }
                }
            }
        }
        quad = warp_sum(quad);
        if (pc.cm > 0) {
#pragma unroll
            for (int c = 0; c < kColsPerSlice; c++) sm[c] = warp_sum(sm[c]);
        }
        if (lane == 0) {
            P.AF[i] = vf.af; P.MAF[i] = vf.maf; P.flip[i] = vf.flip; P.mu[i] = mu; P.quad[i] = quad;
        }
        if (lane < kColsPerSlice) {
            double out = 0.0;
#pragma unroll
            for (int c = 0; c < kColsPerSlice; c++) if (c == lane) out = sm[c];
            P.Sm[i * kColsPerSlice + lane] = out;
        }
    }
}

struct FinParams {
    int64_t p; int q;
    const double *mu, *quad, *Sm;
    const long long *hi, *lo;
    const double *scale, *cov;
    double *Score, *Score_se, *pl, *Est, *Est_se;
};

__global__ void __launch_bounds__(128) finalize_packed_kernel(FinParams P)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.p; i += (int64_t)gridDim.x * blockDim.x) {
        const double mu = P.mu[i];
        double L[kColsPerSlice];
        for (int c = 0; c <= P.q; c++) {
            // the integer contraction saw the (already flipped) codes with missing as 0; add mu * sum_missing(y)
            const long long hi = P.hi[i * 16 + c], lo = P.lo[i * 16 + c];
            L[c] = ((double)hi * 16777216.0 + (double)lo) * P.scale[c] + mu * P.Sm[i * kColsPerSlice + c];
        }
        const double U = L[0];
        double tct = 0.0;
        for (int b = 0; b < P.q; b++) {
            double xb = 0.0;
            for (int a = 0; a < P.q; a++) xb += L[1 + a] * P.cov[a + b * P.q];
            tct += xb * L[1 + b];
        }
        const double C = P.quad[i] - tct;
        P.Score[i] = U;
        single_trait_stats(U, C, P.Score_se[i], P.pl[i], P.Est[i], P.Est_se[i]);
    }
}

}  // namespace

int launch_scan_packed(staar_ctx *ctx, const uint8_t *d_packed, size_t stride, int64_t n, int64_t p, int impute,
                       double *dAF, double *dMAF, int32_t *dflip, double *d_mu, double *d_quad, double *d_Sm)
{
    if (p == 0) return STAAR_OK;
    const NullSparse &ns = ctx->ns;
    ScanParams P;
    P.packed = d_packed; P.stride = stride; P.n = n; P.p = p; P.impute = impute;
    P.diag = ns.diag; P.row_ptr = ns.row_ptr; P.col = ns.col; P.val = ns.val;
    P.has_off = ns.nnz_offdiag > 0 ? ns.has_off : nullptr;
    P.Y = ns.Y; P.ldY = ns.ldY; P.ncol = (int)ns.q + 1;
    P.AF = dAF; P.MAF = dMAF; P.flip = dflip; P.mu = d_mu; P.quad = d_quad; P.Sm = d_Sm;
    int64_t blocks = std::min<int64_t>((p + 7) / 8, (int64_t)ctx->sm_count * 8);
    scan_packed_kernel<<<(int)blocks, 256, 0, ctx->stream>>>(P);
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

int launch_finalize_packed(staar_ctx *ctx, int64_t p, const double *d_mu, const double *d_quad,
                           const double *d_Sm, const long long *d_hi, const long long *d_lo, double *dScore,
                           double *dScore_se, double *dpl, double *dEst, double *dEst_se)
{
    if (p == 0) return STAAR_OK;
    const NullSparse &ns = ctx->ns;
    FinParams P;
    P.p = p; P.q = (int)ns.q; P.mu = d_mu; P.quad = d_quad; P.Sm = d_Sm;
    P.hi = d_hi; P.lo = d_lo; P.scale = ns.col_scale; P.cov = ns.cov;
    P.Score = dScore; P.Score_se = dScore_se; P.pl = dpl; P.Est = dEst; P.Est_se = dEst_se;
    int blocks = (int)std::min<int64_t>((p + 127) / 128, (int64_t)ctx->sm_count * 8);
    finalize_packed_kernel<<<blocks, 128, 0, ctx->stream>>>(P);
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

}  // namespace staar
