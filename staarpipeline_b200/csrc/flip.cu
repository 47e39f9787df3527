// flip.cu — K1: allele frequency, missing-genotype imputation and flip to the minor allele.
//
// Replaces matrix_flip_mean (reference src/matrix_flip_mean.cpp:8-74) and matrix_flip_minor
// (src/matrix_flip_minor.cpp:8-99).  Two forms:
//   * flip_dense_kernel  — FP64 n x p genotype matrix (what the Rcpp boundary delivers);
//   * flip_stats_packed  — 2-bit packed dosage columns: integer counts only, so AF / MAF / flip
//     decisions are bit-exact (AF = sum/2/num is one exact halving and one IEEE division).
// HBM-bound byte/integer work: one CTA (dense) or one warp (packed) per variant column,
// coalesced 128-bit loads, popcount arithmetic, no tensor cores.
#include "common.cuh"
#include "packed.cuh"

namespace staar {

// ---------------------------------------------------------------- dense FP64
// One CTA per column.  Pass 1: sum and count of entries with G > -1 (matrix_flip_mean.cpp:23-41).
// Pass 2: impute, recompute AF for `minor` (matrix_flip_minor.cpp:68-78), flip (…:56-70 / 81-95).
// For 0/1/2 dosages every partial sum is an exact integer, so the tree reduction reproduces the
// reference's in-order sum bit for bit; for real-valued G only 1e-10 parity is possible.
__global__ void __launch_bounds__(256) flip_dense_kernel(double *__restrict__ G, int64_t n, int64_t p, int minor,
                                                         int write_geno, double *__restrict__ AF,
                                                         double *__restrict__ MAF)
{
    __shared__ double red[32];
    for (int64_t i = blockIdx.x; i < p; i += gridDim.x) {
        double *g = G + i * n;
        double s = 0.0, c = 0.0;
        for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
            double v = g[j];
            if (v > -1.0) { s += v; c += 1.0; }
        }
        s = block_sum(s, red);
        c = block_sum(c, red);
        double af = s;
        if (c > 0.0) af = s / 2.0 / c;
        double fill;
        if (!minor) {
            fill = af * 2.0;
        } else {
            fill = (af <= 0.5) ? 0.0 : 2.0;
            // recomputed AF over all n entries after imputation
            // (sum over the imputed column) = s + fill * #missing: exact for integer dosages
            af = (s + fill * ((double)n - c)) / 2.0 / (double)n;
        }
        const bool flip = !(af <= 0.5);
        if (threadIdx.x == 0) {
            AF[i] = af;
            MAF[i] = flip ? 1.0 - af : af;
        }
        if (write_geno) {
            for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
                double v = g[j];
                if (!(v > -1.0)) v = fill;
                if (flip) v = 2.0 - v;
                g[j] = v;
            }
        }
        __syncthreads();
    }
}

int launch_flip_dense(staar_ctx *ctx, double *dG, int64_t n, int64_t p, bool minor, bool write_geno, double *dAF,
                      double *dMAF)
{
    if (p == 0) return STAAR_OK;
    int grid = (int)std::min<int64_t>(p, (int64_t)ctx->sm_count * 8);
    flip_dense_kernel<<<grid, 256, 0, ctx->stream>>>(dG, n, p, minor ? 1 : 0, write_geno ? 1 : 0, dAF, dMAF);
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

// ---------------------------------------------------------------- packed 2-bit
__global__ void __launch_bounds__(256) flip_stats_packed_kernel(const uint8_t *__restrict__ packed, size_t stride,
                                                                int64_t n, int64_t p, int impute,
                                                                double *__restrict__ AF, double *__restrict__ MAF,
                                                                int32_t *__restrict__ flip, int32_t *__restrict__ nmiss)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < p; i += nwarps) {
        PackedCounts pc = count_packed_column(packed + (size_t)i * stride, stride, lane);
        if (lane == 0) {
            VariantFlip vf = flip_from_counts(pc, n, impute);
            AF[i] = vf.af;
            MAF[i] = vf.maf;
            if (flip) flip[i] = vf.flip;
            if (nmiss) nmiss[i] = pc.cm;
        }
    }
}

int launch_flip_stats_packed(staar_ctx *ctx, const uint8_t *d_packed, size_t stride, int64_t n, int64_t p, int impute,
                             double *dAF, double *dMAF, int32_t *dflip, int32_t *dnmiss)
{
    if (p == 0) return STAAR_OK;
    int64_t blocks = std::min<int64_t>((p + 7) / 8, (int64_t)ctx->sm_count * 8);
    flip_stats_packed_kernel<<<(int)blocks, 256, 0, ctx->stream>>>(d_packed, stride, n, p, impute, dAF, dMAF, dflip,
                                                                    dnmiss);
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

}  // namespace staar
