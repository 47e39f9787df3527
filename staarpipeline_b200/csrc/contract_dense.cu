// contract_dense.cu — K2 for real-valued genotypes: L = G' * Y as an FP64 tensor-core contraction.
//
// G is the FP64 n x p matrix the Rcpp boundary delivers (column-major; also the AI-weighted
// real-valued genotypes of R/AI_Individual_Analysis.R:190-206), Y = [residuals | Sigma_iX | ...] is
// n x ncols row-major.  This is the first factor of every Cov in the reference:
//   Uscore = trans(residuals)*G, tSigma_iX_G = trans(Sigma_iX)*G   (src/Individual_Score_Test.cpp:17,42)
// M = variants, N = columns (skinny), K = samples, on mma.sync m8n8k4 f64 (DMMA; tcgen05 has no f64
// kind).  Each genotype is read from HBM exactly once (8 B/genotype => HBM-bound: 4 flop/B vs a
// ridge of 5.7); the Y tile is staged in shared memory and shared by the CTA's warps.
// Split-K over samples fills the 148 SMs when p is small; partials are reduced in a fixed order.
#include "common.cuh"

namespace staar {

namespace {

constexpr int kWarps = 4;          // warps per CTA, 8 variants each
constexpr int kKT = 64;            // samples per shared-memory Y tile
constexpr int kMaxNT = 8;          // n-tiles (8 columns each) per launch

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// grid = (variant tiles of 32, k-splits).  part: [ksplit][p][ldL]
template <int NT>
__global__ void __launch_bounds__(kWarps * 32) contract_dense_kernel(const double *__restrict__ G, int64_t n, int64_t p,
                                                                     const double *__restrict__ Y, int ldY, int col0,
                                                                     double *__restrict__ part, int ldL,
                                                                     int64_t k_per_split)
{
    constexpr int NC = NT * 8;
    constexpr int RS = NC + 1;                       // odd row stride: conflict-free B-fragment reads
    __shared__ double sY[kKT * RS];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int a = lane & 3, b = lane >> 2;
    const int64_t v = (int64_t)blockIdx.x * (kWarps * 8) + warp * 8 + b;   // this lane's variant (A row)
    const bool v_ok = v < p;
    const double *gcol = G + (v_ok ? v : 0) * n;
    const int64_t k_begin = (int64_t)blockIdx.y * k_per_split;
    const int64_t k_end = min(n, k_begin + k_per_split);

    double acc[NT][2];
#pragma unroll
    for (int t = 0; t < NT; t++) acc[t][0] = acc[t][1] = 0.0;

    for (int64_t k0 = k_begin; k0 < k_end; k0 += kKT) {
        // A fragments for this tile: 16 samples per lane (k0 + 16*a' ... see mapping below), issued before
        // the barrier so the global loads overlap the Y-tile staging
        double av[kKT / 4];
#pragma unroll
        for (int s = 0; s < kKT / 4; s++) {
            const int64_t k = k0 + (int64_t)(s >> 2) * 16 + 4 * a + (s & 3);
            av[s] = (v_ok && k < k_end) ? gcol[k] : 0.0;
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < kKT * NC; idx += kWarps * 32) {
            const int r = idx / NC, c = idx - r * NC;
            const int64_t k = k0 + r;
            sY[r * RS + c] = (k < k_end) ? Y[k * ldY + col0 + c] : 0.0;
        }
        __syncthreads();
        // step s covers samples k0 + 16*(s/4) + {4*a + s%4 : a = 0..3}  (A col a <-> that sample)
#pragma unroll
        for (int s = 0; s < kKT / 4; s++) {
            const int r = (s >> 2) * 16 + 4 * a + (s & 3);
#pragma unroll
            for (int t = 0; t < NT; t++) dmma884(acc[t][0], acc[t][1], av[s], sY[r * RS + t * 8 + b]);
        }
    }
    // C fragment: row b (variant), columns 2a, 2a+1 of each n-tile
    if (v_ok) {
        double *dst = part + ((size_t)blockIdx.y * p + v) * ldL + col0;
#pragma unroll
        for (int t = 0; t < NT; t++) {
            dst[t * 8 + 2 * a] = acc[t][0];
            dst[t * 8 + 2 * a + 1] = acc[t][1];
        }
    }
}

// fixed-order reduction of the split-K partials: L[i][c] = sum_s part[s][i][c]
__global__ void reduce_splits_kernel(const double *__restrict__ part, int64_t count, int ksplit, double *__restrict__ L)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < ksplit; k++) s += part[(size_t)k * count + i];
        L[i] = s;
    }
}

template <int NT>
void launch_nt(staar_ctx *ctx, dim3 grid, const double *G, int64_t n, int64_t p, const double *Y, int ldY, int col0,
               double *part, int ldL, int64_t kps)
{
    contract_dense_kernel<NT><<<grid, kWarps * 32, 0, ctx->stream>>>(G, n, p, Y, ldY, col0, part, ldL, kps);
}

}  // namespace

int launch_contract_dense(staar_ctx *ctx, const double *dG, int64_t n, int64_t p, const double *dY, int ldY, int ncols,
                          double *dL, int ldL)
{
    if (p == 0 || ncols == 0) return STAAR_OK;
    if (ldY % 8 != 0 || ldL < ldY || ncols > ldY) {
        set_error("contract_dense: ldY must be a multiple of 8 and ldL >= ldY >= ncols");
        return STAAR_ERR_ARG;
    }
    const int64_t vtiles = (p + kWarps * 8 - 1) / (kWarps * 8);
    // enough CTAs for ~4 per SM; each split covers a multiple of kKT samples
    int64_t ksplit = std::max<int64_t>(1, (4LL * ctx->sm_count + vtiles - 1) / vtiles);
    const int64_t ktiles = (n + kKT - 1) / kKT;
    ksplit = std::min<int64_t>(ksplit, std::max<int64_t>(1, ktiles / 8));
    const int64_t kps = ((ktiles + ksplit - 1) / ksplit) * kKT;
    ksplit = (n + kps - 1) / kps;
    double *part = dL;
    if (ksplit > 1) {
        STAAR_TRY(ctx->tmp2.reserve(sizeof(double) * (size_t)ksplit * p * ldL));
        part = ctx->tmp2.as<double>();
    }
    const int total_nt = (ncols + 7) / 8;
    for (int t0 = 0; t0 < total_nt; t0 += kMaxNT) {
        const int nt = std::min(kMaxNT, total_nt - t0);
        dim3 grid((unsigned)vtiles, (unsigned)ksplit);
        switch (nt) {
            case 1: launch_nt<1>(ctx, grid, dG, n, p, dY, ldY, t0 * 8, part, ldL, kps); break;
            case 2: launch_nt<2>(ctx, grid, dG, n, p, dY, ldY, t0 * 8, part, ldL, kps); break;
            case 3: launch_nt<3>(ctx, grid, dG, n, p, dY, ldY, t0 * 8, part, ldL, kps); break;
            case 4: launch_nt<4>(ctx, grid, dG, n, p, dY, ldY, t0 * 8, part, ldL, kps); break;
            case 5: launch_nt<5>(ctx, grid, dG, n, p, dY, ldY, t0 * 8, part, ldL, kps); break;
            case 6: launch_nt<6>(ctx, grid, dG, n, p, dY, ldY, t0 * 8, part, ldL, kps); break;
            case 7: launch_nt<7>(ctx, grid, dG, n, p, dY, ldY, t0 * 8, part, ldL, kps); break;
            default: launch_nt<8>(ctx, grid, dG, n, p, dY, ldY, t0 * 8, part, ldL, kps); break;
        }
        ctx->launches++;
        STAAR_CUDA(cudaGetLastError());
    }
    if (ksplit > 1) {
        const int64_t count = p * (int64_t)ldL;
        int blocks = (int)std::min<int64_t>((count + 255) / 256, (int64_t)ctx->sm_count * 8);
        reduce_splits_kernel<<<blocks, 256, 0, ctx->stream>>>(part, count, (int)ksplit, dL);
        ctx->launches++;
        STAAR_CUDA(cudaGetLastError());
    }
    return STAAR_OK;
}

}  // namespace staar
