// misc.cu — small layout kernels around the hot path (operand assembly, sparse-genotype scatter,
// Sigma_i * X_adj for the conditional test).  All HBM-bound streaming work.
#include "common.cuh"

namespace staar {

namespace {

// scatter a CSC genotype matrix (what as(Geno,"dgCMatrix") hands to the _sp entry points,
// R/Individual_Analysis.R:343) into a zeroed dense n x p column-major matrix
__global__ void csc_scatter_kernel(const double *__restrict__ val, const int64_t *__restrict__ row,
                                   const int64_t *__restrict__ colptr, int64_t n, int64_t p, double *__restrict__ G)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = warp; c < p; c += nwarps) {
        const int64_t e1 = colptr[c + 1];
        // duplicate (row, column) entries add up, as in Armadillo's sp_mat; a column is owned by one warp, and within the
        // warp the atomic add makes entries of the same row in different lanes safe (entries are few: order effects
        // only arise for duplicates, which sorted dgCMatrix input never has)
        for (int64_t e = colptr[c] + lane; e < e1; e += 32) atomicAdd(&G[c * n + row[e]], val[e]);
    }
}

// Y (row-major n x ldY) <- columns gathered from up to four column-major blocks:
//   col 0 = residuals, cols 1..q = Sigma_iX, then X_adj (m cols) and PX_adj (m cols) for the conditional test.
__global__ void build_Y_kernel(int64_t n, int ldY, const double *__restrict__ res, const double *__restrict__ SX, int q,
                               const double *__restrict__ A, const double *__restrict__ PA, int m, double *__restrict__ Y)
{
    const int64_t total = n * (int64_t)ldY;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = idx / ldY;
        const int c = (int)(idx - j * ldY);
        double v = 0.0;
        if (c == 0) v = res ? res[j] : 0.0;
        else if (c <= q) v = SX[j + (int64_t)(c - 1) * n];
        else if (c <= q + m) v = A[j + (int64_t)(c - 1 - q) * n];
        else if (c <= q + 2 * m) v = PA ? PA[j + (int64_t)(c - 1 - q - m) * n] : 0.0;
        Y[idx] = v;
    }
}

// out = r - A coef  (A n x m column-major, m <= 64): the residuals of a least-squares fit (conditional refit)
__global__ void residual_update_kernel(int64_t n, const double *__restrict__ A, int m, const double *__restrict__ coef,
                                       const double *__restrict__ r, double *__restrict__ out)
{
    __shared__ double c[64];
    if (threadIdx.x < m) c[threadIdx.x] = coef[threadIdx.x];
    __syncthreads();
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int a = 0; a < m; a++) s += A[j + (int64_t)a * n] * c[a];
        out[j] = r[j] - s;
    }
}

// only column 0 (residuals) of an existing Y
__global__ void set_Y_col0_kernel(int64_t n, int ldY, const double *__restrict__ res, double *__restrict__ Y)
{
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x)
        Y[j * ldY] = res[j];
}

// out (n x m col-major) = S (CSR) * A (n x m col-major) - SX (n x q) * T1 (q x m)   [PX_adj, cond.cpp:48]
__global__ void px_adj_kernel(int64_t n, const int64_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                              const double *__restrict__ val, const double *__restrict__ A, int m,
                              const double *__restrict__ SX, int q, const double *__restrict__ T1,
                              double *__restrict__ out)
{
    const int64_t total = n * (int64_t)m;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = idx / n, j = idx - c * n;
        double s = 0.0;
        for (int64_t e = row_ptr[j]; e < row_ptr[j + 1]; e++) s += val[e] * A[col[e] + c * n];
        double t = 0.0;
        for (int a = 0; a < q; a++) t += SX[j + (int64_t)a * n] * T1[a + c * q];
        out[idx] = s - t;
    }
}

__global__ void fill_pairs_kernel(int64_t pp, int k, int32_t *__restrict__ colA, int32_t *__restrict__ colB)
{
    const int64_t total = pp * k * k;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = idx / (k * k);
        const int r = (int)(idx - i * k * k), a = r / k, b = r - a * k;
        colA[idx] = (int32_t)(a * pp + i);      // id_single(k) = k*pp + i  (Individual_Score_Test_multi.cpp:47-50)
        colB[idx] = (int32_t)(b * pp + i);
    }
}

// natural sample order -> null-model sample order for packed 2-bit columns: one CTA per variant, the source column is
// staged in shared memory and every thread assembles whole 16-sample output words (out position j <- source perm[j])
__global__ void __launch_bounds__(256) permute_packed_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out,
                                                             size_t stride, int64_t n, int64_t p,
                                                             const int32_t *__restrict__ perm)
{
    extern __shared__ __align__(16) uint8_t s_col[];
    const int nvec16 = (int)(stride >> 4), nwords = (int)(stride >> 2);
    for (int64_t i = blockIdx.x; i < p; i += gridDim.x) {
        const uint4 *src = reinterpret_cast<const uint4 *>(in + (size_t)i * stride);
        for (int k = threadIdx.x; k < nvec16; k += blockDim.x) reinterpret_cast<uint4 *>(s_col)[k] = __ldg(src + k);
        __syncthreads();
        uint32_t *dst = reinterpret_cast<uint32_t *>(out + (size_t)i * stride);
        for (int k = threadIdx.x; k < nwords; k += blockDim.x) {
            const int64_t base = (int64_t)k * 16;
            uint32_t w = 0u;
            if (base + 16 <= n) {
                const int4 *pp = reinterpret_cast<const int4 *>(perm + base);
#pragma unroll
                for (int v = 0; v < 4; v++) {
                    const int4 q4 = __ldg(pp + v);
                    const int sj[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
                    for (int f = 0; f < 4; f++)
                        w |= (((uint32_t)s_col[sj[f] >> 2] >> (2 * (sj[f] & 3))) & 3u) << (2 * (4 * v + f));
                }
            } else {
                for (int f = 0; f < 16 && base + f < n; f++) {
                    const int sj = __ldg(perm + base + f);
                    w |= (((uint32_t)s_col[sj >> 2] >> (2 * (sj & 3))) & 3u) << (2 * f);
                }
            }
            dst[k] = w;
        }
        __syncthreads();
    }
}

// FP64 `$dosage` block in PINNED (device-visible) host memory -> packed 2-bit columns written straight to HBM: the
// kernel reads the host matrix over PCIe (zero copy), 32 samples = 256 contiguous bytes per warp instruction, and
// builds the two code bit planes with warp ballots.  Used for a share of a chunk's variants while the host threads
// pack the rest (abi.cu: chunk driver), so that PCIe and host-memory bandwidth add up.
__device__ __forceinline__ uint32_t spread16(uint32_t x)      // bit f of the low half -> bit 2f
{
    x &= 0xffffu;
    x = (x | (x << 8)) & 0x00ff00ffu;
    x = (x | (x << 4)) & 0x0f0f0f0fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}
__global__ void __launch_bounds__(256) pack_f64_kernel(const double *__restrict__ G, int64_t n, int64_t pcount,
                                                       uint8_t *__restrict__ packed, size_t stride, int *__restrict__ bad)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t nseg = (int64_t)(stride >> 3), total = pcount * nseg;
    for (int64_t item = warp; item < total; item += nwarps) {
        const int64_t i = item / nseg, sgm = item - i * nseg, j = sgm * 32 + lane;
        const double v = j < n ? __ldg(G + i * n + j) : 0.0;
        const bool is1 = v == 1.0, is2 = v == 2.0, is0 = v == 0.0, miss = !(v > -1.0);   // NaN or <= -1 (matrix_flip_mean.cpp:26)
        const uint32_t lo = __ballot_sync(0xffffffffu, is1 || miss), hi = __ballot_sync(0xffffffffu, is2 || miss);
        const uint32_t okb = __ballot_sync(0xffffffffu, is0 || is1 || is2 || miss);
        if (lane == 0) {
            if (okb != 0xffffffffu) atomicOr(bad, 1);
            uint2 w;
            w.x = spread16(lo) | (spread16(hi) << 1);
            w.y = spread16(lo >> 16) | (spread16(hi >> 16) << 1);
            *reinterpret_cast<uint2 *>(packed + (size_t)i * stride + (size_t)sgm * 8) = w;
        }
    }
}

inline int grid_for(staar_ctx *ctx, int64_t work, int threads)
{
    return (int)std::max<int64_t>(1, std::min<int64_t>((work + threads - 1) / threads, (int64_t)ctx->sm_count * 16));
}

}  // namespace

int launch_csc_to_dense(staar_ctx *ctx, const double *d_val, const int64_t *d_row, const int64_t *d_colptr, int64_t n,
                        int64_t p, double *dG)
{
    STAAR_CUDA(cudaMemsetAsync(dG, 0, sizeof(double) * (size_t)n * p, ctx->stream));
    if (p == 0) return STAAR_OK;
    csc_scatter_kernel<<<grid_for(ctx, p * 32, 256), 256, 0, ctx->stream>>>(d_val, d_row, d_colptr, n, p, dG);
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

int launch_build_Y(staar_ctx *ctx, int64_t n, int ldY, const double *d_res, const double *d_SX, int q, const double *dA,
                   const double *dPA, int m, double *dY)
{
    build_Y_kernel<<<grid_for(ctx, n * ldY, 256), 256, 0, ctx->stream>>>(n, ldY, d_res, d_SX, q, dA, dPA, m, dY);
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

int launch_residual_update(staar_ctx *ctx, int64_t n, const double *dA, int m, const double *d_coef, const double *d_r, double *d_out)
{
    residual_update_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(n, dA, m, d_coef, d_r, d_out);
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

int launch_set_Y_col0(staar_ctx *ctx, int64_t n, int ldY, const double *d_res, double *dY)
{
    set_Y_col0_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(n, ldY, d_res, dY);
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

int launch_px_adj(staar_ctx *ctx, int64_t n, const int64_t *row_ptr, const int32_t *col, const double *val,
                  const double *dA, int m, const double *dSX, int q, const double *dT1, double *dOut)
{
    px_adj_kernel<<<grid_for(ctx, n * m, 256), 256, 0, ctx->stream>>>(n, row_ptr, col, val, dA, m, dSX, q, dT1, dOut);
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

int launch_permute_packed(staar_ctx *ctx, const uint8_t *d_in, uint8_t *d_out, size_t stride, int64_t n, int64_t p,
                          const int32_t *d_perm)
{
    if (p == 0) return STAAR_OK;
    if (stride > 227 * 1024) { set_error("permute_packed: column of %zu bytes does not fit in shared memory", stride); return STAAR_ERR_ARG; }
    if (!ctx->attr_permute) {                   // per context: the attribute belongs to the context's device
        STAAR_CUDA(cudaFuncSetAttribute(permute_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ctx->attr_permute = true;
    }
    const int grid = (int)std::min<int64_t>(p, (int64_t)ctx->sm_count * 4);
    permute_packed_kernel<<<grid, 256, stride, ctx->stream>>>(d_in, d_out, stride, n, p, d_perm);
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

int launch_pack_f64(staar_ctx *ctx, const double *G_dev_visible, int64_t n, int64_t pcount, uint8_t *d_packed, size_t stride,
                    int *d_bad)
{
    if (pcount == 0) return STAAR_OK;
    pack_f64_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(G_dev_visible, n, pcount, d_packed, stride, d_bad);
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

int launch_fill_pairs(staar_ctx *ctx, int64_t pp, int k, int32_t *d_colA, int32_t *d_colB)
{
    if (pp == 0) return STAAR_OK;
    fill_pairs_kernel<<<grid_for(ctx, pp * k * k, 256), 256, 0, ctx->stream>>>(pp, k, d_colA, d_colB);
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

}  // namespace staar
