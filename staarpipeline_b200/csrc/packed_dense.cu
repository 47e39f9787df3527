// packed_dense.cu — dense-GRM score test (reference src/Individual_Score_Test_denseGRM.cpp:9-61 / _sp_denseGRM.cpp) on
// RESIDENT 2-bit genotype columns: BASELINE config 3 without the FP64 matrix of the Rcpp boundary.
//
// Per variant: code counts -> AF / MAF / flip / imputation value exactly as matrix_flip_mean|minor (packed.cuh); the
// column's non-zero entries after flip and imputation become a carrier list (row, value) in sample order; then
//   Uscore = sum v_i r_i                      (quadform.cu:score_carriers)
//   g'Pg   = sum_i sum_j v_i v_j P[i,j]       (quadform.cu:quad_pairs_carriers_P)   when the list has <= n/8 entries,
//   g'Pg   = g . (P g) through the DMMA GEMM  (decode to FP64, quadform.cu)          for the common variants,
// and the reference's per-variant statistics (epilogue.cu).  Columns are in the caller's sample order (P's order).
#include <vector>
#include "common.cuh"
#include "packed.cuh"

namespace staar {

namespace {

__device__ __forceinline__ double carrier_value(uint32_t code, const VariantFlip &vf)
{
    if (code == 3u) return vf.mu;
    if (vf.flip && !(code & 1u)) code ^= 2u;
    return (double)code;
}

// a warp per variant: AF, MAF, flip, mu and the number of non-zero entries of the flipped / imputed column
__global__ void __launch_bounds__(256) carrier_count_kernel(const uint8_t *__restrict__ packed, size_t stride, int64_t n, int64_t p,
                                                            int impute, double *__restrict__ AF, double *__restrict__ MAF,
                                                            int32_t *__restrict__ flip, double *__restrict__ mu,
                                                            long long *__restrict__ count)
{
    const int lane = threadIdx.x & 31;
    for (int64_t i = blockIdx.x * 8ll + (threadIdx.x >> 5); i < p; i += gridDim.x * 8ll) {
        const PackedCounts pc = count_packed_column(packed + (size_t)i * stride, stride, lane);
        if (lane == 0) {
            const VariantFlip vf = flip_from_counts(pc, n, impute);
            const long long c0 = (long long)n - pc.c1 - pc.c2 - pc.cm;
            AF[i] = vf.af; MAF[i] = vf.maf; flip[i] = vf.flip; mu[i] = vf.mu;
            count[i] = pc.c1 + (vf.flip ? c0 : (long long)pc.c2) + (vf.mu != 0.0 ? pc.cm : 0);
        }
    }
}

// a warp per variant: the carrier list in increasing sample order (32 words = 512 samples per step, warp prefix sums)
__global__ void __launch_bounds__(256) carrier_fill_kernel(const uint8_t *__restrict__ packed, size_t stride, int64_t n, int64_t p,
                                                           const double *__restrict__ AF, const int32_t *__restrict__ flip,
                                                           const double *__restrict__ mu, const long long *__restrict__ colptr,
                                                           double *__restrict__ val, int64_t *__restrict__ row)
{
    const int lane = threadIdx.x & 31;
    const int64_t nwords = (n + 15) / 16;
    for (int64_t i = blockIdx.x * 8ll + (threadIdx.x >> 5); i < p; i += gridDim.x * 8ll) {
        const uint32_t *col = reinterpret_cast<const uint32_t *>(packed + (size_t)i * stride);
        const VariantFlip vf{AF[i], 0.0, mu[i], flip[i]};
        long long base = colptr[i];
        for (int64_t w0 = 0; w0 < nwords; w0 += 32) {
            const int64_t w = w0 + lane;
            uint32_t word = w < nwords ? __ldg(col + w) : 0u;
            const int valid = w < nwords ? (int)min(16ll, (long long)(n - 16 * w)) : 0;
            int cnt = 0;
            for (int s = 0; s < valid; s++) cnt += carrier_value((word >> (2 * s)) & 3u, vf) != 0.0;
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
            long long at = base + incl - cnt;
            for (int s = 0; s < valid; s++) {
                const double v = carrier_value((word >> (2 * s)) & 3u, vf);
                if (v != 0.0) { val[at] = v; row[at] = 16 * w + s; at++; }
            }
            base += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
}

// selected packed columns -> flipped, imputed FP64 columns (the matrix the reference's dense entry point receives)
__global__ void __launch_bounds__(256) decode_columns_kernel(const uint8_t *__restrict__ packed, size_t stride, int64_t n,
                                                             const int32_t *__restrict__ list, int64_t first, int64_t count,
                                                             const double *__restrict__ AF, const int32_t *__restrict__ flip,
                                                             const double *__restrict__ mu, double *__restrict__ G)
{
    for (int64_t c = blockIdx.x; c < count; c += gridDim.x) {
        const int64_t i = list[first + c];
        const uint8_t *col = packed + (size_t)i * stride;
        const VariantFlip vf{AF[i], 0.0, mu[i], flip[i]};
        double *g = G + (size_t)c * n;
        for (int64_t j = threadIdx.x; j < n; j += blockDim.x)
            g[j] = carrier_value(((uint32_t)__ldg(col + (j >> 2)) >> (2 * (j & 3))) & 3u, vf);
    }
}

// Selected variants of a resident block (null-model sample order) -> the flipped, imputed FP64 columns the reference's
// FP64 entry points receive, in the caller's sample order (perm: position -> original sample, nullptr = identity).  One
// CTA per variant: popcounts -> flip / imputation value as matrix_flip_mean|minor, then every code as a double.
__global__ void __launch_bounds__(256) decode_selected_kernel(const uint8_t *__restrict__ packed, size_t stride, int64_t n,
                                                              const long long *__restrict__ variant, int64_t count, int impute,
                                                              const int32_t *__restrict__ perm, double *__restrict__ G,
                                                              double *__restrict__ AF, double *__restrict__ MAF)
{
    __shared__ int s_cnt[3][8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int64_t c = blockIdx.x; c < count; c += gridDim.x) {
        const uint8_t *col = packed + (size_t)variant[c] * stride;
        int c1 = 0, c2 = 0, cm = 0;
        const uint4 *v4 = reinterpret_cast<const uint4 *>(col);
        for (int k = threadIdx.x; k < (int)(stride >> 4); k += blockDim.x) {
            const uint4 q = __ldg(v4 + k);
            count_word(q.x, c1, c2, cm); count_word(q.y, c1, c2, cm); count_word(q.z, c1, c2, cm); count_word(q.w, c1, c2, cm);
        }
        c1 = __reduce_add_sync(0xffffffffu, c1); c2 = __reduce_add_sync(0xffffffffu, c2); cm = __reduce_add_sync(0xffffffffu, cm);
        __syncthreads();
        if (lane == 0) { s_cnt[0][w] = c1; s_cnt[1][w] = c2; s_cnt[2][w] = cm; }
        __syncthreads();
        c1 = c2 = cm = 0;
        for (int ww = 0; ww < 8; ww++) { c1 += s_cnt[0][ww]; c2 += s_cnt[1][ww]; cm += s_cnt[2][ww]; }
        const VariantFlip vf = flip_from_counts(PackedCounts{c1, c2, cm}, n, impute);
        if (threadIdx.x == 0) { if (AF) AF[c] = vf.af; if (MAF) MAF[c] = vf.maf; }
        double *g = G + (size_t)c * n;
        for (int64_t pos = threadIdx.x; pos < n; pos += blockDim.x)
            g[perm ? perm[pos] : pos] = carrier_value(((uint32_t)__ldg(col + (pos >> 2)) >> (2 * (pos & 3))) & 3u, vf);
    }
}

__global__ void scatter_kernel(const double *__restrict__ src, const int32_t *__restrict__ list, int64_t first, int64_t count,
                               double *__restrict__ dst)
{
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < count; c += (int64_t)gridDim.x * blockDim.x)
        dst[list[first + c]] = src[c];
}

// Kronecker expansion Diagonal(k) (x) G0 of the carrier lists (R/Individual_Analysis.R:267): column t*pp + j holds the
// entries of G0's column j at rows t*n0 + i
__global__ void kron_expand_kernel(const double *__restrict__ val, const int64_t *__restrict__ row, const long long *__restrict__ colptr,
                                   int64_t n0, int64_t pp, int k, long long nnz, double *__restrict__ val_k,
                                   int64_t *__restrict__ row_k, long long *__restrict__ colptr_k)
{
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = tid; e < nnz * k; e += nthreads) {
        const int64_t t = e / nnz, e0 = e - t * nnz;
        val_k[e] = val[e0];
        row_k[e] = row[e0] + t * n0;
    }
    for (int64_t c = tid; c <= pp * k; c += nthreads) {
        const int64_t t = c == pp * k ? k - 1 : c / pp, j = c - t * pp;
        colptr_k[c] = t * nnz + colptr[j];
    }
}

// Multi-trait kernels: one CTA per variant of G0, variants drawn from an atomic counter (a common variant has 10^4-10^5
// carriers, a rare one 10^1-10^3: with a warp per column the common ones set the makespan).  Both read the carrier list
// of G0's column once for all k traits; the Kronecker expansion exists only as index arithmetic (row t*n0 + i).
//
// contraction: L[(t, j), c] = sum_e v_e Y[t*n0 + i_e, c], c < ncols = 1 + kq <= 64.  Warp w takes carriers w, w + 8, ...
// (a carrier's Y row is one coalesced read per trait; lanes = columns c and c + 32), partial sums meet in shared memory
// and are added in warp order.
template <int K>
__global__ void __launch_bounds__(256) contract_multi_packed_kernel(const double *__restrict__ val, const int64_t *__restrict__ row,
                                                                    const long long *__restrict__ colptr, int64_t n0, int64_t pp,
                                                                    const double *__restrict__ Y, int ldY, int ncols,
                                                                    double *__restrict__ L, int ldL, unsigned long long *work)
{
    __shared__ double part[8][K][64];
    __shared__ long long s_j;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool on0 = lane < ncols, on1 = lane + 32 < ncols;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_j = (long long)atomicAdd(work, 1ull);
        __syncthreads();
        const int64_t j = s_j;
        if (j >= pp) break;
        double a0[K], a1[K];
#pragma unroll
        for (int t = 0; t < K; t++) { a0[t] = 0.0; a1[t] = 0.0; }
        const long long e1 = colptr[j + 1];
        for (long long e = colptr[j] + warp; e < e1; e += 16) {               // two carriers per warp in flight
            const bool two = e + 8 < e1;
            const int64_t i0 = row[e], i1 = two ? row[e + 8] : 0;
            const double v0 = val[e], v1 = two ? val[e + 8] : 0.0;
#pragma unroll
            for (int t = 0; t < K; t++) {
                const double *y0 = Y + (size_t)(t * n0 + i0) * ldY, *y1 = Y + (size_t)(t * n0 + i1) * ldY;
                const double p0 = on0 ? y0[lane] : 0.0, p1 = on1 ? y0[lane + 32] : 0.0;
                const double q0 = (two && on0) ? y1[lane] : 0.0, q1 = (two && on1) ? y1[lane + 32] : 0.0;
                a0[t] += v0 * p0; a1[t] += v0 * p1;
                a0[t] += v1 * q0; a1[t] += v1 * q1;
            }
        }
#pragma unroll
        for (int t = 0; t < K; t++) { part[warp][t][lane] = a0[t]; part[warp][t][lane + 32] = a1[t]; }
        __syncthreads();
        for (int idx = threadIdx.x; idx < K * ncols; idx += 256) {
            const int t = idx / ncols, c = idx - t * ncols;
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < 8; w++) tot += part[w][t][c];
            L[((size_t)t * pp + j) * ldL + c] = tot;
        }
    }
}

// All k x k quadratic forms g_(s,j)' Sigma_i g_(t,j) of one variant in one sweep over its carriers (threads = carriers):
// for a carrier i and a trait block s, the row s*n0 + i of Sigma_i lists (trait t, relative i') entries, and g_(t,j)[i']
// is read straight from the variant's 2-bit column — no search in a carrier list.  Output index j*k*k + s*k + t as
// misc.cu:fill_pairs_kernel orders the pairs for the block test (src/Individual_Score_Test_sp_multi.cpp:47-62).
template <int K>
__global__ void __launch_bounds__(256) quad_multi_packed_kernel(const uint8_t *__restrict__ packed, size_t stride, int64_t n0, int64_t pp,
                                                                const double *__restrict__ AF, const int32_t *__restrict__ flip,
                                                                const double *__restrict__ mu, const double *__restrict__ val,
                                                                const int64_t *__restrict__ row, const long long *__restrict__ colptr,
                                                                const int64_t *__restrict__ S_row_ptr, const int32_t *__restrict__ S_col,
                                                                const double *__restrict__ S_val, double *__restrict__ quad,
                                                                unsigned long long *work)
{
    __shared__ double part[8][K * K];
    __shared__ long long s_j;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_j = (long long)atomicAdd(work, 1ull);
        __syncthreads();
        const int64_t j = s_j;
        if (j >= pp) break;
        const uint8_t *col = packed + (size_t)j * stride;
        const VariantFlip vf{AF[j], 0.0, mu[j], flip[j]};
        double acc[K * K];
#pragma unroll
        for (int a = 0; a < K * K; a++) acc[a] = 0.0;
        for (long long e = colptr[j] + threadIdx.x; e < colptr[j + 1]; e += 256) {
            const int64_t i = row[e];
            const double v = val[e];
#pragma unroll
            for (int s_ = 0; s_ < K; s_++) {
                const int64_t r = s_ * n0 + i;
                for (int64_t f = S_row_ptr[r]; f < S_row_ptr[r + 1]; f++) {
                    const int64_t c = S_col[f];
                    const int t = (int)(c / n0);
                    const int64_t i2 = c - t * n0;
                    const double gb = carrier_value(((uint32_t)__ldg(col + (i2 >> 2)) >> (2 * (i2 & 3))) & 3u, vf);
                    const double term = v * S_val[f] * gb;
#pragma unroll
                    for (int t_ = 0; t_ < K; t_++) if (t == t_) acc[s_ * K + t_] += term;
                }
            }
        }
#pragma unroll
        for (int a = 0; a < K * K; a++) {
            const double tot = warp_sum(acc[a]);
            if (lane == 0) part[warp][a] = tot;
        }
        __syncthreads();
        if (threadIdx.x < K * K) {
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < 8; w++) tot += part[w][threadIdx.x];
            quad[j * (K * K) + threadIdx.x] = tot;
        }
    }
}

}  // namespace

// Individual_Score_Test_multi / _sp_multi (src/Individual_Score_Test_sp_multi.cpp:9-68) on RESIDENT 2-bit columns of G0
// (n0 samples, caller's order): the null model is the nk x nk one of staar_nullmodel_set_sparse (trait-major stacking).
int packed_score_multi(staar_ctx *ctx, const uint8_t *d_packed, size_t stride, int64_t n0, int64_t pp, int k, int impute,
                       double *d_AF, double *d_MAF, double *d_Score, double *d_pl)
{
    const NullSparse &ns = ctx->ns;
    cudaStream_t st = ctx->stream;
    const int64_t p = pp * k;
    const size_t pi = (size_t)((pp + 1) / 2 * 2);
    STAAR_TRY(ctx->tmp2.reserve(sizeof(int32_t) * pi + sizeof(double) * (size_t)pp + sizeof(long long) * ((size_t)pp + 2 + (size_t)p + 2)));
    int32_t *dflip = ctx->tmp2.as<int32_t>();
    double *dmu = reinterpret_cast<double *>(dflip + pi);
    long long *dcount = reinterpret_cast<long long *>(dmu + pp), *dcolptr_k = dcount + pp + 2;
    const int grid8 = (int)std::min<int64_t>((pp + 7) / 8, (int64_t)ctx->sm_count * 8);
    carrier_count_kernel<<<grid8, 256, 0, st>>>(d_packed, stride, n0, pp, impute, d_AF, d_MAF, dflip, dmu, dcount);
    ctx->launches++;
    std::vector<long long> cnt((size_t)pp + 1);
    STAAR_CUDA(cudaMemcpyAsync(cnt.data(), dcount, sizeof(long long) * pp, cudaMemcpyDeviceToHost, st));
    STAAR_CUDA(cudaStreamSynchronize(st));
    long long run = 0;
    for (int64_t i = 0; i < pp; i++) { const long long c = cnt[i]; cnt[i] = run; run += c; }
    cnt[pp] = run;
    const long long nnz = run;
    STAAR_TRY(ctx->pk2.reserve((size_t)nnz * 16 * (size_t)(k + 1) + 64));
    double *dval = ctx->pk2.as<double>(), *dval_k = dval + nnz;
    int64_t *drow = reinterpret_cast<int64_t *>(dval_k + nnz * k), *drow_k = drow + nnz;
    STAAR_CUDA(cudaMemcpyAsync(dcount, cnt.data(), sizeof(long long) * (pp + 1), cudaMemcpyHostToDevice, st));
    carrier_fill_kernel<<<grid8, 256, 0, st>>>(d_packed, stride, n0, pp, d_AF, dflip, dmu, dcount, dval, drow);
    ctx->launches++;
    STAAR_CUDA(cudaStreamSynchronize(st));                        // cnt goes out of scope
    const int ldY = ns.ldY, ncols = (int)(1 + ns.q);
    STAAR_TRY(ctx->L.reserve(sizeof(double) * (size_t)p * ldY));
    STAAR_CUDA(cudaMemsetAsync(ctx->L.p, 0, sizeof(double) * (size_t)p * ldY, st));
    const int64_t npairs = pp * k * k;
    STAAR_TRY(ctx->tmp.reserve(sizeof(int32_t) * 2 * (size_t)npairs + sizeof(double) * (size_t)npairs + 64));
    double *dquad = ctx->tmp.as<double>();
    if (k <= 4 && ncols <= 64) {
        STAAR_TRY(ctx->ctr.reserve(2 * sizeof(unsigned long long)));
        unsigned long long *work = ctx->ctr.as<unsigned long long>();
        STAAR_CUDA(cudaMemsetAsync(work, 0, 2 * sizeof(unsigned long long), st));
        const int grid = (int)std::min<int64_t>(pp, (int64_t)ctx->sm_count * 8);
#define STAAR_MULTI(K)                                                                                                          \
    do {                                                                                                                        \
        contract_multi_packed_kernel<K><<<grid, 256, 0, st>>>(dval, drow, dcount, n0, pp, ns.Y, ldY, ncols, ctx->L.as<double>(), \
                                                              ldY, work);                                                       \
        quad_multi_packed_kernel<K><<<grid, 256, 0, st>>>(d_packed, stride, n0, pp, d_AF, dflip, dmu, dval, drow, dcount,        \
                                                          ns.row_ptr, ns.col, ns.val, dquad, work + 1);                         \
    } while (0)
        if (k == 1) STAAR_MULTI(1); else if (k == 2) STAAR_MULTI(2); else if (k == 3) STAAR_MULTI(3); else STAAR_MULTI(4);
#undef STAAR_MULTI
        ctx->launches += 2;
        STAAR_CUDA(cudaGetLastError());
    } else {                                  // generic: explicit Kronecker carrier lists and the pair kernels of quadform.cu
        kron_expand_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(dval, drow, dcount, n0, pp, k, nnz, dval_k, drow_k, dcolptr_k);
        ctx->launches++;
        const int64_t *cp = reinterpret_cast<const int64_t *>(dcolptr_k);
        STAAR_TRY(launch_contract_carriers(ctx, dval_k, drow_k, cp, p, ns.Y, ldY, ncols, ctx->L.as<double>(), ldY));
        int32_t *dcolA = reinterpret_cast<int32_t *>(dquad + npairs), *dcolB = dcolA + npairs;
        STAAR_TRY(launch_fill_pairs(ctx, pp, k, dcolA, dcolB));
        STAAR_TRY(launch_quad_pairs_carriers_csr(ctx, dval_k, drow_k, cp, ns.row_ptr, ns.col, ns.val, dcolA, dcolB, npairs, dquad));
    }
    STAAR_TRY(ctx->out.reserve(sizeof(double) * 5 * (size_t)p));
    double *o = ctx->out.as<double>();
    STAAR_TRY(launch_epilogue_general(ctx, p, k, ctx->L.as<double>(), ldY, (int)ns.q, ns.cov, dquad, 0, nullptr, nullptr, 0, 0,
                                      o, o + p, o + 2 * p, o + 3 * p, o + 4 * p));
    STAAR_CUDA(cudaMemcpyAsync(d_Score, o, sizeof(double) * p, cudaMemcpyDeviceToDevice, st));
    STAAR_CUDA(cudaMemcpyAsync(d_pl, o + 2 * p, sizeof(double) * pp, cudaMemcpyDeviceToDevice, st));
    return STAAR_OK;
}

int launch_decode_selected(staar_ctx *ctx, const uint8_t *d_packed, size_t stride, int64_t n, const long long *d_variant,
                           int64_t count, int impute, const int32_t *d_perm, double *dG, double *dAF, double *dMAF)
{
    if (count == 0) return STAAR_OK;
    decode_selected_kernel<<<(int)std::min<int64_t>(count, (int64_t)ctx->sm_count * 8), 256, 0, ctx->stream>>>(
        d_packed, stride, n, d_variant, count, impute, d_perm, dG, dAF, dMAF);
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

int packed_score_dense(staar_ctx *ctx, const uint8_t *d_packed, size_t stride, int64_t n, int64_t p, int impute,
                       double *d_AF, double *d_MAF, double *d_Score, double *d_Score_se, double *d_pl, double *d_Est,
                       double *d_Est_se)
{
    const NullDense &nd = ctx->nd;
    cudaStream_t st = ctx->stream;
    // per-variant scratch: flip (i32) | mu | count / colptr (i64, p + 1) | quad | L (p x 8)
    const size_t pi = (size_t)((p + 1) / 2 * 2);
    STAAR_TRY(ctx->tmp2.reserve(sizeof(int32_t) * pi + sizeof(double) * ((size_t)p * 10 + 8) + sizeof(long long) * ((size_t)p + 2)));
    int32_t *dflip = ctx->tmp2.as<int32_t>();
    double *dmu = reinterpret_cast<double *>(dflip + pi);
    long long *dcount = reinterpret_cast<long long *>(dmu + p);
    double *dquad = reinterpret_cast<double *>(dcount + p + 2), *dL = dquad + p;
    const int grid8 = (int)std::min<int64_t>((p + 7) / 8, (int64_t)ctx->sm_count * 8);
    carrier_count_kernel<<<grid8, 256, 0, st>>>(d_packed, stride, n, p, impute, d_AF, d_MAF, dflip, dmu, dcount);
    ctx->launches++;
    // column pointers and the two routes are decided on the host from the p counts (8 B per variant over PCIe)
    std::vector<long long> cnt((size_t)p + 1);
    STAAR_CUDA(cudaMemcpyAsync(cnt.data(), dcount, sizeof(long long) * p, cudaMemcpyDeviceToHost, st));
    STAAR_CUDA(cudaStreamSynchronize(st));
    std::vector<int32_t> lists((size_t)p);                       // sparse-route variants first, dense-route after
    int64_t ns_ = 0, nd_ = 0;
    long long run = 0;
    for (int64_t i = 0; i < p; i++) {
        const long long c = cnt[i];
        cnt[i] = run; run += c;
        if (c <= n / 8) lists[ns_++] = (int32_t)i;
    }
    cnt[p] = run;
    for (int64_t i = 0; i < p; i++)
        if (cnt[i + 1] - cnt[i] > n / 8) lists[ns_ + nd_++] = (int32_t)i;
    const long long nnz = run;
    STAAR_TRY(ctx->pk2.reserve((size_t)nnz * 16 + sizeof(int32_t) * (size_t)p + 64));
    double *dval = ctx->pk2.as<double>();
    int64_t *drow = reinterpret_cast<int64_t *>(dval + nnz);
    int32_t *dlist = reinterpret_cast<int32_t *>(drow + nnz);
    STAAR_CUDA(cudaMemcpyAsync(dcount, cnt.data(), sizeof(long long) * (p + 1), cudaMemcpyHostToDevice, st));
    STAAR_CUDA(cudaMemcpyAsync(dlist, lists.data(), sizeof(int32_t) * p, cudaMemcpyHostToDevice, st));
    carrier_fill_kernel<<<grid8, 256, 0, st>>>(d_packed, stride, n, p, d_AF, dflip, dmu, dcount, dval, drow);
    ctx->launches++;
    const int ldL = 8;
    STAAR_CUDA(cudaMemsetAsync(dL, 0, sizeof(double) * (size_t)p * ldL, st));
    STAAR_TRY(launch_score_carriers(ctx, dval, drow, reinterpret_cast<const int64_t *>(dcount), p, nd.residuals, dL, ldL));
    // sparse route: pairs (i, i) of the listed variants, results scattered to their variant
    STAAR_TRY(ctx->aux.reserve(sizeof(double) * (size_t)std::max<int64_t>(p, 1)));
    double *dtmp = ctx->aux.as<double>();
    if (ns_ > 0) {
        STAAR_TRY(launch_quad_pairs_carriers_P(ctx, dval, drow, reinterpret_cast<const int64_t *>(dcount), n, nd.P, dlist, dlist, ns_, dtmp));
        scatter_kernel<<<(unsigned)std::min<int64_t>((ns_ + 255) / 256, 1024), 256, 0, st>>>(dtmp, dlist, 0, ns_, dquad);
        ctx->launches++;
    }
    // dense route: batches of decoded FP64 columns through P*G
    if (nd_ > 0) {
        const int64_t batch = std::max<int64_t>(1, std::min<int64_t>(nd_, (1ll << 30) / (8 * n)));
        STAAR_TRY(ctx->G.reserve(sizeof(double) * (size_t)n * batch));
        STAAR_TRY(ctx->tmp.reserve(sizeof(double) * (size_t)n * batch + sizeof(int32_t) * 2 * (size_t)batch + 64));
        double *dW = ctx->tmp.as<double>();
        int32_t *didx = reinterpret_cast<int32_t *>(dW + (size_t)n * batch);
        std::vector<int32_t> idx((size_t)batch);
        for (int64_t c = 0; c < batch; c++) idx[c] = (int32_t)c;
        STAAR_CUDA(cudaMemcpyAsync(didx, idx.data(), sizeof(int32_t) * batch, cudaMemcpyHostToDevice, st));
        for (int64_t first = 0; first < nd_; first += batch) {
            const int64_t count = std::min(batch, nd_ - first);
            decode_columns_kernel<<<(int)std::min<int64_t>(count, (int64_t)ctx->sm_count * 8), 256, 0, st>>>(
                d_packed, stride, n, dlist, ns_ + first, count, d_AF, dflip, dmu, ctx->G.as<double>());
            ctx->launches++;
            STAAR_TRY(launch_dense_P_times_G(ctx, nd.P, ctx->G.as<double>(), n, count, dW));
            STAAR_TRY(launch_dot_pairs(ctx, ctx->G.as<double>(), dW, n, didx, didx, count, dtmp));
            scatter_kernel<<<(unsigned)std::min<int64_t>((count + 255) / 256, 1024), 256, 0, st>>>(dtmp, dlist, ns_ + first, count, dquad);
            ctx->launches++;
        }
        STAAR_CUDA(cudaStreamSynchronize(st));                    // idx goes out of scope
    }
    STAAR_TRY(launch_epilogue_general(ctx, p, 0, dL, ldL, 0, nullptr, dquad, 0, nullptr, nullptr, 0, 0, d_Score, d_Score_se, d_pl,
                                      d_Est, d_Est_se));
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

}  // namespace staar
