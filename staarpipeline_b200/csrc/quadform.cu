// quadform.cu — K3/K4: the quadratic forms g_a' * S * g_b behind diag(Cov) and the k x k trait blocks.
//
//  * quad_pairs_csr  : S = Sigma_i sparse (CSR, symmetric, full rows) — the (trans(G)*Sigma_i)*G factor of
//                      src/Individual_Score_Test.cpp:43, _sp.cpp:43, _multi.cpp:40, _cond.cpp:49, but only the
//                      entries the reference actually reads: the diagonal, or Cov(id_single,id_single)
//                      (src/Individual_Score_Test_multi.cpp:47-52).
//  * dense_P_times_G : W = P*G for the dense-GRM family (src/Individual_Score_Test_denseGRM.cpp:37) as an
//                      FP64 tensor-core GEMM (mma.sync m8n8k4 DMMA, 128x64 CTA tile, next tiles
//                      prefetched into registers under the MMAs); dot_pairs then takes g_a' * W[:,b].
//  * quad_pairs_carriers_P / score_carriers : the same dense-GRM quantities for SPARSE genotype columns
//                      (src/Individual_Score_Test_sp_denseGRM.cpp:17,37 — the route of rare variants, MAF < 0.05,
//                      R/Individual_Analysis.R:372) taken over the carriers only: g_a' P g_b = sum_{i in a} sum_{j in b}
//                      v_i v_j P[i,j], c_a*c_b gathers from P instead of 2 n^2 flop of P*G (a singleton with 40
//                      carriers at n = 20 000: 820 loads against 8e8 flop).
// FP64 throughout (1e-10 parity); deterministic tree reductions, no atomics.
#include "common.cuh"

namespace staar {

namespace {

__global__ void __launch_bounds__(256) quad_pairs_csr_kernel(const double *__restrict__ G, int64_t n,
                                                             const int64_t *__restrict__ row_ptr,
                                                             const int32_t *__restrict__ col,
                                                             const double *__restrict__ val,
                                                             const int32_t *__restrict__ colA,
                                                             const int32_t *__restrict__ colB, int64_t npairs,
                                                             double *__restrict__ out)
{
    __shared__ double red[32];
    for (int64_t pr = blockIdx.x; pr < npairs; pr += gridDim.x) {
        const double *ga = G + (int64_t)colA[pr] * n;
        const double *gb = G + (int64_t)colB[pr] * n;
        double acc = 0.0;
        for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
            const double a = ga[j];
            if (a != 0.0) {
                double s = 0.0;
                const int64_t e1 = row_ptr[j + 1];
                for (int64_t e = row_ptr[j]; e < e1; e++) s += val[e] * gb[col[e]];
                acc += a * s;
            }
        }
        acc = block_sum(acc, red);
        if (threadIdx.x == 0) out[pr] = acc;
    }
}

__global__ void __launch_bounds__(256) dot_pairs_kernel(const double *__restrict__ G, const double *__restrict__ W,
                                                        int64_t n, const int32_t *__restrict__ colA,
                                                        const int32_t *__restrict__ colB, int64_t npairs,
                                                        double *__restrict__ out)
{
    __shared__ double red[32];
    for (int64_t pr = blockIdx.x; pr < npairs; pr += gridDim.x) {
        const double *ga = G + (int64_t)colA[pr] * n;
        const double *wb = W + (int64_t)colB[pr] * n;
        double acc = 0.0;
        for (int64_t j = threadIdx.x; j < n; j += blockDim.x) acc += ga[j] * wb[j];
        acc = block_sum(acc, red);
        if (threadIdx.x == 0) out[pr] = acc;
    }
}


// One CTA per (a, b) pair of sparse columns; the b list sits in shared memory in tiles, the a list is walked serially
// with the threads spread over b.  For a == b only j >= i is visited (P is symmetric: Sigma^-1 - Sigma^-1 X cov X' Sigma^-1).
// ASSUMPTION: P is symmetric, as GMMAT's P always is; the reference's expression uses every entry of P, so for an
// asymmetric P the a == b shortcut (off-diagonal terms counted twice from the upper triangle) would differ from it.
constexpr int kCarrierTile = 1024;
__global__ void __launch_bounds__(256) quad_pairs_carriers_P_kernel(const double *__restrict__ val, const int64_t *__restrict__ row,
                                                                    const int64_t *__restrict__ colptr, int64_t n,
                                                                    const double *__restrict__ P,
                                                                    const int32_t *__restrict__ colA,
                                                                    const int32_t *__restrict__ colB, int64_t npairs,
                                                                    double *__restrict__ out)
{
    __shared__ double red[32];
    __shared__ double bv[kCarrierTile];
    __shared__ int32_t bi[kCarrierTile];
    for (int64_t pr = blockIdx.x; pr < npairs; pr += gridDim.x) {
        const int ca = colA[pr], cb = colB[pr];
        const int64_t a0 = colptr[ca], a1 = colptr[ca + 1], b0 = colptr[cb], b1 = colptr[cb + 1];
        const bool same = ca == cb;
        double acc = 0.0;
        for (int64_t t0 = b0; t0 < b1; t0 += kCarrierTile) {
            const int tn = (int)min((long long)kCarrierTile, (long long)(b1 - t0));
            __syncthreads();
            for (int t = threadIdx.x; t < tn; t += blockDim.x) { bv[t] = val[t0 + t]; bi[t] = (int32_t)row[t0 + t]; }
            __syncthreads();
            // same column: entry e of the a list pairs with entries >= e of the b list (off-diagonal terms twice)
            const int64_t e_end = same ? min((long long)a1, (long long)(t0 + tn)) : a1;
            for (int64_t e = a0; e < e_end; e += 4) {                  // four rows of P in flight per thread
                double va[4], s[4];
                const double *Pi[4];
                int first[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int64_t ek = e + k;
                    const bool live = ek < e_end;
                    va[k] = live ? val[ek] : 0.0;
                    Pi[k] = P + (size_t)row[live ? ek : e] * n;        // column i of P (= row i)
                    first[k] = !live ? tn : (same ? (int)max(0ll, (long long)(ek - t0)) : 0);
                    s[k] = 0.0;
                }
                for (int t = first[0] + threadIdx.x; t < tn; t += blockDim.x) {
                    const double b = bv[t];
                    const int32_t j = bi[t];
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (t >= first[k]) {
                            const double term = b * Pi[k][j];
                            s[k] += (same && t0 + t != e + k) ? 2.0 * term : term;
                        }
                }
#pragma unroll
                for (int k = 0; k < 4; k++) acc += va[k] * s[k];
            }
        }
        acc = block_sum(acc, red);
        if (threadIdx.x == 0) out[pr] = acc;
    }
}

// Uscore of sparse columns: L[j*ldL] = sum_e val[e] * residuals[row[e]]   (a warp per column)
__global__ void __launch_bounds__(256) score_carriers_kernel(const double *__restrict__ val, const int64_t *__restrict__ row,
                                                             const int64_t *__restrict__ colptr, int64_t p,
                                                             const double *__restrict__ residuals, double *__restrict__ L, int ldL)
{
    const int lane = threadIdx.x & 31;
    for (int64_t j = blockIdx.x * 8ll + (threadIdx.x >> 5); j < p; j += gridDim.x * 8ll) {
        double s = 0.0;
        for (int64_t e = colptr[j] + lane; e < colptr[j + 1]; e += 32) s += val[e] * residuals[row[e]];
        s = warp_sum(s);
        if (lane == 0) L[j * ldL] = s;
    }
}


// ---------------------------------------------------------------- sparse genotype columns, sparse Sigma_i
// The rare-variant route of the reference (Individual_Score_Test_sp / _sp_multi, src/Individual_Score_Test_sp.cpp:17,42-43;
// _sp_multi.cpp:18,39-40) keeps G sparse.  Here too: everything is taken over the stored entries of a column.
//   contract_carriers : L[j, c] = sum_e val[e] * Y[row[e], c]          (a warp per column, lanes = columns of Y)
//   quad_pairs_carriers_csr : g_a' Sigma_i g_b = sum_{e in a} val[e] * sum_{f in row(row[e]) of Sigma_i} S_f * g_b[col_f],
//                             g_b looked up by binary search in b's sorted row list (a warp per pair)
__global__ void __launch_bounds__(256) contract_carriers_kernel(const double *__restrict__ val, const int64_t *__restrict__ row,
                                                                const int64_t *__restrict__ colptr, int64_t p,
                                                                const double *__restrict__ Y, int ldY, int ncols,
                                                                double *__restrict__ L, int ldL)
{
    const int lane = threadIdx.x & 31;
    for (int64_t j = blockIdx.x * 8ll + (threadIdx.x >> 5); j < p; j += gridDim.x * 8ll) {
        const int64_t e0 = colptr[j], e1 = colptr[j + 1];
        for (int c0 = 0; c0 < ncols; c0 += 32) {
            const int c = c0 + lane;
            const bool on = c < ncols;
            double acc = 0.0;
            int64_t e = e0;
            for (; e + 4 <= e1; e += 4) {                              // four rows of Y in flight
                double v[4], y[4];
#pragma unroll
                for (int k = 0; k < 4; k++) { v[k] = val[e + k]; y[k] = on ? Y[(size_t)row[e + k] * ldY + c] : 0.0; }
#pragma unroll
                for (int k = 0; k < 4; k++) acc += v[k] * y[k];
            }
            for (; e < e1; e++) acc += val[e] * (on ? Y[(size_t)row[e] * ldY + c] : 0.0);
            if (on) L[j * ldL + c] = acc;
        }
    }
}

__device__ __forceinline__ double carrier_lookup(const double *__restrict__ val, const int64_t *__restrict__ row, int64_t b0,
                                                 int64_t b1, int64_t want)
{
    while (b0 < b1) {
        const int64_t mid = (b0 + b1) >> 1;
        const int64_t r = row[mid];
        if (r == want) return val[mid];
        if (r < want) b0 = mid + 1; else b1 = mid;
    }
    return 0.0;
}

__global__ void __launch_bounds__(256) quad_pairs_carriers_csr_kernel(const double *__restrict__ val, const int64_t *__restrict__ row,
                                                                      const int64_t *__restrict__ colptr,
                                                                      const int64_t *__restrict__ S_row_ptr,
                                                                      const int32_t *__restrict__ S_col,
                                                                      const double *__restrict__ S_val,
                                                                      const int32_t *__restrict__ colA,
                                                                      const int32_t *__restrict__ colB, int64_t npairs,
                                                                      double *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    for (int64_t pr = blockIdx.x * 8ll + (threadIdx.x >> 5); pr < npairs; pr += gridDim.x * 8ll) {
        const int ca = colA[pr], cb = colB[pr];
        const int64_t a0 = colptr[ca], a1 = colptr[ca + 1], b0 = colptr[cb], b1 = colptr[cb + 1];
        double acc = 0.0;
        if (b1 > b0)
            for (int64_t e = a0 + lane; e < a1; e += 32) {
                const int64_t i = row[e];
                double s = 0.0;
                for (int64_t f = S_row_ptr[i]; f < S_row_ptr[i + 1]; f++) {
                    const double gb = carrier_lookup(val, row, b0, b1, S_col[f]);
                    if (gb != 0.0) s += S_val[f] * gb;
                }
                acc += val[e] * s;
            }
        acc = warp_sum(acc);
        if (lane == 0) out[pr] = acc;
    }
}

// ---------------------------------------------------------------- W = P * G  (FP64 DMMA GEMM)
constexpr int BM = 128, BN = 64, BK = 16;
constexpr int RSA = BM + 4;      // sP[k][i], stride = 4 mod 16 doubles -> conflict-free A fragments
constexpr int RSB = BK + 4;      // sG[v][k], stride = 4 mod 16 doubles -> conflict-free B fragments

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// P: n x n column-major; G: n x p column-major; W: n x p column-major.  256 threads = 8 warps (4 x 2),
// warp tile 32 x 32.
__global__ void __launch_bounds__(256) dgemm_PG_kernel(const double *__restrict__ P, const double *__restrict__ G,
                                                       int64_t n, int64_t p, double *__restrict__ W)
{
    __shared__ double sP[BK * RSA];
    __shared__ double sG[BN * RSB];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int a = lane & 3, b = lane >> 2;
    const int wm = warp & 3, wn = warp >> 2;
    const int64_t i0 = (int64_t)blockIdx.x * BM, v0 = (int64_t)blockIdx.y * BN;

    double acc[4][4][2];
#pragma unroll
    for (int mt = 0; mt < 4; mt++)
#pragma unroll
        for (int nt = 0; nt < 4; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

    double rp[8], rg[4];
    auto load_tiles = [&](int64_t k0) {
#pragma unroll
        for (int u = 0; u < 8; u++) {           // P tile: 16 k-columns x 128 rows
            const int idx = tid + u * 256, kk = idx >> 7, ii = idx & 127;
            const int64_t i = i0 + ii, k = k0 + kk;
            rp[u] = (i < n && k < n) ? P[i + k * n] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {           // G tile: 64 variants x 16 samples
            const int idx = tid + u * 256, vv = idx >> 4, kk = idx & 15;
            const int64_t v = v0 + vv, k = k0 + kk;
            rg[u] = (v < p && k < n) ? G[k + v * n] : 0.0;
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int idx = tid + u * 256, kk = idx >> 7, ii = idx & 127;
            sP[kk * RSA + ii] = rp[u];
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int idx = tid + u * 256, vv = idx >> 4, kk = idx & 15;
            sG[vv * RSB + kk] = rg[u];
        }
    };

    const int64_t nk = (n + BK - 1) / BK;
    load_tiles(0);
    for (int64_t kt = 0; kt < nk; kt++) {
        store_tiles();
        __syncthreads();
        if (kt + 1 < nk) load_tiles((kt + 1) * BK);   // global loads in flight under the DMMAs
#pragma unroll
        for (int ks = 0; ks < BK / 4; ks++) {
            double af[4], bf[4];
#pragma unroll
            for (int mt = 0; mt < 4; mt++) af[mt] = sP[(ks * 4 + a) * RSA + wm * 32 + mt * 8 + b];
#pragma unroll
            for (int nt = 0; nt < 4; nt++) bf[nt] = sG[(wn * 32 + nt * 8 + b) * RSB + ks * 4 + a];
#pragma unroll
            for (int mt = 0; mt < 4; mt++)
#pragma unroll
                for (int nt = 0; nt < 4; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int mt = 0; mt < 4; mt++)
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
            const int64_t i = i0 + wm * 32 + mt * 8 + b;
            const int64_t v = v0 + wn * 32 + nt * 8 + 2 * a;
            if (i < n) {
                if (v < p) W[i + v * n] = acc[mt][nt][0];
                if (v + 1 < p) W[i + (v + 1) * n] = acc[mt][nt][1];
            }
        }
}

}  // namespace

int launch_quad_pairs_csr(staar_ctx *ctx, const double *dG, int64_t n, const int64_t *row_ptr, const int32_t *col,
                          const double *val, const int32_t *d_colA, const int32_t *d_colB, int64_t npairs,
                          double *d_out)
{
    if (npairs == 0) return STAAR_OK;
    int grid = (int)std::min<int64_t>(npairs, (int64_t)ctx->sm_count * 16);
    quad_pairs_csr_kernel<<<grid, 256, 0, ctx->stream>>>(dG, n, row_ptr, col, val, d_colA, d_colB, npairs, d_out);
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

int launch_dot_pairs(staar_ctx *ctx, const double *dG, const double *dW, int64_t n, const int32_t *d_colA,
                     const int32_t *d_colB, int64_t npairs, double *d_out)
{
    if (npairs == 0) return STAAR_OK;
    int grid = (int)std::min<int64_t>(npairs, (int64_t)ctx->sm_count * 16);
    dot_pairs_kernel<<<grid, 256, 0, ctx->stream>>>(dG, dW, n, d_colA, d_colB, npairs, d_out);
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

int launch_quad_pairs_carriers_P(staar_ctx *ctx, const double *d_val, const int64_t *d_row, const int64_t *d_colptr, int64_t n,
                                 const double *dP, const int32_t *d_colA, const int32_t *d_colB, int64_t npairs, double *d_out)
{
    if (npairs == 0) return STAAR_OK;
    const int grid = (int)std::min<int64_t>(npairs, (int64_t)ctx->sm_count * 8);
    quad_pairs_carriers_P_kernel<<<grid, 256, 0, ctx->stream>>>(d_val, d_row, d_colptr, n, dP, d_colA, d_colB, npairs, d_out);
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

int launch_score_carriers(staar_ctx *ctx, const double *d_val, const int64_t *d_row, const int64_t *d_colptr, int64_t p,
                          const double *d_res, double *dL, int ldL)
{
    if (p == 0) return STAAR_OK;
    const int grid = (int)std::min<int64_t>((p + 7) / 8, (int64_t)ctx->sm_count * 8);
    score_carriers_kernel<<<grid, 256, 0, ctx->stream>>>(d_val, d_row, d_colptr, p, d_res, dL, ldL);
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

int launch_contract_carriers(staar_ctx *ctx, const double *d_val, const int64_t *d_row, const int64_t *d_colptr, int64_t p,
                            const double *dY, int ldY, int ncols, double *dL, int ldL)
{
    if (p == 0) return STAAR_OK;
    const int grid = (int)std::min<int64_t>((p + 7) / 8, (int64_t)ctx->sm_count * 8);
    contract_carriers_kernel<<<grid, 256, 0, ctx->stream>>>(d_val, d_row, d_colptr, p, dY, ldY, ncols, dL, ldL);
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

int launch_quad_pairs_carriers_csr(staar_ctx *ctx, const double *d_val, const int64_t *d_row, const int64_t *d_colptr,
                                   const int64_t *S_row_ptr, const int32_t *S_col, const double *S_val, const int32_t *d_colA,
                                   const int32_t *d_colB, int64_t npairs, double *d_out)
{
    if (npairs == 0) return STAAR_OK;
    const int grid = (int)std::min<int64_t>((npairs + 7) / 8, (int64_t)ctx->sm_count * 8);
    quad_pairs_carriers_csr_kernel<<<grid, 256, 0, ctx->stream>>>(d_val, d_row, d_colptr, S_row_ptr, S_col, S_val, d_colA, d_colB,
                                                                 npairs, d_out);
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

int launch_dense_P_times_G(staar_ctx *ctx, const double *dP, const double *dG, int64_t n, int64_t p, double *dW)
{
    if (p == 0 || n == 0) return STAAR_OK;
    dim3 grid((unsigned)((n + BM - 1) / BM), (unsigned)((p + BN - 1) / BN));
    dgemm_PG_kernel<<<grid, 256, 0, ctx->stream>>>(dP, dG, n, p, dW);
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

}  // namespace staar
