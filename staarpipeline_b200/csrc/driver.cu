// driver.cu — the per-chunk filtering of the single-variant driver on the device.
//
// R/Individual_Analysis.R:242,324 keeps a variant when it is common (MAF >= 0.05) or rare with
// (mac_cutoff - 0.5)/n/2 < MAF < 0.05 — i.e. when MAF > (mac_cutoff - 0.5)/n/2 — and, for imbalanced binary traits,
// sends the variants with exp(-pvalue_log) < 0.05 on to the SPA stage (:290-297).  Here the seven per-variant arrays
// of the resident fast path are compacted in variant order (stable, deterministic: block counts -> exclusive scan ->
// scatter), so that only the tested rows and one count cross PCIe.
#include "common.cuh"
#include "packed.cuh"

namespace staar {

namespace {

constexpr int kCompactThreads = 256;
constexpr int kCompactPerBlock = 4 * kCompactThreads;       // variants per block

struct CompactParams {
    int64_t p;
    const double *in[7];            // AF, MAF, Score, Score_se, pvalue_log, Est, Est_se
    double maf_cut;                 // keep when MAF > maf_cut
    double p_cut;                   // select when exp(-pvalue_log) < p_cut (<= 0: nothing selected)
    long long *block_keep, *block_sel;      // per-block counts, then exclusive offsets
    long long *index;               // compacted variant indices
    double *out[7];
    long long *sel_index;           // indices INTO THE COMPACTED TABLE of the selected (SPA-stage) rows
    long long *totals;              // [0] = kept, [1] = selected
};

__device__ __forceinline__ bool keep_variant(const CompactParams &P, int64_t i) { return P.in[1][i] > P.maf_cut; }
__device__ __forceinline__ bool select_variant(const CompactParams &P, int64_t i)
{
    return P.p_cut > 0.0 && P.in[1][i] > P.maf_cut && exp(-P.in[4][i]) < P.p_cut;
}

__global__ void __launch_bounds__(kCompactThreads) compact_count_kernel(CompactParams P)
{
    const int64_t base = (int64_t)blockIdx.x * kCompactPerBlock;
    int k = 0, s = 0;
    for (int t = threadIdx.x; t < kCompactPerBlock; t += kCompactThreads) {
        const int64_t i = base + t;
        if (i < P.p) { k += keep_variant(P, i); s += select_variant(P, i); }
    }
    __shared__ int sk[kCompactThreads / 32], ss[kCompactThreads / 32];
    k = __reduce_add_sync(0xffffffffu, k);
    s = __reduce_add_sync(0xffffffffu, s);
    if ((threadIdx.x & 31) == 0) { sk[threadIdx.x >> 5] = k; ss[threadIdx.x >> 5] = s; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int tk = 0, ts = 0;
        for (int w = 0; w < kCompactThreads / 32; w++) { tk += sk[w]; ts += ss[w]; }
        P.block_keep[blockIdx.x] = tk;
        P.block_sel[blockIdx.x] = ts;
    }
}

// exclusive scan of the per-block counts (one block; the number of blocks is p / 1024)
__global__ void __launch_bounds__(1024) compact_scan_kernel(long long *block_keep, long long *block_sel, int64_t nblocks,
                                                            long long *totals)
{
    __shared__ long long carry[2];
    __shared__ long long warp_sum[2][32];
    if (threadIdx.x == 0) { carry[0] = 0; carry[1] = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int64_t b0 = 0; b0 < nblocks; b0 += 1024) {
        const int64_t b = b0 + threadIdx.x;
        long long v[2] = {b < nblocks ? block_keep[b] : 0, b < nblocks ? block_sel[b] : 0};
        long long incl[2];
        for (int a = 0; a < 2; a++) {
            long long x = v[a];
            for (int o = 1; o < 32; o <<= 1) { const long long y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            incl[a] = x;
            if (lane == 31) warp_sum[a][w] = x;
        }
        __syncthreads();
        for (int a = 0; a < 2; a++) {
            long long before = carry[a];
            for (int ww = 0; ww < w; ww++) before += warp_sum[a][ww];
            const long long excl = before + incl[a] - v[a];
            if (b < nblocks) (a == 0 ? block_keep : block_sel)[b] = excl;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int a = 0; a < 2; a++) { long long t = 0; for (int ww = 0; ww < 32; ww++) t += warp_sum[a][ww]; carry[a] += t; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { totals[0] = carry[0]; totals[1] = carry[1]; }
}

__global__ void __launch_bounds__(kCompactThreads) compact_scatter_kernel(CompactParams P)
{
    __shared__ int warp_keep[kCompactThreads / 32], warp_sel[kCompactThreads / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * kCompactPerBlock;
    long long off_keep = P.block_keep[blockIdx.x], off_sel = P.block_sel[blockIdx.x];
    for (int t0 = 0; t0 < kCompactPerBlock; t0 += kCompactThreads) {          // variant order is preserved
        const int64_t i = base + t0 + threadIdx.x;
        const bool k = i < P.p && keep_variant(P, i), s = i < P.p && select_variant(P, i);
        const uint32_t bk = __ballot_sync(0xffffffffu, k), bs = __ballot_sync(0xffffffffu, s);
        if (lane == 0) { warp_keep[w] = __popc(bk); warp_sel[w] = __popc(bs); }
        __syncthreads();
        long long pk = off_keep, ps = off_sel;
        int tot_k = 0, tot_s = 0;
        for (int ww = 0; ww < kCompactThreads / 32; ww++) {
            if (ww < w) { pk += warp_keep[ww]; ps += warp_sel[ww]; }
            tot_k += warp_keep[ww]; tot_s += warp_sel[ww];
        }
        pk += __popc(bk & ((1u << lane) - 1u));
        ps += __popc(bs & ((1u << lane) - 1u));
        if (k) {
            P.index[pk] = i;
#pragma unroll
            for (int a = 0; a < 7; a++) P.out[a][pk] = P.in[a][i];
            if (s) P.sel_index[ps] = pk;
        }
        off_keep += tot_k; off_sel += tot_s;
        __syncthreads();
    }
}

}  // namespace

int launch_compact_results(staar_ctx *ctx, int64_t p, const double *const d_in[7], double maf_cut, double p_cut,
                           long long *d_index, double *const d_out[7], long long *d_sel_index, long long *d_totals)
{
    if (p == 0) {
        STAAR_CUDA(cudaMemsetAsync(d_totals, 0, 2 * sizeof(long long), ctx->stream));
        return STAAR_OK;
    }
    const int64_t nblocks = (p + kCompactPerBlock - 1) / kCompactPerBlock;
    STAAR_TRY(ctx->aux.reserve(sizeof(long long) * 2 * (size_t)nblocks));
    CompactParams P;
    P.p = p; P.maf_cut = maf_cut; P.p_cut = p_cut;
    for (int a = 0; a < 7; a++) { P.in[a] = d_in[a]; P.out[a] = d_out[a]; }
    P.block_keep = ctx->aux.as<long long>(); P.block_sel = P.block_keep + nblocks;
    P.index = d_index; P.sel_index = d_sel_index; P.totals = d_totals;
    compact_count_kernel<<<(unsigned)nblocks, kCompactThreads, 0, ctx->stream>>>(P);
    compact_scan_kernel<<<1, 1024, 0, ctx->stream>>>(P.block_keep, P.block_sel, nblocks, d_totals);
    compact_scatter_kernel<<<(unsigned)nblocks, kCompactThreads, 0, ctx->stream>>>(P);
    ctx->launches += 3;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

}  // namespace staar
