// contract_packed.cu — K2 for packed dosages: G'[residuals | Sigma_iX] as an EXACT integer tensor-core contraction.
//
// This is the first factor of every Cov in the reference (Uscore = trans(residuals)*G, tSigma_iX_G =
// trans(Sigma_iX)*G; src/Individual_Score_Test.cpp:17,42) for genotypes resident in HBM as 2-bit codes.
//
// Why not FP64 MMA: B200 measures 37 TFLOP/s for both DFMA and mma.sync f64 (profiles/r01_pipe_peaks.json), which
// caps a dense FP64 contraction at ~11 M variants/s at n = 1e5 — 4 % of the HBM roofline for 2-bit genotypes.
// Genotypes are exact small integers, so instead the FP64 operand Y = [r | Sigma_iX] is split ONCE, at null-model
// set-up, into 7 balanced base-256 digits per entry (54-bit fixed point per column, Ozaki-style slicing):
//     Y[j,c] ~= 2^(e_c-54) * sum_s d_s[j,c] * 256^(6-s),   d_s in [-128,127]
// and G' * d_s is an INT8 x INT8 -> INT32 tensor-core GEMM whose result is exact (|acc| <= 3*128*n < 2^31).
// The 7 partial results are recombined in 64-bit integer arithmetic, so the contraction is bit-reproducible for any
// tiling / split / GPU count, and the only rounding left is the 2^-54 truncation of Y (relative to the column
// maximum) — far inside the 1e-10 budget.
// The kernel contracts the RAW 2-bit codes (0, 1, 2 and 3 = missing, taken as the value 3): the flip to the minor
// allele g -> 2 - g and the imputation of missing entries are linear in these exact integer sums and are applied
// analytically by finalize_packed_kernel (scan_packed.cu), so the contraction does not depend on the scan.
//
// This file is the LEGACY tensor path (mma.sync.m16n8k32.s8; measured 1.14 POP/s on B200), kept as a cross-check of
// the tcgen05 kernel (STAAR_B200_CONTRACT=mma) and because the host-side slicing of Y lives here.
// Kernel: M = variants (2 m16 tiles per warp, 8 warps per CTA), N = 7 slices x 16 columns = 14 n8 tiles, K = samples.
// The digit operand is pre-arranged in fragment-major order so a k-chunk (256 samples, 28 KiB) is ONE contiguous
// bulk copy: it is streamed from L2 into a 4-stage shared-memory ring by cp.async.bulk (TMA 1-D) completing on
// mbarriers, and read back with conflict-free 128-bit LDS.  Genotype words go HBM -> registers (128-bit,
// L1::no_allocate) and are expanded 2-bit -> int8 with shift/mask only; every genotype byte is read once per launch.
#include "common.cuh"
#include "packed.cuh"

namespace staar {

namespace {

constexpr int kWarpsC = 8;
constexpr int kMT = 2;                       // m16 tiles per warp
constexpr int kVarPerCta = kWarpsC * kMT * 16;   // 256
constexpr int kStages = 4;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void imma16832(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// field f (0..3) of each of the 4 bytes of w as 4 int8 lanes (raw codes 0..3)
__device__ __forceinline__ uint32_t fields_to_s8(uint32_t w, int f)
{
    return (w >> (2 * f)) & 0x03030303u;
}

// grid = (ceil(p / 256), ksplit).  out_hi / out_lo: p x 16 int64, zero-initialised when ksplit > 1.
__global__ void __launch_bounds__(kWarpsC * 32, 1)
contract_packed_kernel(const uint8_t *__restrict__ packed, size_t stride, int64_t p,
                       const int8_t *__restrict__ Yfrag,
                       int64_t n_chunks, int64_t chunks_per_split, long long *__restrict__ out_hi,
                       long long *__restrict__ out_lo, int use_atomics)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *sB = smem;                                                  // kStages * kChunkBytesB
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + kStages * kChunkBytesB);
    uint64_t *empty = full + kStages;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int64_t kc_begin = (int64_t)blockIdx.y * chunks_per_split;
    const int64_t kc_end = min(n_chunks, kc_begin + chunks_per_split);
    const int64_t nkc = kc_end - kc_begin;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], kWarpsC); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // Thread 0 doubles as the producer of the digit ring (a 9th warp would put 3 warps on one SM sub-partition and
    // cap the kernel at 168 registers).  produce(must) issues chunks in order: it BLOCKS only for chunks <= must
    // (needed next by this very warp, so their slot is released by work that is already fully fed) and otherwise
    // only takes slots that test free, so it can never wait on something it has not yet issued itself.
    int64_t produced = 0;
    auto produce = [&](int64_t must) {
        while (produced < nkc && produced < must + kStages) {
            const int stage = (int)(produced % kStages);
            if (produced >= kStages) {
                const uint32_t par = (uint32_t)(((produced / kStages) - 1) & 1);
                if (produced <= must) mbar_wait(&empty[stage], par);
                else if (!mbar_test(&empty[stage], par)) break;
            }
            mbar_expect_tx(&full[stage], kChunkBytesB);
            bulk_g2s(sB + (size_t)stage * kChunkBytesB, Yfrag + (size_t)(kc_begin + produced) * kChunkBytesB,
                     kChunkBytesB, &full[stage]);
            produced++;
        }
    };

    // this lane's four genotype rows: (mt, half) -> variant v0 + mt*16 + half*8 + g
    const int64_t v0 = (int64_t)blockIdx.x * kVarPerCta + warp * (kMT * 16);
    const uint8_t *rowp[kMT][2];
    bool rok[kMT][2];
#pragma unroll
    for (int mt = 0; mt < kMT; mt++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int64_t v = v0 + mt * 16 + h * 8 + g;
            rok[mt][h] = v < p;
            rowp[mt][h] = packed + (size_t)(rok[mt][h] ? v : 0) * stride;
        }
    auto load_A = [&](int64_t kc, uint4 (&dst)[kMT][2]) {
        const size_t off = (size_t)kc * 64 + (size_t)t * 16;
#pragma unroll
        for (int mt = 0; mt < kMT; mt++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
                // raw codes only: touching the loaded value here would stall on the prefetch
                if (rok[mt][h] && off + 16 <= stride) dst[mt][h] = ldg_nc_u4(rowp[mt][h] + off);
                else dst[mt][h] = make_uint4(0, 0, 0, 0);
            }
    };

    int acc[kMT][kNT][4];
#pragma unroll
    for (int mt = 0; mt < kMT; mt++) {
#pragma unroll
        for (int j = 0; j < kNT; j++)
#pragma unroll
            for (int e = 0; e < 4; e++) acc[mt][j][e] = 0;
    }

    auto adopt = [&](uint4 (&cur)[kMT][2], const uint4 (&nxt)[kMT][2]) {
#pragma unroll
        for (int mt = 0; mt < kMT; mt++)
#pragma unroll
            for (int h = 0; h < 2; h++) cur[mt][h] = nxt[mt][h];
    };
    uint4 Acur[kMT][2], Anext[kMT][2];
    if (nkc > 0) { load_A(kc_begin, Anext); adopt(Acur, Anext); }

    for (int64_t it = 0; it < nkc; it++) {
        const int stage = (int)(it % kStages);
        const uint32_t phase = (uint32_t)((it / kStages) & 1);
        if (it + 1 < nkc) load_A(kc_begin + it + 1, Anext);
        if (threadIdx.x == 0) produce(it);
        mbar_wait(&full[stage], phase);
        const uint8_t *sb = sB + (size_t)stage * kChunkBytesB;
#pragma unroll
        for (int s = 0; s < 8; s++) {
            // A fragments of step s: word s/2 of each row, fields 2*(s%2) and 2*(s%2)+1
            uint32_t a[kMT][4];
#pragma unroll
            for (int mt = 0; mt < kMT; mt++) {
                const uint32_t w0 = (s < 2) ? Acur[mt][0].x : (s < 4) ? Acur[mt][0].y : (s < 6) ? Acur[mt][0].z : Acur[mt][0].w;
                const uint32_t w1 = (s < 2) ? Acur[mt][1].x : (s < 4) ? Acur[mt][1].y : (s < 6) ? Acur[mt][1].z : Acur[mt][1].w;
                const int f = 2 * (s & 1);
                a[mt][0] = fields_to_s8(w0, f);      // row g,   k = 4t + y
                a[mt][1] = fields_to_s8(w1, f);      // row g+8, k = 4t + y
                a[mt][2] = fields_to_s8(w0, f + 1);  // row g,   k = 16 + 4t + y
                a[mt][3] = fields_to_s8(w1, f + 1);  // row g+8, k = 16 + 4t + y
            }
            const uint4 *bs = reinterpret_cast<const uint4 *>(sb + (size_t)s * kStepBytesB);
#pragma unroll
            for (int jj = 0; jj < kNT / 2; jj++) {
                const uint4 b = bs[jj * 32 + lane];   // n-tiles 2jj (b.x,b.y) and 2jj+1 (b.z,b.w)
#pragma unroll
                for (int mt = 0; mt < kMT; mt++) {
                    imma16832(acc[mt][2 * jj], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b.x, b.y);
                    imma16832(acc[mt][2 * jj + 1], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b.z, b.w);
                }
            }
        }
        // release the stage back to the producer
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (it + 1 < nkc) adopt(Acur, Anext);
    }

    // recombine the 7 digit sums per (variant, column) in 64-bit integers: X = hi * 256^3 + lo
#pragma unroll
    for (int mt = 0; mt < kMT; mt++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (!rok[mt][h]) continue;
            const int64_t v = v0 + mt * 16 + h * 8 + g;
#pragma unroll
            for (int cg = 0; cg < 2; cg++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int c = cg * 8 + 2 * t + e;
                    long long hi = 0, lo = 0;
#pragma unroll
                    for (int sl = 0; sl < 4; sl++) hi = hi * 256 + (long long)acc[mt][2 * sl + cg][h * 2 + e];
#pragma unroll
                    for (int sl = 4; sl < kSlices; sl++) lo = lo * 256 + (long long)acc[mt][2 * sl + cg][h * 2 + e];
                    if (use_atomics) {
                        atomicAdd(reinterpret_cast<unsigned long long *>(out_hi + v * 16 + c), (unsigned long long)hi);
                        atomicAdd(reinterpret_cast<unsigned long long *>(out_lo + v * 16 + c), (unsigned long long)lo);
                    } else {
                        out_hi[v * 16 + c] = hi;
                        out_lo[v * 16 + c] = lo;
                    }
                }
        }
}

}  // namespace

int launch_contract_packed(staar_ctx *ctx, const uint8_t *d_packed, size_t stride, int64_t n, int64_t p,
                           long long *d_hi, long long *d_lo)
{
    if (p == 0) return STAAR_OK;
    const NullSparse &ns = ctx->ns;
    if (!ns.sliced) { set_error("packed contraction: sliced operand not built (q > 15?)"); return STAAR_ERR_STATE; }
    const int64_t n_chunks = (n + kChunkSamples - 1) / kChunkSamples;
    const int64_t vt = (p + kVarPerCta - 1) / kVarPerCta;
    // split K only when the variant tiles alone cannot fill the SMs; integer atomics keep it exact and order-free
    int64_t ksplit = 1;
    if (vt < ctx->sm_count) ksplit = std::min<int64_t>(std::max<int64_t>(1, n_chunks / 16), (2 * ctx->sm_count + vt - 1) / vt);
    const int64_t cps = (n_chunks + ksplit - 1) / ksplit;
    ksplit = (n_chunks + cps - 1) / cps;
    if (ksplit > 1) {
        STAAR_CUDA(cudaMemsetAsync(d_hi, 0, sizeof(long long) * (size_t)p * 16, ctx->stream));
        STAAR_CUDA(cudaMemsetAsync(d_lo, 0, sizeof(long long) * (size_t)p * 16, ctx->stream));
    }
    const size_t smem = (size_t)kStages * kChunkBytesB + 2 * kStages * sizeof(uint64_t);
    if (!ctx->attr_contract_packed) {           // per context: the attribute belongs to the context's device
        STAAR_CUDA(cudaFuncSetAttribute(contract_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->attr_contract_packed = true;
    }
    dim3 grid((unsigned)vt, (unsigned)ksplit);
    contract_packed_kernel<<<grid, kWarpsC * 32, smem, ctx->stream>>>(d_packed, stride, p, ns.Yfrag, n_chunks, cps, d_hi,
                                                                      d_lo, ksplit > 1 ? 1 : 0);
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Null-model set-up: slice Y = [r | Sigma_iX] into fragment-major INT8 digits (host, once per null model).
// Layout of one k-chunk (256 samples): [step s = 0..7][n-tile pair jj = 0..6][lane = 0..31][16 bytes], where the
// 16 bytes are {b0,b1} of n-tile 2jj then {b0,b1} of n-tile 2jj+1; n-tile j = 2*slice + colgroup; lane = 4*g + t;
// b0 byte y = digit of column 8*colgroup + g at sample 256*kc + 64*t + 16*(s/2) + 4*y + 2*(s%2); b1: that sample + 1.
// (The sample permutation is free: the kernel's A fragments are built with the same map.)
// ---------------------------------------------------------------------------------------------------------
int build_sliced_operand(staar_ctx *ctx, const double *hY)
{
    NullSparse &ns = ctx->ns;
    const int64_t n = ns.n;
    const int ncol = (int)ns.q + 2, ldY = ns.ldY;      // [r | Sigma_iX | diag(Sigma_i)]
    ns.sliced = false;
    cudaFree(ns.Yfrag); cudaFree(ns.col_scale); cudaFree(ns.col_total); cudaFree(ns.Ytc); cudaFree(ns.Ypair);
    ns.Yfrag = nullptr; ns.col_scale = nullptr; ns.col_total = nullptr; ns.Ytc = nullptr; ns.Ypair = nullptr;
    if (ncol > kColsPerSlice - 1) return STAAR_OK;  // wider designs take the FP64 path (slot 15 is reserved)
    std::vector<double> pulled;
    if (!hY) {                                      // current operand (incl. an updated residual column) from HBM
        pulled.resize((size_t)n * ldY);
        STAAR_CUDA(cudaMemcpyAsync(pulled.data(), ns.Y, sizeof(double) * pulled.size(), cudaMemcpyDeviceToHost, ctx->stream));
        STAAR_CUDA(cudaStreamSynchronize(ctx->stream));
        hY = pulled.data();
        STAAR_TRY(refresh_permuted_Y(ctx, hY));      // the scan's missing-sample sums gather rows of Yp
    }
    // packed columns are in null-model sample order: digit position smp <- row perm[smp] of the natural operand
    const int32_t *hperm = ns.permuted ? ns.h_perm.data() : nullptr;
    auto row_of = [&](int64_t smp) { return hperm ? (int64_t)hperm[smp] : smp; };
    const int64_t n_chunks = (n + kChunkSamples - 1) / kChunkSamples;
    std::vector<int8_t> frag((size_t)n_chunks * kChunkBytesB, 0);
    std::vector<double> scale(kColsPerSlice, 1.0);
    std::vector<int8_t> digits((size_t)n * kSlices);
    std::vector<int8_t> alldigits((size_t)n * kColsPerSlice * kSlices, 0);   // [position][column][slice] for the tcgen05 image
    std::vector<long long> tot(2 * kColsPerSlice, 0);          // exact column totals of the fixed-point operand: hi | lo
    // column kColsPerSlice - 1 is the constant 1 with X = 1 exactly (digits 0,...,0,1): its contraction is the exact
    // sum of a variant's codes, from which the scan gets allele counts without popcounting the column again
    for (int cc = 0; cc <= ncol; cc++) {
        const bool ones = cc == ncol;
        const int c = ones ? kColsPerSlice - 1 : cc;
        double mx = 0.0;
        for (int64_t j = 0; j < n && !ones; j++) {
            const double a = fabs(hY[j * ldY + c]);
            if (!(a <= 1.79e308)) { set_error("null-model operand column %d holds a non-finite value", c); return STAAR_ERR_ARG; }
            if (a > mx) mx = a;
        }
        int e = 0;
        if (mx > 0.0) frexp(mx, &e);                 // mx = f * 2^e, f in [0.5, 1)  =>  |Y| < 2^e
        scale[c] = ones ? 1.0 : ldexp(1.0, e - 54);
        for (int64_t j = 0; j < n; j++) {
            long long X = ones ? 1 : llrint(ldexp(hY[row_of(j) * ldY + c], 54 - e));     // |X| <= 2^54; digits[] is in position order
            int8_t d[kSlices];
            for (int s = kSlices - 1; s >= 1; s--) {
                const long long dd = ((X + 128) & 255) - 128;
                d[s] = (int8_t)dd;
                X = (X - dd) >> 8;
            }
            d[0] = (int8_t)X;                                          // |X| <= 65 here
            for (int s = 0; s < kSlices; s++) {
                digits[(size_t)j * kSlices + s] = d[s];
                alldigits[((size_t)j * kColsPerSlice + c) * kSlices + s] = d[s];
            }
            long long hi = 0, lo = 0;                                  // same split as the kernel: X = hi * 2^24 + lo
            for (int s = 0; s < 4; s++) hi = hi * 256 + d[s];
            for (int s = 4; s < kSlices; s++) lo = lo * 256 + d[s];
            tot[c] += hi; tot[kColsPerSlice + c] += lo;
        }
        const int cg = c / 8, g = c % 8;
        for (int64_t kc = 0; kc < n_chunks; kc++)
            for (int s = 0; s < 8; s++)
                for (int sl = 0; sl < kSlices; sl++) {
                    const int j = 2 * sl + cg, jj = j / 2, within = j % 2;
                    for (int t = 0; t < 4; t++) {
                        int8_t *dst = frag.data() + (size_t)kc * kChunkBytesB + (size_t)s * kStepBytesB +
                                      ((size_t)jj * 32 + (4 * g + t)) * 16 + within * 8;
                        for (int half = 0; half < 2; half++)
                            for (int y = 0; y < 4; y++) {
                                const int64_t smp = kc * 256 + 64 * t + 16 * (s / 2) + 4 * y + 2 * (s % 2) + half;
                                dst[half * 4 + y] = smp < n ? digits[(size_t)smp * kSlices + sl] : (int8_t)0;
                            }
                    }
                }
    }
    STAAR_CUDA(cudaMalloc(&ns.Yfrag, frag.size()));
    STAAR_CUDA(cudaMalloc(&ns.col_scale, sizeof(double) * kColsPerSlice));
    STAAR_CUDA(cudaMalloc(&ns.col_total, sizeof(long long) * 2 * kColsPerSlice));
    ns.reg(&ns.Yfrag, frag.size()); ns.reg(&ns.col_scale, sizeof(double) * kColsPerSlice);
    ns.reg(&ns.col_total, sizeof(long long) * 2 * kColsPerSlice);
    STAAR_CUDA(cudaMemcpyAsync(ns.col_total, tot.data(), sizeof(long long) * tot.size(), cudaMemcpyHostToDevice, ctx->stream));
    STAAR_CUDA(cudaMemcpyAsync(ns.Yfrag, frag.data(), frag.size(), cudaMemcpyHostToDevice, ctx->stream));
    STAAR_CUDA(cudaMemcpyAsync(ns.col_scale, scale.data(), sizeof(double) * kColsPerSlice, cudaMemcpyHostToDevice, ctx->stream));
    STAAR_CUDA(cudaStreamSynchronize(ctx->stream));
    STAAR_TRY(build_tc_operand(ctx, alldigits.data(), kColsPerSlice));
    ns.n_chunks = n_chunks;
    ns.sliced = true;
    return STAAR_OK;
}

}  // namespace staar
