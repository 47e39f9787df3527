// fused_tc.cu — the packed score test in ONE pass over the genotypes (sm_100a): the exact tensor-core contraction of
// contract_tc.cu and the sparse, non-linear terms of scan_packed.cu computed from the SAME shared-memory tiles, so
// every genotype byte is read from HBM exactly once per launch (the reference evaluates both factors from one G:
// src/Individual_Score_Test.cpp:42-43).
//
// A CTA owns 128 * kTiles variants (kTiles = 2: two accumulator tiles share every digit stage, which halves the L2
// traffic of the digit operand per variant — the measured limiter of the one-tile kernel) and walks the samples in
// k-blocks of 256 (8 MMA k-steps):
//   * genotype producer (1 thread): one TMA tensor load per k-block, raw 2-bit tile of kRows x 64 B, 64-byte swizzle,
//     4-stage ring; digit producer (1 thread): one 28 KiB bulk copy of the digit operand per k-block, 3-stage ring;
//   * expander warps (thread = variant row): 2-bit -> int8 in registers, tcgen05.st into the TMEM A stage of their
//     tile; MMA issuer (1 thread): per k-step one tcgen05.mma.kind::i8 (M = 128, N = 112) per tile against the same
//     digit k-step; accumulators (2 x 112 columns) stay in TMEM for the whole column;
//   * scanner warps (half-warp = one variant row at a time, lane = one 16-sample word): from the same tile they
//     find (a) the missing samples — position and word are appended to the variant's list, the finaliser gathers
//     the rows of Y; (b) hom-minor carriers — counted where diag(Sigma_i) is uniform, 2 s_jj gathered elsewhere;
//     (c) 4-sample groups holding two related carriers — one look-up in the group's kinship table.
//     The orientation (flip) is not known while streaming: it is guessed per variant from the first k-block, the
//     finaliser compares the guess with the exact decision and sends the (rare) mismatches, and variants whose
//     missing list overflowed, to the block-per-variant kernel of scan_packed.cu.
// Nothing here is a floating-point atomic: per-(row, lane) partial sums are combined in a fixed order.
#include <climits>
#include "common.cuh"
#include "packed.cuh"
#include "tc_common.cuh"
#include <cuda.h>

namespace staar {

namespace {

using namespace tc;

constexpr int kFKS = 8;                       // k-steps (K = 32 samples) per k-block
constexpr int kFSamples = 32 * kFKS;          // 256 samples per k-block
constexpr int kFRowBytes = kFSamples / 4;     // 64 raw bytes per variant row and k-block
constexpr int kFStepBytes = 2 * 14 * 128;     // digit operand of one k-step (same layout as contract_tc.cu)
constexpr int kFBlockBytesB = kFKS * kFStepBytes;
constexpr int kFStagesB = 3;
constexpr int kFRing = 4;
constexpr int kScanWarps = 8;
constexpr int kFColA = 256;                   // TMEM: accumulators of tile t at column 128 t, A stages from 256
static_assert(kColsPerSlice * kSlices == 112, "N = 112");

template <int kTiles> struct FusedCfg {
    static constexpr int kRows = 128 * kTiles;
    static constexpr int kExpWarps = 4 * kTiles;
    static constexpr int kExpThreads = 128 * kTiles;
    static constexpr int kStagesA = kTiles == 2 ? 2 : 4;
    static constexpr int kTileBytes = kFRowBytes * kRows;
    static constexpr int kThreads = (kExpWarps + kScanWarps + 3) * 32;
    static constexpr int kRowsPerScanWarp = kRows / kScanWarps;
    static constexpr int kSteps = kRowsPerScanWarp / 2;          // a warp scans two rows per step
    static_assert(kFColA + kStagesA * kTiles * 8 * kFKS <= 512, "TMEM budget");
    static_assert(kSteps <= 16, "one count register per lane");
    static constexpr size_t kSmem = (size_t)kFRing * kTileBytes + (size_t)kFStagesB * kFBlockBytesB +
                                    sizeof(double) * kRows * 16 + sizeof(int) * kRows + 64 * sizeof(uint64_t);
};

struct FusedParams {
    int64_t p, n, n_kblocks;
    const int8_t *Ytc;
    long long *out_hi, *out_lo;
    // scanner operands (null-model sample order)
    int64_t tab_words, x_words;
    int uvec;                                  // 64-sample vectors >= uvec: uniform diag(Sigma_i) (INT_MAX: none)
    const uint32_t *kmask;
    const NullSparse::WordDesc *wdesc;
    const NullSparse::KinTerm *xterms;
    const double *lut;
    const double *diag2;                       // 2 * Sigma_i[j,j] by position
    FusedScanOut *scan;
    uint2 *mlist; int mcap;
};

// --------------------------------------------------------------------------------------------------------------
template <int kTiles>
__global__ void __launch_bounds__(FusedCfg<kTiles>::kThreads, 1)
fused_tc_kernel(const __grid_constant__ CUtensorMap gmap, FusedParams P)
{
    using Cfg = FusedCfg<kTiles>;
    constexpr int kRows = Cfg::kRows, kStagesA = Cfg::kStagesA, kTileBytes = Cfg::kTileBytes;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sG = smem;                                                       // kFRing x kTileBytes
    uint8_t *sB = smem + (size_t)kFRing * kTileBytes;                          // kFStagesB x kFBlockBytesB
    double *s_acc = reinterpret_cast<double *>(sB + (size_t)kFStagesB * kFBlockBytesB);   // [kRows][16]
    int *s_cm = reinterpret_cast<int *>(s_acc + kRows * 16);                   // [kRows] missing samples so far
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_cm + kRows);
    uint64_t *a_full = bars, *a_empty = bars + kStagesA;
    uint64_t *b_full = bars + 2 * kStagesA, *b_empty = b_full + kFStagesB;
    uint64_t *g_full = b_empty + kFStagesB, *g_empty = g_full + kFRing;
    uint64_t *d_full = g_empty + kFRing;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(d_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kScanWarp0 = Cfg::kExpWarps, kMmaWarp = kScanWarp0 + kScanWarps, kProdWarp = kMmaWarp + 1;
    const int64_t n_kblocks = P.n_kblocks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStagesA; s++) { bar_init(&a_full[s], Cfg::kExpThreads); bar_init(&a_empty[s], 1); }
        for (int s = 0; s < kFStagesB; s++) { bar_init(&b_full[s], 1); bar_init(&b_empty[s], 1); }
        for (int s = 0; s < kFRing; s++) { bar_init(&g_full[s], 1); bar_init(&g_empty[s], Cfg::kExpThreads + kScanWarps * 32); }
        bar_init(d_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_addr(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < kScanWarp0) {
        // ===================================================================== expanders: thread = variant row
        const int row = threadIdx.x, tile = row >> 7;
        const int64_t v = (int64_t)blockIdx.x * kRows + row;
        const bool rok = v < P.p;
        const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        // K index 4c + y of k-step ks <-> sample 32 ks + 16 (c >> 2) + 4 y + (c & 3) (the digit operand uses the same map)
        uint32_t x[8 * kFKS];
        int ring_slot = 0;
        uint32_t g_par = 0u;
        const uint8_t *my_row = sG + (size_t)row * kFRowBytes;
        const int swz = (row >> 1) & 3;                                      // 64-byte swizzle: chunk c sits at c ^ swz
        auto expand = [&]() {
            bar_wait(&g_full[ring_slot], g_par);
            const uint8_t *rowp = my_row + (size_t)ring_slot * kTileBytes;
            uint32_t w[kFRowBytes / 4];
#pragma unroll
            for (int c = 0; c < kFRowBytes / 16; c++) {
                const uint4 q = *reinterpret_cast<const uint4 *>(rowp + 16 * (c ^ swz));
                w[4 * c] = q.x; w[4 * c + 1] = q.y; w[4 * c + 2] = q.z; w[4 * c + 3] = q.w;
            }
            bar_arrive(&g_empty[ring_slot]);
            if (++ring_slot == kFRing) { ring_slot = 0; g_par ^= 1u; }
#pragma unroll
            for (int c = 0; c < 8 * kFKS; c++)
                x[c] = ((c & 3) == 0 ? w[c >> 2] : __umulhi(w[c >> 2], 1u << (32 - 2 * (c & 3)))) & 0x03030303u;
        };
        if (n_kblocks > 0) expand();
        int sa = 0;
        uint32_t a_par = 1u;
        for (int64_t kb = 0; kb < n_kblocks; kb++) {
            bar_wait(&a_empty[sa], a_par);
            tc_fence_after();
#pragma unroll
            for (int b = 0; b < 2; b++)
                tmem_st32(lane_base + kFColA + (sa * kTiles + tile) * (8 * kFKS) + 32 * b, *reinterpret_cast<const uint32_t (*)[32]>(&x[32 * b]));
            if (kb + 1 < n_kblocks) expand();
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            bar_arrive(&a_full[sa]);
            if (++sa == kStagesA) { sa = 0; a_par ^= 1u; }
        }
        // ------------------------------------------------------------- epilogue: 7 digit sums -> two 64-bit integers
        bar_wait(d_full, 0u);
        tc_fence_after();
        long long hi[16], lo[16];
#pragma unroll
        for (int c = 0; c < 16; c++) { hi[c] = 0; lo[c] = 0; }
#pragma unroll
        for (int sl = 0; sl < kSlices; sl++) {
            int t16[16];
            tmem_ld16(lane_base + 128 * tile + 16 * sl, t16);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 16; c++) {
                if (sl < 4) hi[c] = hi[c] * 256 + (long long)t16[c];
                else lo[c] = lo[c] * 256 + (long long)t16[c];
            }
        }
        if (rok) {
#pragma unroll
            for (int c = 0; c < 16; c++) { P.out_hi[v * 16 + c] = hi[c]; P.out_lo[v * 16 + c] = lo[c]; }
        }
        tc_fence_before();
    } else if (warp < kMmaWarp) {
        // ===================================================================== scanners
        constexpr int kSteps = Cfg::kSteps;
        const int sw = warp - kScanWarp0, half = lane >> 4, l16 = lane & 15;
        const int wrow0 = sw * Cfg::kRowsPerScanWarp;                          // even
        for (int r = lane; r < Cfg::kRowsPerScanWarp * 16; r += 32) s_acc[wrow0 * 16 + r] = 0.0;
        if (lane < Cfg::kRowsPerScanWarp) s_cm[wrow0 + lane] = 0;
        __syncwarp();
        const int64_t v0 = (int64_t)blockIdx.x * kRows;
        uint32_t gmask = 0u;          // bit s: guessed orientation (1 = flip) of the row this half-warp scans at step s
        int cnt = 0;                  // lane 16 h + s: hom-minor carriers in the uniform-diagonal region, row wrow0 + 2 s + h
        int ring_slot = 0;
        uint32_t g_par = 0u;
        const int tab_words = (int)P.tab_words, x_words = (int)P.x_words;
        const int n32 = (int)P.n;
        // byte offset of this lane's word inside a row: 16-byte chunk (l16 >> 2) ^ swz(row), word l16 & 3
        const int wofs = (l16 & 3) * 4;
        for (int64_t kb = 0; kb < n_kblocks; kb++) {
            const int wi = (int)kb * 16 + l16;                                // global 16-sample word index of this lane
            const int left = n32 - wi * 16;                                   // real samples from this word on
            const uint32_t vmask = left >= 16 ? 0x55555555u : (left <= 0 ? 0u : (((1u << (2 * left)) - 1u) & 0x55555555u));
            const bool unif = (wi >> 2) >= P.uvec;
            const uint32_t cmask = unif ? 0xffffffffu : 0u;                   // count (uniform diagonal) vs gather
            const bool kb_kin = (int)kb * 16 < tab_words;                     // warp-uniform: k-block touches table words
            const bool kb_unif = (((int)kb * 16) >> 2) >= P.uvec;             // warp-uniform: every word counts only
            const uint32_t km = wi < tab_words ? __ldg(P.kmask + wi) : 0u;
            const bool xw = wi < x_words;
            const double *dg = P.diag2 + (int64_t)wi * 16;
            const double *lutw = P.lut + ((int64_t)wi << 10);                  // 4 groups x 256 entries per word
            bar_wait(&g_full[ring_slot], g_par);
            const uint8_t *tile = sG + (size_t)ring_slot * kTileBytes;
            if (kb == 0) {
                // orientation guess: more hom-2 than hom-0 codes among the first 256 samples -> the variant will be flipped
#pragma unroll 1
                for (int s = 0; s < kSteps; s++) {
                    const int row = wrow0 + 2 * s + half;
                    const uint32_t w = *reinterpret_cast<const uint32_t *>(tile + row * kFRowBytes + ((((l16 >> 2) ^ ((row >> 1) & 3))) << 4) + wofs);
                    const uint32_t sh = w >> 1;
                    int d = __popc(sh & ~w & vmask) - __popc(~(sh | w) & vmask);
                    d += __shfl_xor_sync(0xffffffffu, d, 8); d += __shfl_xor_sync(0xffffffffu, d, 4);
                    d += __shfl_xor_sync(0xffffffffu, d, 2); d += __shfl_xor_sync(0xffffffffu, d, 1);
                    if (d > 0) gmask |= 1u << s;
                }
            }
#pragma unroll 2
            for (int s = 0; s < kSteps; s++) {
                const int row = wrow0 + 2 * s + half;
                const uint32_t w = *reinterpret_cast<const uint32_t *>(tile + row * kFRowBytes + ((((l16 >> 2) ^ ((row >> 1) & 3))) << 4) + wofs);
                const uint32_t sh = w >> 1;
                const uint32_t mis = w & sh & 0x55555555u;
                const bool fl = (gmask >> s) & 1u;
                const uint32_t gm = fl ? 0xffffffffu : 0u;
                const uint32_t hom = (sh ^ gm) & ~w & vmask;                 // raw code 0 (flip) or 2 (no flip), real samples only
                // uniform-diagonal words: count; both half-warps' sums ride in one warp reduction (<= 256 each)
                const uint32_t c = (uint32_t)__popc(hom & cmask) << (16 * half);
                const uint32_t tot = __reduce_add_sync(0xffffffffu, c);
                if (l16 == s) cnt += (int)((tot >> (16 * half)) & 0xffffu);
                const uint32_t mbal = __ballot_sync(0xffffffffu, mis != 0u);
                if (mbal) {
                    // rare: append (position, word) of every missing sample to the variant's list, ascending positions
                    uint32_t b = mbal;
                    while (b) {
                        const int src = __ffs(b) - 1;
                        b &= b - 1u;
                        uint32_t m = __shfl_sync(0xffffffffu, mis, src);
                        const uint32_t wv = __shfl_sync(0xffffffffu, w, src);
                        if (lane == 0) {
                            const int r = wrow0 + 2 * s + (src >> 4);
                            int cmr = s_cm[r];
                            const uint32_t pos0 = (uint32_t)((int)kb * 16 + (src & 15)) * 16u;
                            uint2 *ml = P.mlist + (size_t)(v0 + r) * P.mcap;
                            while (m) {
                                const int bit = __ffs(m) - 1;
                                m &= m - 1u;
                                if (cmr < P.mcap) ml[cmr] = make_uint2(pos0 + (uint32_t)(bit >> 1), wv);
                                cmr++;
                            }
                            s_cm[r] = cmr;
                        }
                    }
                    __syncwarp();
                }
                if (kb_unif && !kb_kin) continue;                            // nothing to gather in this k-block
                uint32_t homg = hom & ~cmask;
                uint32_t two = 0u, wz = 0u;
                bool cross = false;
                if (kb_kin) {
                    const uint32_t lo = w & 0x55555555u;
                    const uint32_t wf = w ^ ((~lo & gm & 0x55555555u) << 1);  // g -> 2 - g on the codes 0 and 2
                    wz = wf & ~(mis * 3u) & km;                              // related, non-missing codes
                    const uint32_t hit = (wz | (wz >> 1)) & 0x55555555u;
                    two = hit & ((hit | 0x80808080u) - 0x01010101u);         // byte != 0 <=> >= 2 carriers in the group
                    cross = xw && (hit & (hit - 1u)) != 0u;
                }
                if (!__any_sync(0xffffffffu, (homg | two) != 0u || cross)) continue;
                double acc = 0.0;
                while (homg) {                                               // two independent gathers per round
                    const int b0 = __ffs(homg) - 1;
                    homg &= homg - 1u;
                    double a1 = 0.0;
                    const double a0 = __ldg(dg + (b0 >> 1));
                    if (homg) { const int b1 = __ffs(homg) - 1; homg &= homg - 1u; a1 = __ldg(dg + (b1 >> 1)); }
                    acc += a0 + a1;
                }
                if (two) {
                    double t[4];
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        t[g] = 0.0;
                        if (two & (0xffu << (8 * g))) t[g] = __ldg(lutw + (g << 8) + ((wz >> (8 * g)) & 0xffu));
                    }
                    acc += (t[0] + t[1]) + (t[2] + t[3]);
                }
                if (cross) {                                                 // families of 5..16 members: member x other chunk
                    const uint2 wd = __ldg(reinterpret_cast<const uint2 *>(P.wdesc) + wi);
                    const uint2 *tp = reinterpret_cast<const uint2 *>(P.xterms) + wd.x;
                    for (uint32_t t = 0; t < wd.y; t++) {
                        const uint2 tr = __ldg(tp + t);
                        const uint32_t idx = (wz >> (tr.y & 0xffu)) & ((tr.y >> 8) & 0xffu);
                        const uint32_t gj = (wz >> ((tr.y >> 16) & 0xffu)) & 3u;
                        if (gj != 0u && idx != 0u) acc += (double)gj * __ldg(P.lut + tr.x + idx);
                    }
                }
                if (acc != 0.0) s_acc[row * 16 + l16] += acc;
            }
            __syncwarp();
            bar_arrive(&g_empty[ring_slot]);
            if (++ring_slot == kFRing) { ring_slot = 0; g_par ^= 1u; }
        }
        // ---- per-row totals: the 16 word lanes in a fixed tree
        __syncwarp();
#pragma unroll 1
        for (int s = 0; s < kSteps; s++) {
            const int row = wrow0 + 2 * s + half;
            double t = s_acc[row * 16 + l16];
            t += __shfl_xor_sync(0xffffffffu, t, 8); t += __shfl_xor_sync(0xffffffffu, t, 4);
            t += __shfl_xor_sync(0xffffffffu, t, 2); t += __shfl_xor_sync(0xffffffffu, t, 1);
            const int c = __shfl_sync(0xffffffffu, cnt, 16 * half + s);
            if (l16 == 0 && v0 + row < P.p) {
                FusedScanOut o;
                o.acc = t; o.cnt = c; o.cm = s_cm[row]; o.guess = (int)((gmask >> s) & 1u); o.pad = 0;
                P.scan[v0 + row] = o;
            }
        }
    } else if (warp == kMmaWarp) {
        // ===================================================================== MMA issuer (one thread)
        if (lane == 0) {
            constexpr uint32_t kIdesc = idesc_i8(128, 112);
            int sa = 0, sb = 0;
            uint32_t a_par = 0u, b_par = 0u;
            for (int64_t kb = 0; kb < n_kblocks; kb++) {
                bar_wait(&b_full[sb], b_par);
                bar_wait(&a_full[sa], a_par);
                tc_fence_after();
                const uint64_t bdesc = smem_desc(s_addr(sB + (size_t)sb * kFBlockBytesB), 14 * 128, 128);
#pragma unroll
                for (int ks = 0; ks < kFKS; ks++)
#pragma unroll
                    for (int t = 0; t < kTiles; t++)
                        umma_i8_ts(tmem + 128 * t, tmem + kFColA + (sa * kTiles + t) * (8 * kFKS) + 8 * ks,
                                   bdesc + (uint64_t)(ks * (kFStepBytes >> 4)), kIdesc, (kb > 0 || ks > 0) ? 1u : 0u);
                tc_commit(&a_empty[sa]);
                tc_commit(&b_empty[sb]);
                if (++sa == kStagesA) { sa = 0; a_par ^= 1u; }
                if (++sb == kFStagesB) { sb = 0; b_par ^= 1u; }
            }
            tc_commit(d_full);
        }
        __syncwarp();
    } else if (warp == kProdWarp) {
        // ===================================================================== digit producer (one thread)
        if (lane == 0) {
            int sb = 0;
            uint32_t b_par = 1u;
            for (int64_t kb = 0; kb < n_kblocks; kb++) {
                bar_wait_relaxed(&b_empty[sb], b_par);
                bar_expect_tx(&b_full[sb], kFBlockBytesB);
                bulk_load(sB + (size_t)sb * kFBlockBytesB, P.Ytc + (size_t)kb * kFBlockBytesB, kFBlockBytesB, &b_full[sb]);
                if (++sb == kFStagesB) { sb = 0; b_par ^= 1u; }
            }
        }
        __syncwarp();
    } else {
        // ===================================================================== genotype producer (one thread)
        if (lane == 0) {
            int sg = 0;
            uint32_t g_par = 1u;
            const int32_t row0 = (int32_t)(blockIdx.x * kRows);
            for (int64_t kb = 0; kb < n_kblocks; kb++) {
                bar_wait(&g_empty[sg], g_par);
                bar_expect_tx(&g_full[sg], kTileBytes);
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(s_addr(sG + (size_t)sg * kTileBytes)), "l"(&gmap), "r"((int32_t)(kb * kFRowBytes)), "r"(row0),
                               "r"(s_addr(&g_full[sg])) : "memory");
                if (++sg == kFRing) { sg = 0; g_par ^= 1u; }
            }
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// --------------------------------------------------------------------------------------------------------------
// Finaliser of the fused path: one warp per variant.  Exact flip / AF / MAF from integer counts (sum of codes from
// the contraction's constant-1 column, missing count from the scanners), column sums of Y over the listed missing
// samples (lane = column, two list entries per instruction), kinship correction for missing related samples from the
// listed words, then the statistics of src/Individual_Score_Test.cpp:45-64.  A variant whose orientation guess was
// wrong or whose missing list overflowed goes to the redo list instead.
struct FusedFinParams {
    int64_t p, n; int q, impute;
    const long long *hi, *lo, *tot;
    const double *scale, *cov;
    const FusedScanOut *scan;
    const uint2 *mlist; int mcap;
    const double *Y; int ldY;
    int64_t tab_words;
    const uint32_t *kmask;
    const NullSparse::SampleInfo *sinfo;
    const NullSparse::OffEntry *off_ent;
    double dval;
    double *AF, *MAF, *Score, *Score_se, *pl, *Est, *Est_se;
    long long *redo; unsigned long long *redo_count;
};

__global__ void __launch_bounds__(256) fused_finalize_kernel(FusedFinParams P)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int c16 = lane & 15, hw = lane >> 4;
    for (int64_t i = warp0; i < P.p; i += nwarps) {
        const FusedScanOut so = P.scan[i];
        const long long S = P.lo[i * 16 + 15];
        const int cm = so.cm;
        const long long gsum = S - 3ll * cm, gnum = (long long)P.n - cm;
        bool flip, mu_zero;
        if (P.impute == STAAR_IMPUTE_MEAN) {
            flip = gsum > gnum;
            mu_zero = flip ? (gsum == 2 * gnum) : (gsum == 0);
        } else {
            const long long fill = gsum > gnum ? 2 : 0;
            flip = gsum + fill * cm > (long long)P.n;
            mu_zero = flip ? (fill == 2) : (fill == 0);
        }
        const bool mono = gsum == 0 || gsum == 2 * gnum;                 // every genotype 0 after flip and imputation
        if (!mono && (cm > P.mcap || (int)flip != so.guess)) {
            if (lane == 0) P.redo[atomicAdd(P.redo_count, 1ull)] = i;
            continue;
        }
        const PackedCounts pc{(int)gsum, 0, cm};
        const VariantFlip vf = flip_from_counts(pc, P.n, P.impute);
        if (mono) {
            if (lane == 0) {
                P.AF[i] = vf.af; P.MAF[i] = vf.maf; P.Score[i] = 0.0;
                single_trait_stats(0.0, 0.0, P.Score_se[i], P.pl[i], P.Est[i], P.Est_se[i]);
            }
            continue;
        }
        const double mu = vf.mu;
        // ---- missing samples: column sums of Y (lane = column, half-warp h takes entries h, h + 2, ...)
        const uint2 *ml = P.mlist + (size_t)i * P.mcap;
        double sm = 0.0;
        if (c16 <= P.q + 1)
            for (int e = hw; e < cm; e += 2) sm += __ldg(P.Y + (int64_t)ml[e].x * P.ldY + c16);
        sm += __shfl_xor_sync(0xffffffffu, sm, 16);
        // ---- missing related samples: imputed to mu, the tables saw them as 0
        double corr = 0.0;
        if (!mu_zero) {
            for (int e = lane; e < cm; e += 32) {
                const uint2 en = ml[e];
                const int64_t j = en.x;
                if ((int64_t)(j >> 4) >= P.tab_words) continue;
                if (((__ldg(P.kmask + (j >> 4)) >> (2 * (j & 15))) & 3u) == 0u) continue;
                const NullSparse::SampleInfo si = P.sinfo[j];
                double sacc = 0.0;
                for (uint32_t t = 0; t < si.off_count; t++) {
                    const NullSparse::OffEntry oe = P.off_ent[si.off_start + t];
                    uint32_t code = (en.y >> (2 * (oe.col & 15))) & 3u;         // relatives of a table family share the word
                    if (flip && !(code & 1u)) code ^= 2u;
                    const double gv = code == 3u ? ((int64_t)oe.col > j ? mu : 0.0) : (double)code;
                    sacc += oe.val * gv;
                }
                corr += 2.0 * mu * sacc;
            }
        }
        corr = warp_sum(corr);
        // ---- G'[r | Sigma_iX | diag]: lane c < q + 2
        const int q = P.q;
        double L = 0.0;
        if (c16 <= q + 1) {
            long long h = P.hi[i * 16 + c16], l = P.lo[i * 16 + c16];
            if (flip) { h = 2 * P.tot[c16] - h; l = 2 * P.tot[kColsPerSlice + c16] - l; }
            L = ((double)h * 16777216.0 + (double)l) * P.scale[c16];
        }
        const double wm = flip ? mu + 1.0 : mu - 3.0;
        const double Lc = L + wm * sm;                                     // columns 0..q
        double xb = 0.0;                                                  // lane b < q: sum_a L[1 + a] cov[a, b]
        for (int a = 0; a < q; a++) {
            const double La = __shfl_sync(0xffffffffu, Lc, 1 + a);
            if (lane < q) xb += La * P.cov[a + lane * q];
        }
        const double Lb = __shfl_sync(0xffffffffu, Lc, (lane + 1) & 31);
        double tct = lane < q ? xb * Lb : 0.0;
        tct = warp_sum(tct);
        const double U = __shfl_sync(0xffffffffu, Lc, 0);
        const double smd = __shfl_sync(0xffffffffu, sm, q + 1);
        const double lin = __shfl_sync(0xffffffffu, L, q + 1) + (flip ? 1.0 : -3.0) * smd;
        if (lane == 0) {
            const double quad = lin + (so.acc + 2.0 * P.dval * (double)so.cnt) + mu * mu * smd + corr;
            const double C = quad - tct;
            P.AF[i] = vf.af; P.MAF[i] = vf.maf; P.Score[i] = U;
            single_trait_stats(U, C, P.Score_se[i], P.pl[i], P.Est[i], P.Est_se[i]);
        }
    }
}

int encode_tile_map(CUtensorMap *map, const uint8_t *d_packed, size_t stride, int64_t p, int rows)
{
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    STAAR_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled is not available in this driver"); return STAAR_ERR_CUDA; }
    const EncodeFn encode = reinterpret_cast<EncodeFn>(fn);
    // rank-2 byte tensor: dim 0 = bytes of a column, dim 1 = variants; box = 64 bytes x rows; zero fill outside
    const cuuint64_t dims[2] = {(cuuint64_t)stride, (cuuint64_t)p};
    const cuuint64_t strides[1] = {(cuuint64_t)stride};
    const cuuint32_t box[2] = {(cuuint32_t)kFRowBytes, (cuuint32_t)rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t *>(d_packed), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (CUresult %d): packed base must be 16-byte aligned", (int)r); return STAAR_ERR_CUDA; }
    return STAAR_OK;
}

}  // namespace

int fused_missing_cap(int64_t n)
{
    int cap = 64;
    while ((int64_t)cap * 400 < n) cap *= 2;
    return cap;
}

bool fused_path_available(const staar_ctx *ctx)
{
    const NullSparse &ns = ctx->ns;
    // families of more than 16 members need relatives' genotypes from anywhere in the column (scan_packed.cu keeps the
    // whole column in shared memory); every other null model streams
    return ns.set && ns.sliced && ns.Ytc && ns.diag2_p && ns.n_generic == 0 && ns.kin_words == ns.tab_words;
}

int launch_fused_tc(staar_ctx *ctx, const uint8_t *d_packed, size_t stride, int64_t n, int64_t p, long long *d_hi,
                    long long *d_lo, FusedScanOut *d_scan, uint2 *d_mlist, int mcap)
{
    if (p == 0) return STAAR_OK;
    const NullSparse &ns = ctx->ns;
    FusedParams P;
    P.p = p; P.n = n; P.n_kblocks = (n + kFSamples - 1) / kFSamples;
    P.Ytc = ns.Ytc; P.out_hi = d_hi; P.out_lo = d_lo;
    P.tab_words = ns.tab_words; P.x_words = ns.x_words;
    P.uvec = ns.diag_uniform_vec >= 0 ? (int)ns.diag_uniform_vec : INT_MAX;
    P.kmask = ns.kmask; P.wdesc = ns.wdesc; P.xterms = ns.xterms; P.lut = ns.lut; P.diag2 = ns.diag2_p;
    P.scan = d_scan; P.mlist = d_mlist; P.mcap = mcap;
    const int tiles = ctx->fused_tiles == 1 ? 1 : 2;
    CUtensorMap gmap;
    STAAR_TRY(encode_tile_map(&gmap, d_packed, stride, p, 128 * tiles));
    if (tiles == 2) {
        using Cfg = FusedCfg<2>;
        STAAR_CUDA(cudaFuncSetAttribute(fused_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem));
        fused_tc_kernel<2><<<(unsigned)((p + Cfg::kRows - 1) / Cfg::kRows), Cfg::kThreads, Cfg::kSmem, ctx->stream>>>(gmap, P);
    } else {
        using Cfg = FusedCfg<1>;
        STAAR_CUDA(cudaFuncSetAttribute(fused_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem));
        fused_tc_kernel<1><<<(unsigned)((p + Cfg::kRows - 1) / Cfg::kRows), Cfg::kThreads, Cfg::kSmem, ctx->stream>>>(gmap, P);
    }
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

int launch_fused_finalize(staar_ctx *ctx, int64_t n, int64_t p, int impute, const long long *d_hi, const long long *d_lo,
                          const FusedScanOut *d_scan, const uint2 *d_mlist, int mcap, double *dAF, double *dMAF,
                          double *dScore, double *dScore_se, double *dpl, double *dEst, double *dEst_se, long long *d_redo,
                          unsigned long long *d_redo_count)
{
    if (p == 0) return STAAR_OK;
    const NullSparse &ns = ctx->ns;
    FusedFinParams P;
    P.p = p; P.n = n; P.q = (int)ns.q; P.impute = impute; P.hi = d_hi; P.lo = d_lo; P.tot = ns.col_total;
    P.scale = ns.col_scale; P.cov = ns.cov; P.scan = d_scan; P.mlist = d_mlist; P.mcap = mcap; P.Y = ns.Yp; P.ldY = ns.ldY;
    P.tab_words = ns.tab_words; P.kmask = ns.kmask; P.sinfo = ns.sinfo_p; P.off_ent = ns.off_ent_p;
    P.dval = ns.diag_uniform_val;
    P.AF = dAF; P.MAF = dMAF; P.Score = dScore; P.Score_se = dScore_se; P.pl = dpl; P.Est = dEst; P.Est_se = dEst_se;
    P.redo = d_redo; P.redo_count = d_redo_count;
    const int blocks = (int)std::min<int64_t>((p + 7) / 8, (int64_t)ctx->sm_count * 8);
    fused_finalize_kernel<<<blocks, 256, 0, ctx->stream>>>(P);
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

}  // namespace staar
