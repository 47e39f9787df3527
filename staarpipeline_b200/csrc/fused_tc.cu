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
#include <type_traits>
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
constexpr int kScanWarps = 16;
constexpr int kListCap = 1024;                // cells listed per k-block and CTA (the rest is served by the lane that found it)
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
                                    sizeof(double) * kRows * 16 + 2 * kListCap * sizeof(uint2) + sizeof(double) * kScanWarps * 4 * 32 +
                                    sizeof(uint16_t) * kRows * 16 + kRows + 16 +
                                    32 * 8 * sizeof(unsigned int) + 64 * sizeof(uint64_t);
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
    const double *gtab;                        // kinship tables, then 2 * Sigma_i[j,j] by position from diag2_off
    uint32_t diag2_off;
    FusedScanOut *scan;
    uint2 *mlist; int msub;                    // missing samples: 16 sub-lists of msub entries per variant (position, word)
    uint16_t *msub_cnt;                        // p x 16 entries found per sub-list (may exceed msub: then the variant is redone)
    int dbg;                                   // STAAR_B200_FUSED_DBG: switch-off experiments (timing only; results invalid)
    unsigned int *tim;                         // dbg & 128: per CTA and warp, clocks spent per phase (8 counters), else NULL
};

// phase timers (dbg & 128): lane 0 of a warp adds the clocks since the previous mark to counter `ph`
struct PhaseTimer {
    unsigned int *slot; unsigned int last; bool on;
    __device__ __forceinline__ void start(unsigned int *s_tim, int warp, bool enable)
    {
        on = enable && (threadIdx.x & 31) == 0; slot = s_tim + warp * 8;
        if (on) { for (int j = 0; j < 8; j++) slot[j] = 0u; last = (unsigned int)clock64(); }
    }
    __device__ __forceinline__ void mark(int ph)
    {
        if (on) { const unsigned int now = (unsigned int)clock64(); slot[ph] += now - last; last = now; }
    }
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
    uint2 *s_cl = reinterpret_cast<uint2 *>(s_acc + kRows * 16);              // [2][kListCap] flagged cells of a k-block (by k-block parity)
    double *s_stage = reinterpret_cast<double *>(s_cl + 2 * kListCap);        // [kScanWarps][4][32] staged look-ups
    uint16_t *s_mc = reinterpret_cast<uint16_t *>(s_stage + kScanWarps * 4 * 32);   // [kRows][16] missing samples per sub-list
    uint8_t *s_guess = reinterpret_cast<uint8_t *>(s_mc + kRows * 16);        // [kRows] orientation guesses
    int *s_ccount = reinterpret_cast<int *>(s_guess + kRows);                 // [4] cells listed for k-block kb % 3
    unsigned int *s_tim = reinterpret_cast<unsigned int *>(s_ccount + 4);     // [32 warps][8] phase clocks (dbg & 128)
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_tim + 32 * 8);
    uint64_t *a_full = bars, *a_empty = bars + kStagesA;
    uint64_t *b_full = bars + 2 * kStagesA, *b_empty = b_full + kFStagesB;
    uint64_t *g_full = b_empty + kFStagesB, *g_empty = g_full + kFRing;
    uint64_t *d_full = g_empty + kFRing;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(d_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kScanWarp0 = Cfg::kExpWarps, kMmaWarp = kScanWarp0 + kScanWarps, kProdWarp = kMmaWarp + 1;
    const int64_t n_kblocks = P.n_kblocks;

    if (threadIdx.x == 0) {
        for (int j = 0; j < 4; j++) s_ccount[j] = 0;
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
#ifdef FUSED_TIMERS
    PhaseTimer tm;
    tm.start(s_tim, warp, (P.dbg & 128) != 0);
#define TM_MARK(ph) tm.mark(ph)
#else
#define TM_MARK(ph) ((void)0)
#endif

    if (warp < kScanWarp0) {
        // ===================================================================== expanders: thread = variant row
        const int row = threadIdx.x, tile = row >> 7;
        const int64_t v = (int64_t)blockIdx.x * kRows + row;
        const bool rok = v < P.p;
        const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        // K index 4c + y of k-step ks <-> sample 32 ks + 16 (c >> 2) + 4 y + (c & 3) (the digit operand uses the same map)
        uint32_t w[kFRowBytes / 4];                                          // raw words of the k-block about to be expanded
        int ring_slot = 0;
        uint32_t g_par = 0u;
        const uint8_t *my_row = sG + (size_t)row * kFRowBytes;
        const int swz = (row >> 1) & 3;                                      // 64-byte swizzle: chunk c sits at c ^ swz
        auto fetch = [&]() {
            TM_MARK(2);
            bar_wait_hint(&g_full[ring_slot], g_par);
            TM_MARK(1);
            const uint8_t *rowp = my_row + (size_t)ring_slot * kTileBytes;
#pragma unroll
            for (int c = 0; c < kFRowBytes / 16; c++) {
                const uint4 q = *reinterpret_cast<const uint4 *>(rowp + 16 * (c ^ swz));
                w[4 * c] = q.x; w[4 * c + 1] = q.y; w[4 * c + 2] = q.z; w[4 * c + 3] = q.w;
            }
            bar_arrive(&g_empty[ring_slot]);
            if (++ring_slot == kFRing) { ring_slot = 0; g_par ^= 1u; }
        };
        if (n_kblocks > 0) fetch();
        int sa = 0;
        uint32_t a_par = 1u;
        for (int64_t kb = 0; kb < n_kblocks; kb++) {
            TM_MARK(2);
            bar_wait_hint(&a_empty[sa], a_par);
            TM_MARK(0);
            tc_fence_after();
            if (!(P.dbg & 8)) {
#pragma unroll
                for (int b = 0; b < 2; b++) {                                // 32 TMEM columns (4 k-steps) at a time: 32 registers
                    uint32_t x[32];
#pragma unroll
                    for (int c = 0; c < 32; c++) {
                        const uint32_t ww = w[8 * b + (c >> 2)];
                        x[c] = ((c & 3) == 0 ? ww : __umulhi(ww, 1u << (32 - 2 * (c & 3)))) & 0x03030303u;
                    }
                    tmem_st32(lane_base + kFColA + (sa * kTiles + tile) * (8 * kFKS) + 32 * b, x);
                }
            }
            if (kb + 1 < n_kblocks) fetch();                                 // overlaps the store latency
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            bar_arrive(&a_full[sa]);
            if (++sa == kStagesA) { sa = 0; a_par ^= 1u; }
        }
        // ------------------------------------------------------------- epilogue: 7 digit sums -> two 64-bit integers
        bar_wait(d_full, 0u);
        tc_fence_after();
#pragma unroll
        for (int h8 = 0; h8 < 2; h8++) {                                     // eight columns at a time (register budget)
            long long hi[8], lo[8];
#pragma unroll
            for (int c = 0; c < 8; c++) { hi[c] = 0; lo[c] = 0; }
#pragma unroll
            for (int sl = 0; sl < kSlices; sl++) {
                int t8[8];
                tmem_ld8(lane_base + 128 * tile + 16 * sl + 8 * h8, t8);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    if (sl < 4) hi[c] = hi[c] * 256 + (long long)t8[c];
                    else lo[c] = lo[c] * 256 + (long long)t8[c];
                }
            }
            if (rok) {
#pragma unroll
                for (int c = 0; c < 8; c++) { P.out_hi[v * 16 + 8 * h8 + c] = hi[c]; P.out_lo[v * 16 + 8 * h8 + c] = lo[c]; }
            }
        }
        tc_fence_before();
    } else if (warp < kMmaWarp) {
        // ===================================================================== scanners
        // Lane (half, l16) of scanner warp sw owns, for every step s, the word l16 of row wrow0 + 2 s + half.  Phase 1 of a
        // k-block is a branch-free unrolled sweep that only counts hom-minor carriers (uniform-diagonal words) and flags
        // the cells = (row, word) pairs that need more: missing samples, table look-ups.  Cells are not served where they
        // are found — a lane with a hit would run alone while 31 wait, and a warp that owns several common variants would
        // hold the whole CTA back: every scanner warp appends its flagged cells (with the post-flip word) to ONE list per
        // CTA and k-block (shared memory), and after a barrier among the scanner warps the list is served 32 cells at a
        // time, batches dealt round-robin to the warps.  Look-ups are cp.async copies into a per-lane staging area (a
        // load outstanding in a register would keep its scoreboard slot and stall the next shared-memory read); those of a
        // warp's last batch stay pending and are added to the cell's accumulator slot during the next k-block.  A cell's
        // slot s_acc[row][word] receives one value per k-block, in k-block order, whichever warp serves it; missing
        // samples go to sub-list `word` of the row's variant in k-block order: results do not depend on the schedule.
        constexpr int kSteps = Cfg::kSteps;
        const int sw = warp - kScanWarp0, half = lane >> 4, l16 = lane & 15;
        const int wrow0 = sw * Cfg::kRowsPerScanWarp;                          // even
        for (int r = lane; r < Cfg::kRowsPerScanWarp * 16; r += 32) { s_acc[wrow0 * 16 + r] = 0.0; s_mc[wrow0 * 16 + r] = 0; }
        __syncwarp();
        const int64_t v0 = (int64_t)blockIdx.x * kRows;
        uint32_t gmask = 0u;          // bit s: guessed orientation (1 = flip) of the row this half-warp scans at step s
        int cnt_u[kSteps];            // this lane's hom-minor count (uniform-diagonal words) of the row scanned at step s
#pragma unroll
        for (int s = 0; s < kSteps; s++) cnt_u[s] = 0;
        double *stage = s_stage + (sw * 4) * 32 + lane;                      // look-up j of this lane: stage[32 j]
        int pslot = -1;
        auto consume = [&]() {
            if (pslot >= 0) {
                asm volatile("cp.async.wait_all;" ::: "memory");
                s_acc[pslot] += (stage[0] + stage[32]) + (stage[64] + stage[96]);
                pslot = -1;
            }
        };
        int ring_slot = 0;
        uint32_t g_par = 0u;
        const int tab_words = (int)P.tab_words, x_words = (int)P.x_words, uvec4 = P.uvec > (INT_MAX >> 2) ? INT_MAX : 4 * P.uvec;
        const int n32 = (int)P.n, nkb = (int)n_kblocks;
        const int msub = P.msub;
        uint32_t km_next = l16 < tab_words ? __ldg(P.kmask + l16) : 0u;
        // byte offset of this lane's word inside a row: 16-byte chunk (l16 >> 2) ^ swz(row), word l16 & 3
        const int wofs = (l16 & 3) * 4, chunk = l16 >> 2;
        const int swz0 = wrow0 >> 1;                                          // (row >> 1) & 3 == (swz0 + s) & 3
        const uint32_t rbase = (uint32_t)(((wrow0 + half) << 4) | l16);       // cell id = row << 4 | word; step s adds 2 rows
        int par3 = 0;                                                         // kb % 3
        for (int kb = 0; kb < nkb; kb++) {
            const int wi = kb * 16 + l16;                                     // global 16-sample word index of this lane
            uint32_t vmask = 0x55555555u;                                     // real samples of the word (tail of the column)
            if (kb == nkb - 1) {
                const int left = n32 - wi * 16;
                vmask = left >= 16 ? 0x55555555u : (left <= 0 ? 0u : (((1u << (2 * left)) - 1u) & 0x55555555u));
            }
            const bool kb_kin = kb * 16 < tab_words;                          // warp-uniform: k-block touches table words
            const bool kb_unif = kb * 16 >= uvec4;                            // warp-uniform: every word counts only
            const bool kb_nouni = kb * 16 + 15 < uvec4;                       // warp-uniform: no word counts
            const uint32_t cmask = wi >= uvec4 ? 0xffffffffu : 0u;            // count (uniform diagonal) vs look up
            const uint32_t km = km_next;                                      // fetched one k-block ahead
            km_next = wi + 16 < tab_words ? __ldg(P.kmask + wi + 16) : 0u;
            const bool xw = wi < x_words;
            TM_MARK(6);
            bar_wait_hint(&g_full[ring_slot], g_par);
            TM_MARK(0);
            if (P.dbg & 1) { bar_arrive(&g_empty[ring_slot]); if (++ring_slot == kFRing) { ring_slot = 0; g_par ^= 1u; } continue; }
            const uint8_t *lane_tile = sG + (size_t)ring_slot * kTileBytes + (wrow0 + half) * kFRowBytes + wofs;
            auto word_at = [&](int s) {
                return *reinterpret_cast<const uint32_t *>(lane_tile + s * (2 * kFRowBytes) + ((chunk ^ ((swz0 + s) & 3)) << 4));
            };
            if (kb == 0) {
                // orientation guess: more hom-2 than hom-0 codes among the first 256 samples -> the variant will be flipped
#pragma unroll 1
                for (int s = 0; s < kSteps; s++) {
                    const uint32_t w = word_at(s);
                    const uint32_t sh = w >> 1;
                    int d = __popc(sh & ~w & vmask) - __popc(~(sh | w) & vmask);
                    d += __shfl_xor_sync(0xffffffffu, d, 8); d += __shfl_xor_sync(0xffffffffu, d, 4);
                    d += __shfl_xor_sync(0xffffffffu, d, 2); d += __shfl_xor_sync(0xffffffffu, d, 1);
                    if (d > 0) gmask |= 1u << s;
                }
            }
            // ---- phase 1: branch-free sweep over the steps; ev = steps at which this lane's cell needs serving.
            // Shifts are written as multiplications (FMA pipe): the sweep is bound by the ALU pipe (LOP3, SEL).
            uint32_t ev = 0u;
            auto sweep = [&](auto kin_c, auto count_c, auto gather_c) {
                constexpr bool kKin = decltype(kin_c)::value, kCount = decltype(count_c)::value, kGather = decltype(gather_c)::value;
#pragma unroll
                for (int s = 0; s < kSteps; s++) {
                    const uint32_t w = word_at(s);
                    const uint32_t gm = __umulhi(gmask << (31 - s), 2u) * 0xffffffffu;   // all ones: this row is flipped
                    const uint32_t wf = w ^ (((~w & 0x55555555u) * 2u) & gm);            // g -> 2 - g on the codes 0 and 2
                    const uint32_t fh = __umulhi(wf, 0x80000000u);                       // wf >> 1
                    const uint32_t mis = wf & fh & 0x55555555u;
                    const uint32_t hom = fh & ~wf & vmask;                   // hom-minor under the guess, real samples only
                    if (kCount) cnt_u[s] += __popc(kGather ? (hom & cmask) : hom);
                    uint32_t e = mis | (kGather ? (hom & ~cmask) : 0u);
                    if (kKin) {
                        const uint32_t wz = wf & ~(mis * 3u) & km;           // related, non-missing codes
                        const uint32_t hit = (wz | __umulhi(wz, 0x80000000u)) & 0x55555555u;
                        // >= 2 carriers in a 4-sample group (byte != 0), in a cross word: >= 2 anywhere
                        e |= xw ? (hit & (hit - 1u)) : (hit & ((hit | 0x80808080u) - 0x01010101u));
                    }
                    ev |= (e != 0u ? 1u : 0u) << s;
                }
            };
            using T = std::true_type;
            using F = std::false_type;
            if (kb_kin) { if (kb_nouni) sweep(T{}, F{}, T{}); else sweep(T{}, T{}, T{}); }
            else if (kb_unif) sweep(F{}, T{}, F{});
            else sweep(F{}, T{}, T{});
            TM_MARK(1);
            consume();                                                       // look-ups issued in the previous k-block
            // ---- phase 2: list the flagged cells of the CTA, then serve them in dense batches
            if (P.dbg & 4) ev = 0u;
            uint2 *clist = s_cl + (kb & 1) * kListCap;
            // serve one cell (post-flip word wf of 16-sample word wi_c of row `row`, that word's constants): missing samples
            // -> the variant's sub-list; the first four look-ups -> asynchronous copies into the lane's staging slots
            // (unused slots zeroed); the rare rest — more than four, cross terms of families of 5..16 members — is loaded on
            // the spot and returned
            auto serve = [&](uint32_t slot, uint32_t wf, uint32_t km_c, uint32_t vm_c, uint32_t cm_c, int wi_c) -> double {
                const uint32_t fh = wf >> 1;
                uint32_t mis = wf & fh & 0x55555555u;
                uint32_t h = fh & ~wf & vm_c & ~cm_c;                         // hom-minor carriers outside the uniform-diagonal region
                const uint32_t wa = wf & ~(mis * 3u);                        // all non-missing codes (table index)
                if (mis && !(P.dbg & 2)) {
                    int k = s_mc[slot];
                    // entry k of sub-list `word` of a variant lives at [k][word]: the first entries of all 16 sub-lists
                    // share a 128-byte line, so a variant's list is a few full lines for the finaliser
                    uint2 *ml = P.mlist + (size_t)(v0 + (slot >> 4)) * 16 * msub + (slot & 15u);
                    do {
                        const int bit = __ffs(mis) - 1;
                        mis &= mis - 1u;
                        if (k < msub) ml[16 * k] = make_uint2((uint32_t)wi_c * 16u + (uint32_t)(bit >> 1), wf);
                        k++;
                    } while (mis);
                    s_mc[slot] = (uint16_t)k;
                }
                const uint32_t wz = wa & km_c;                               // related ones
                const uint32_t hit = (wz | (wz >> 1)) & 0x55555555u;
                const uint32_t two = hit & ((hit | 0x80808080u) - 0x01010101u);
                // bit 8 g + 7: group g needs its table — two related carriers, or (table words: the table also
                // carries 2 s_jj [g_j == 2]) any hom-minor carrier
                uint32_t z = (two + 0x7f7f7f7fu) & 0x80808080u;
                if (wi_c < tab_words) { z |= (h + 0x7f7f7f7fu) & 0x80808080u; h = 0u; }
                const double *lup = P.gtab + ((int64_t)wi_c << 10);           // the word's 4 group tables
                const double *dgp = P.gtab + P.diag2_off + (int64_t)wi_c * 16;   // 2 * Sigma_i[j,j] of the word's samples
                uint32_t sdst = s_addr(stage);
                int used = 0;
                double extra = 0.0;
                while (z) {
                    const int g8 = (__ffs(z) - 1) & ~7;                       // 8 * group
                    z &= z - 1u;
                    if (!(P.dbg & 256)) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sdst), "l"(lup + (g8 << 5) + ((wa >> g8) & 0xffu)) : "memory");
                    else stage[32 * used] = (double)((wa >> g8) & 0xffu);
                    sdst += 256; used++;
                }
                while (h) {
                    const int b = __ffs(h) - 1;
                    h &= h - 1u;
                    if (used < 4) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sdst), "l"(dgp + (b >> 1)) : "memory"); sdst += 256; used++; }
                    else extra += __ldg(dgp + (b >> 1));                     // more than four look-ups in one word
                }
                for (int j = used; j < 4; j++) stage[32 * j] = 0.0;
                if (!(P.dbg & 32) && wi_c < x_words && (hit & (hit - 1u)) != 0u) {   // families of 5..16 members: member x other chunk
                    const uint2 wd = __ldg(reinterpret_cast<const uint2 *>(P.wdesc) + wi_c);
                    const uint2 *tp = reinterpret_cast<const uint2 *>(P.xterms) + wd.x;
                    for (uint32_t t = 0; t < wd.y; t += 4) {                 // four terms in flight (descriptors are L1 hits)
                        double xv[4];
                        uint32_t gjv[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            xv[u] = 0.0; gjv[u] = 0u;
                            if (t + u < wd.y) {
                                const uint2 tr = __ldg(tp + t + u);
                                const uint32_t idx = (wz >> (tr.y & 0xffu)) & ((tr.y >> 8) & 0xffu);
                                gjv[u] = (wz >> ((tr.y >> 16) & 0xffu)) & 3u;
                                if (gjv[u] != 0u && idx != 0u) xv[u] = __ldg(P.gtab + tr.x + idx);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 4; u++) extra += (double)gjv[u] * xv[u];
                    }
                }
                return extra;
            };
            {
                // every lane reserves room for its flagged cells; entry = (row << 4 | word, post-flip word)
                const int ne = __popc(ev);
                int pos = ne ? atomicAdd(&s_ccount[par3], ne) : 0;
                while (ev) {
                    const int s = __ffs(ev) - 1;
                    ev &= ev - 1u;
                    const uint32_t w = word_at(s);
                    const uint32_t gm = ((gmask >> s) & 1u) ? 0xffffffffu : 0u;
                    const uint32_t wf = w ^ (((~w & 0x55555555u) << 1) & gm);
                    const uint32_t cell = rbase + ((uint32_t)s << 5);
                    if (pos < kListCap) {
                        clist[pos] = make_uint2(cell, wf);
                    } else {                                                 // list full (most cells of the tile flagged): serve it here
                        const double extra = serve(cell, wf, km, vmask, cmask, wi);
                        asm volatile("cp.async.wait_all;" ::: "memory");
                        s_acc[cell] += ((stage[0] + stage[32]) + (stage[64] + stage[96])) + extra;
                    }
                    pos++;
                }
            }
            __syncwarp();
            bar_arrive(&g_empty[ring_slot]);                                  // the tile is not needed any more
            if (++ring_slot == kFRing) { ring_slot = 0; g_par ^= 1u; }
            TM_MARK(3);
            asm volatile("bar.sync 1, %0;" ::"n"(kScanWarps * 32) : "memory");   // list complete, pending look-ups consumed
            TM_MARK(4);
            const int total = min(*reinterpret_cast<volatile int *>(&s_ccount[par3]), kListCap);
            if (sw == 0 && lane == 0) s_ccount[par3 == 0 ? 2 : par3 - 1] = 0;     // (kb + 2) % 3: last read before this barrier, next filled after the next one
            if (P.tim && sw == 0 && lane == 0) {                              // dbg & 128: cells listed / served in place / batches
                const int cnt_all = *reinterpret_cast<volatile int *>(&s_ccount[par3]);
                atomicAdd(&P.tim[0], (unsigned)total); atomicAdd(&P.tim[1], (unsigned)(cnt_all - total));
                atomicAdd(&P.tim[2], (unsigned)((total + 31) / 32)); atomicAdd(&P.tim[3], total > 0 ? 1u : 0u);
            }
            if (++par3 == 3) par3 = 0;
            for (int base = 32 * sw; base < total; base += 32 * kScanWarps) {
                const bool act = base + lane < total, keep = base + 32 * kScanWarps >= total || (P.dbg & 64);
                const uint2 en = act ? clist[base + lane] : make_uint2(0u, 0u);
                const int w16 = (int)(en.x & 15u);                            // word in the row; en.x >> 4 = row in the CTA
                // per-word constants live in the lanes that own the word
                const uint32_t km_c = __shfl_sync(0xffffffffu, km, w16), vm_c = __shfl_sync(0xffffffffu, vmask, w16);
                const uint32_t cm_c = __shfl_sync(0xffffffffu, cmask, w16);
                TM_MARK(2);
                if (act && !(P.dbg & 512)) {
                    const double extra = serve(en.x, en.y, km_c, vm_c, cm_c, kb * 16 + w16);
                    pslot = (int)en.x;                                        // s_acc[row][word]
                    if (extra != 0.0) s_acc[pslot] += extra;
                    if (!keep) consume();
                }
                __syncwarp();
            }
            TM_MARK(5);
        }
        __syncwarp();
        consume();
        // ---- per-row totals: the 16 word lanes in a fixed tree (other warps may have served this warp's cells)
        asm volatile("bar.sync 1, %0;" ::"n"(kScanWarps * 32) : "memory");
#pragma unroll
        for (int s = 0; s < kSteps; s++) {
            const int row = wrow0 + 2 * s + half;
            double t = s_acc[row * 16 + l16];
            t += __shfl_xor_sync(0xffffffffu, t, 8); t += __shfl_xor_sync(0xffffffffu, t, 4);
            t += __shfl_xor_sync(0xffffffffu, t, 2); t += __shfl_xor_sync(0xffffffffu, t, 1);
            int c = cnt_u[s];
            c += __shfl_xor_sync(0xffffffffu, c, 8); c += __shfl_xor_sync(0xffffffffu, c, 4);
            c += __shfl_xor_sync(0xffffffffu, c, 2); c += __shfl_xor_sync(0xffffffffu, c, 1);
            const int mc = s_mc[row * 16 + l16];
            int cm = mc;
            cm += __shfl_xor_sync(0xffffffffu, cm, 8); cm += __shfl_xor_sync(0xffffffffu, cm, 4);
            cm += __shfl_xor_sync(0xffffffffu, cm, 2); cm += __shfl_xor_sync(0xffffffffu, cm, 1);
            if (v0 + row < P.p) {
                P.msub_cnt[(v0 + row) * 16 + l16] = (uint16_t)mc;
                if (l16 == 0) {
                    FusedScanOut o;
                    o.acc = t; o.cnt = c; o.cm = cm; o.guess = (int)((gmask >> s) & 1u); o.pad = 0;
                    P.scan[v0 + row] = o;
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ===================================================================== MMA issuer (one thread)
        if (lane == 0) {
            constexpr uint32_t kIdesc = idesc_i8(128, 112);
            int sa = 0, sb = 0;
            uint32_t a_par = 0u, b_par = 0u;
            for (int64_t kb = 0; kb < n_kblocks; kb++) {
                TM_MARK(2);
                bar_wait(&b_full[sb], b_par);
                TM_MARK(0);
                bar_wait(&a_full[sa], a_par);
                TM_MARK(1);
                tc_fence_after();
                const uint64_t bdesc = smem_desc(s_addr(sB + (size_t)sb * kFBlockBytesB), 14 * 128, 128);
                if (!(P.dbg & 16))
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
                TM_MARK(1);
                bar_wait(&g_empty[sg], g_par);
                TM_MARK(0);
                bar_expect_tx(&g_full[sg], kTileBytes);
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(s_addr(sG + (size_t)sg * kTileBytes)), "l"(&gmap), "r"((int32_t)(kb * kFRowBytes)), "r"(row0),
                               "r"(s_addr(&g_full[sg])) : "memory");
                if (++sg == kFRing) { sg = 0; g_par ^= 1u; }
            }
        }
        __syncwarp();
    }
    TM_MARK(7);
    __syncthreads();
#ifdef FUSED_TIMERS
    if (P.tim && threadIdx.x < 32 * 8) P.tim[(size_t)blockIdx.x * 256 + threadIdx.x] = s_tim[threadIdx.x];
#endif
    if (warp == kMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// --------------------------------------------------------------------------------------------------------------
// Finaliser of the fused path: one warp per variant.  Exact flip / AF / MAF from integer counts (sum of codes from
// the contraction's constant-1 column, missing count from the scanners).  Missing samples: lane (sub-list l & 15,
// column half l >> 4) walks its sub-list in order and sums eight columns of the listed rows of Y; the 16 sub-lists are
// then combined in a fixed tree (so the result does not depend on which scanner lane found a sample first), and the
// kinship correction for missing related samples comes from the listed words.  Then the statistics of
// src/Individual_Score_Test.cpp:45-64.  A variant whose orientation guess was wrong or with an overflowed sub-list
// goes to the redo list instead.
struct FusedFinParams {
    int64_t p, n; int q, impute;
    const long long *hi, *lo, *tot;
    const double *scale, *cov;
    const FusedScanOut *scan;
    const uint2 *mlist; int msub;
    const uint16_t *msub_cnt;
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
    const int sub = lane & 15, ch = lane >> 4;
    const int q = P.q;
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
        const int mycnt = cm > 0 ? (int)P.msub_cnt[i * 16 + sub] : 0;
        const bool over = __any_sync(0xffffffffu, mycnt > P.msub);
        if (!mono && (over || (int)flip != so.guess)) {
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
        // ---- missing samples of this lane's sub-list: eight columns of Y per listed row, kinship correction (ch == 0)
        double sm[8];
#pragma unroll
        for (int c = 0; c < 8; c++) sm[c] = 0.0;
        double corr = 0.0;
        if (mycnt > 0) {
            const uint2 *ml = P.mlist + (size_t)i * 16 * P.msub + sub;     // entry k at ml[16 k]
            const bool cols = 8 * ch < P.ldY;                              // ldY = 8 or 16
            for (int k0 = 0; k0 < mycnt; k0 += 2) {
                const bool two = k0 + 1 < mycnt;
                const uint2 e0 = ml[16 * k0], e1 = two ? ml[16 * (k0 + 1)] : e0;
                if (cols) {
                    const double2 *r0 = reinterpret_cast<const double2 *>(P.Y + (int64_t)e0.x * P.ldY + 8 * ch);
                    const double2 *r1 = reinterpret_cast<const double2 *>(P.Y + (int64_t)e1.x * P.ldY + 8 * ch);
                    double2 a[4], b[4];
#pragma unroll
                    for (int c = 0; c < 4; c++) { a[c] = __ldg(r0 + c); b[c] = __ldg(r1 + c); }
#pragma unroll
                    for (int c = 0; c < 4; c++) { sm[2 * c] += a[c].x; sm[2 * c + 1] += a[c].y; }
                    if (two) {
#pragma unroll
                        for (int c = 0; c < 4; c++) { sm[2 * c] += b[c].x; sm[2 * c + 1] += b[c].y; }
                    }
                }
                if (ch == 0 && !mu_zero) {
#pragma unroll 1
                    for (int t2 = 0; t2 < (two ? 2 : 1); t2++) {
                        const uint2 en = t2 ? e1 : e0;
                        const int64_t j = en.x;
                        if ((int64_t)(j >> 4) >= P.tab_words) continue;
                        if (((__ldg(P.kmask + (j >> 4)) >> (2 * (j & 15))) & 3u) == 0u) continue;
                        const NullSparse::SampleInfo si = P.sinfo[j];
                        double sacc = 0.0;
                        for (uint32_t t = 0; t < si.off_count; t++) {
                            const NullSparse::OffEntry oe = P.off_ent[si.off_start + t];
                            // relatives of a table family share the word; it was listed post-flip (guess == flip here)
                            const uint32_t code = (en.y >> (2 * (oe.col & 15))) & 3u;
                            const double gv = code == 3u ? ((int64_t)oe.col > j ? mu : 0.0) : (double)code;
                            sacc += oe.val * gv;
                        }
                        corr += 2.0 * mu * sacc;
                    }
                }
            }
        }
        double smc = 0.0;                       // lane l: column 8 (l >> 4) + (l & 7) summed over all sub-lists
        if (cm > 0) {                           // warp-uniform
#pragma unroll
            for (int c = 0; c < 8; c++) {
                double t = sm[c];
                t += __shfl_xor_sync(0xffffffffu, t, 8); t += __shfl_xor_sync(0xffffffffu, t, 4);
                t += __shfl_xor_sync(0xffffffffu, t, 2); t += __shfl_xor_sync(0xffffffffu, t, 1);
                if ((lane & 7) == c) smc = t;
            }
            corr = warp_sum(corr);
        }
        // lane c < 16 takes column c: held by lane 16 (c >> 3) + (c & 7)
        const double smcol = __shfl_sync(0xffffffffu, smc, 16 * ((lane & 15) >> 3) + (lane & 7));
        // ---- G'[r | Sigma_iX | diag]: lane c < q + 2
        const int c16 = lane & 15;
        double L = 0.0;
        if (c16 <= q + 1) {
            long long h = P.hi[i * 16 + c16], l = P.lo[i * 16 + c16];
            if (flip) { h = 2 * P.tot[c16] - h; l = 2 * P.tot[kColsPerSlice + c16] - l; }
            L = ((double)h * 16777216.0 + (double)l) * P.scale[c16];
        }
        const double wm = flip ? mu + 1.0 : mu - 3.0;
        const double Lc = L + wm * smcol;                                  // columns 0..q
        double xb = 0.0;                                                  // lane b < q: sum_a L[1 + a] cov[a, b]
        for (int a = 0; a < q; a++) {
            const double La = __shfl_sync(0xffffffffu, Lc, 1 + a);
            if (lane < q) xb += La * P.cov[a + lane * q];
        }
        const double Lb = __shfl_sync(0xffffffffu, Lc, (lane + 1) & 31);
        double tct = lane < q ? xb * Lb : 0.0;
        tct = warp_sum(tct);
        const double U = __shfl_sync(0xffffffffu, Lc, 0);
        const double smd = __shfl_sync(0xffffffffu, smcol, q + 1);
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

// entries per missing-sample sub-list (16 sub-lists per variant): ~4 x the expectation at 0.1 % missing samples
int fused_missing_sub(int64_t n)
{
    int cap = 8;
    while ((int64_t)cap * 4000 < n) cap *= 2;
    return cap;
}

bool fused_path_available(const staar_ctx *ctx)
{
    const NullSparse &ns = ctx->ns;
    // families of more than 16 members need relatives' genotypes from anywhere in the column (scan_packed.cu keeps the
    // whole column in shared memory); every other null model streams
    return ns.set && ns.sliced && ns.Ytc && ns.lutf && ns.diag2_off > 0 && ns.diag2_off + ns.n + 64 < (1ll << 27) - 1 &&
           ns.n < (1 << 20) && ns.n_generic == 0 && ns.kin_words == ns.tab_words;
}

int launch_fused_tc(staar_ctx *ctx, const uint8_t *d_packed, size_t stride, int64_t n, int64_t p, long long *d_hi,
                    long long *d_lo, FusedScanOut *d_scan, uint2 *d_mlist, int msub, uint16_t *d_msub_cnt)
{
    if (p == 0) return STAAR_OK;
    const NullSparse &ns = ctx->ns;
    FusedParams P;
    P.p = p; P.n = n; P.n_kblocks = (n + kFSamples - 1) / kFSamples;
    P.Ytc = ns.Ytc; P.out_hi = d_hi; P.out_lo = d_lo;
    P.tab_words = ns.tab_words; P.x_words = ns.x_words;
    P.uvec = ns.diag_uniform_vec >= 0 ? (int)ns.diag_uniform_vec : INT_MAX;
    P.kmask = ns.kmask; P.wdesc = ns.wdesc; P.xterms = ns.xterms; P.gtab = ns.lutf; P.diag2_off = (uint32_t)ns.diag2_off;
    P.scan = d_scan; P.mlist = d_mlist; P.msub = msub; P.msub_cnt = d_msub_cnt;
    { const char *dbg = getenv("STAAR_B200_FUSED_DBG"); P.dbg = dbg ? atoi(dbg) : 0; }
    P.tim = nullptr;
    if (P.dbg & 128) {
        const size_t ctas = (size_t)((p + 127) / 128);
        STAAR_TRY(ctx->tmp2.reserve(sizeof(unsigned int) * 256 * ctas));
        STAAR_CUDA(cudaMemsetAsync(ctx->tmp2.p, 0, sizeof(unsigned int) * 256 * ctas, ctx->stream));
        P.tim = ctx->tmp2.as<unsigned int>();
    }
    { const char *t = getenv("STAAR_B200_FUSED_TILES"); if (t) ctx->fused_tiles = atoi(t) == 1 ? 1 : 2; }
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
                          const FusedScanOut *d_scan, const uint2 *d_mlist, int msub, const uint16_t *d_msub_cnt,
                          double *dAF, double *dMAF,
                          double *dScore, double *dScore_se, double *dpl, double *dEst, double *dEst_se, long long *d_redo,
                          unsigned long long *d_redo_count)
{
    if (p == 0) return STAAR_OK;
    const NullSparse &ns = ctx->ns;
    FusedFinParams P;
    P.p = p; P.n = n; P.q = (int)ns.q; P.impute = impute; P.hi = d_hi; P.lo = d_lo; P.tot = ns.col_total;
    P.scale = ns.col_scale; P.cov = ns.cov; P.scan = d_scan; P.mlist = d_mlist; P.msub = msub; P.msub_cnt = d_msub_cnt; P.Y = ns.Yp; P.ldY = ns.ldY;
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
