// contract_tc.cu — K2 for packed dosages on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a.
//
// Same contract as contract_packed.cu (the mma.sync version): for every variant the EXACT integers
//     R_c = sum_j code_j * X_jc   (c = residuals | Sigma_iX | diag(Sigma_i); raw 2-bit codes, missing = 3)
// with X the 54-bit fixed-point image of Y split into 7 balanced base-256 digits (see contract_packed.cu), i.e. the
// first factor of every Cov in the reference (trans(residuals)*G, trans(Sigma_iX)*G; src/Individual_Score_Test.cpp:17,42).
//
// Mapping onto tcgen05.mma.cta_group::1.kind::i8 (int8 x int8 -> int32 accumulators in TMEM):
//   M = 128 variants per CTA (TMEM lane = variant), K = samples (32 per instruction), N = 112 = 7 digit slices x 16
//   columns.  (Measured: one K = 32 instruction costs ~52 clk whatever N <= 112 is — the A tile is read from TMEM —
//   so a separate narrow N = 8 tile for the hom-minor indicator doubled the tensor time; that term now comes from
//   the scan kernel's sparse sweep instead.)
//   * B (digits) is pre-arranged at null-model set-up in the canonical K-major, non-swizzled core-matrix layout, so a
//     512-sample k-block (56 KiB) is ONE cp.async.bulk from L2 into a 2-stage shared-memory ring (mbarrier tx-count).
//   * A (genotypes) never exists as int8 in memory.  The raw 2-bit tile of a k-block (128 variant rows x 128 bytes) is
//     one TMA tensor load (cp.async.bulk.tensor.2d, 128-byte swizzle, zero fill outside the matrix) into a 4-stage
//     ring — a per-thread copy of row-strided 16-byte pieces costs one L1 wavefront per lane and was the measured
//     bottleneck.  Two "expander" threads own each variant row (half a k-block each): they read their row from the
//     swizzled tile (conflict-free 128-bit LDS), expand 2-bit -> int8 with shift/mask in registers (no flip, no
//     missing fix-up: both are applied analytically by the finaliser) and write the int8 tile straight into TMEM with
//     tcgen05.st; the MMA reads A from TMEM (".ts" form), so shared-memory bandwidth is spent on B only.
//   * One elected thread issues the MMAs; tcgen05.commit releases the A (TMEM) and B (smem) stages to the expanders and
//     the bulk-copy producer through mbarriers; the expander warps read the accumulators back with tcgen05.ld and
//     recombine the 7 digit sums per (variant, column) in 64-bit integers.
// Roles: warps 0-7 expanders + epilogue (TMEM lane quarter = warp id % 4, k-block half = warp id / 4), warp 8 MMA issuer
// + TMEM allocator, warp 9 digit producer (bulk copies), warp 10 genotype producer (tensor loads).  Every genotype byte is read from HBM exactly once per launch.
#include "common.cuh"
#include "packed.cuh"
#include "tc_common.cuh"
#include <cuda.h>            // CUtensorMap and the cuTensorMapEncodeTiled prototype (resolved at run time, no -lcuda)

namespace staar {

namespace {

constexpr int kTcRows = 128;                 // variants per CTA
constexpr int kTcThreads = 352;              // 8 expander warps + MMA issuer + digit producer + genotype producer
constexpr int kPairThreads = 384;            // CTA-pair kernel: one more warp (second MMA issuer of the leader)
constexpr int kExpThreads = 256;             // 8 expander warps
// A k-block is the unit of every hand-off (expanders -> MMA issuer -> expanders, producer -> MMA issuer): its fixed
// cost (two mbarrier waits of ~90 clk each in the single issuing thread, fences, commits) is amortised over kKS
// instructions, so k-blocks are 16 k-steps = 512 samples.
constexpr int kKS = 16;                      // k-steps (K = 32 samples each) per k-block
constexpr int kTcSamples = 32 * kKS;         // samples per k-block
constexpr int kHalfCols = 4 * kKS;           // TMEM columns (= registers) one expander thread writes per k-block
constexpr int kHalfBytes = kTcSamples / 8;   // raw 2-bit bytes one expander thread consumes per k-block
constexpr int kStagesA = 3;                  // TMEM stages of the expanded A tiles (8 kKS columns each)
constexpr int kStagesB = 2;                  // shared-memory stages of the digit operand
constexpr int kStepBytes = 2 * 14 * 128;     // one k-step: 2 k-chunks x 14 row groups x (8 rows x 16 B)
constexpr int kTcBlockBytes = kKS * kStepBytes;
constexpr int kRing = 4;                     // stages of raw genotype tiles (TMA)
constexpr int kRowBytes = kTcSamples / 4;    // raw bytes of one variant row per k-block: the TMA box is kRowBytes x 128 rows
constexpr int kTileBytes = kRowBytes * 128;
static_assert(kRowBytes == 128, "the 128-byte swizzle of the tensor load assumes 128-byte rows");
static_assert(kColsPerSlice * kSlices == 112 && 112 + kStagesA * 8 * kKS <= 512, "TMEM budget");
constexpr int kColD = 0, kColA = 128;        // TMEM columns: accumulators [0,112), A stages from 128

using namespace tc;

// switch-off experiments (compile with -DTC_DBG, select with STAAR_B200_TC_DBG): bit 0 expanders skip LDS + expansion,
// 1 skip tcgen05.st, 2 no digit copies, 3 no genotype tensor loads, 4 every MMA reads the same digit k-step, 5 no MMAs
#ifdef TC_WAIT_HINT
#define TC_WAIT_EXP bar_wait_hint
#else
#define TC_WAIT_EXP bar_wait
#endif
#ifdef TC_DBG
__device__ int g_tc_dbg;
#define TC_SW(bit) ((dbg >> (bit)) & 1)
#else
#define TC_SW(bit) 0
#endif

// grid = ceil(p / 128); out_hi / out_lo: p x 16 int64, X = hi * 2^24 + lo
__global__ void __launch_bounds__(kTcThreads, 1)
contract_tc_kernel(const __grid_constant__ CUtensorMap gmap, int64_t p, const int8_t *__restrict__ Ytc,
                   int64_t n_kblocks, long long *__restrict__ out_hi, long long *__restrict__ out_lo,
                   uint2 *__restrict__ mlist, uint32_t *__restrict__ mcnt, int mcap)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sG = smem;                                                       // kRing x kTileBytes (1024-byte aligned: swizzle)
    uint8_t *sB = smem + (size_t)kRing * kTileBytes;                           // kStagesB x kTcBlockBytes
    uint64_t *bars = reinterpret_cast<uint64_t *>(sB + (size_t)kStagesB * kTcBlockBytes);
    uint64_t *a_full = bars, *a_empty = bars + kStagesA;
    uint64_t *b_full = bars + 2 * kStagesA, *b_empty = b_full + kStagesB;
    uint64_t *g_full = b_empty + kStagesB, *g_empty = g_full + kRing;
    uint64_t *d_full = g_empty + kRing;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(d_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kMmaWarp = kExpThreads / 32, kProdWarp = kMmaWarp + 1;
#ifdef TC_DBG
    const int dbg = g_tc_dbg;
#endif

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStagesA; s++) { bar_init(&a_full[s], kExpThreads); bar_init(&a_empty[s], 1); }
        for (int s = 0; s < kStagesB; s++) { bar_init(&b_full[s], 1); bar_init(&b_empty[s], 1); }
        for (int s = 0; s < kRing; s++) { bar_init(&g_full[s], 1); bar_init(&g_empty[s], kExpThreads); }
        bar_init(d_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {                                                   // whole warp: allocate all 512 TMEM columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_addr(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < kMmaWarp) {
        // ===================================================================== expanders: (variant row, k-block half)
        const int row = threadIdx.x & (kTcRows - 1), half = threadIdx.x >> 7;   // TMEM lane quarter = warp % 4
        const int64_t v = (int64_t)blockIdx.x * kTcRows + row;
        const bool rok = v < p;
        const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        // Software pipeline: the int8 tile of k-block kb+1 is expanded into registers while the TMEM stores of k-block
        // kb are still in flight; only then does the thread wait for them and publish stage kb to the MMA issuer.
        // K index 4c + y of k-step ks <-> sample 32 ks + 16 (c >> 2) + 4 y + (c & 3): column c of a k-step is field
        // (c & 3) of word 2 ks + (c >> 2); the digit operand is laid out with the same map (build_tc_operand).
        uint32_t x[kHalfCols];
        int ring_slot = 0;                                                   // kb % kRing, kept incrementally (no 64-bit %)
        uint32_t g_par = 0u;                                                 // parity to wait for on g_full[ring_slot]
        // this thread's row of the swizzled tile: logical 16-byte chunk c sits at physical chunk c ^ (row & 7)
        const uint8_t *my_row = sG + (size_t)row * kRowBytes;
        // Missing samples (code 3) are listed here, where every word of the column passes through registers anyway and the
        // expander threads have slack (they wait ~40 % of a k-block for the tensor pipe): thread (row, half) appends one
        // record {32-sample segment, 32-bit mask} per segment that holds missing samples, in increasing order, to its own
        // sub-list mlist[(2 v + half) * mcap ..] (mask bit 2 f + w = sample 16 w + f of the segment, the scan's own
        // encoding); the scan kernel then needs no detection pass and knows the missing count up front.  A record count
        // above mcap marks the variant for the scan's own detection path.
        uint2 *my_list = mlist ? mlist + ((size_t)v * 2 + half) * (size_t)mcap : nullptr;
        uint32_t mc = 0;                                                     // records written (or counted, beyond mcap)
#ifdef TC_TIMERS
        long long tm_ae = 0, tm_st = 0, tm_gf = 0, tm_ex = 0, tm_ws = 0, tm_t0;
        const bool tm_on = threadIdx.x == 0 && blockIdx.x == 500;
#define TCT(acc) do { if (tm_on) { const long long t_ = clock64(); acc += t_ - tm_t0; tm_t0 = t_; } } while (0)
#else
#define TCT(acc) do { } while (0)
#endif
        auto expand = [&](int64_t kbx) {
            TC_WAIT_EXP(&g_full[ring_slot], g_par);                          // the tensor load of this k-block has landed
            TCT(tm_gf);
            const uint8_t *rowp = my_row + (size_t)ring_slot * kTileBytes;
            uint32_t w[kHalfBytes / 4];
            if (TC_SW(0)) {
                bar_arrive(&g_empty[ring_slot]);
                if (++ring_slot == kRing) { ring_slot = 0; g_par ^= 1u; }
                return;
            }
#pragma unroll
            for (int c = 0; c < kHalfBytes / 16; c++) {
                const int chunk = (half * (kHalfBytes / 16) + c) ^ (row & 7);
                const uint4 q = *reinterpret_cast<const uint4 *>(rowp + 16 * chunk);
                w[4 * c] = q.x; w[4 * c + 1] = q.y; w[4 * c + 2] = q.z; w[4 * c + 3] = q.w;
            }
            bar_arrive(&g_empty[ring_slot]);                                 // words are in registers: the stage may be refilled
            if (++ring_slot == kRing) { ring_slot = 0; g_par ^= 1u; }
            if (my_list) {
                uint32_t anyc[kHalfBytes / 16];
#pragma unroll
                for (int c = 0; c < kHalfBytes / 16; c++)
                    anyc[c] = ((w[4 * c] & (w[4 * c] >> 1)) | (w[4 * c + 1] & (w[4 * c + 1] >> 1)) |
                               (w[4 * c + 2] & (w[4 * c + 2] >> 1)) | (w[4 * c + 3] & (w[4 * c + 3] >> 1))) & 0x55555555u;
                uint32_t any = 0u;
#pragma unroll
                for (int c = 0; c < kHalfBytes / 16; c++) any |= anyc[c];
                if (any) {                                                   // rare per thread: ~0.1 % of the samples
                    const uint32_t seg0 = ((uint32_t)kbx * kTcSamples + (uint32_t)half * (kTcSamples / 2)) >> 5;
#pragma unroll
                    for (int c = 0; c < kHalfBytes / 16; c++)
                        if (anyc[c]) {
#pragma unroll
                            for (int hh = 0; hh < 2; hh++) {
                                const uint32_t wa = w[4 * c + 2 * hh], wb = w[4 * c + 2 * hh + 1];
                                const uint32_t mask = (wa & (wa >> 1) & 0x55555555u) | ((wb & (wb >> 1) & 0x55555555u) << 1);
                                if (mask) {
                                    if (mc < (uint32_t)mcap) my_list[mc] = make_uint2(seg0 + 2u * c + hh, mask);
                                    mc++;
                                }
                            }
                        }
                }
            }
#pragma unroll
            // right shifts as IMAD.HI (w * 2^(32-s) >> 32) run on the FMA pipe, leaving the ALU pipe the masks only
            for (int c = 0; c < kHalfCols; c++)
                x[c] = ((c & 3) == 0 ? w[c >> 2] : __umulhi(w[c >> 2], 1u << (32 - 2 * (c & 3)))) & 0x03030303u;
        };
#ifdef TC_TIMERS
        tm_t0 = clock64();
#endif
        if (n_kblocks > 0) expand(0);
        int sa = 0;
        uint32_t a_par = 1u;                                                 // parity to wait for on a_empty[sa]
        for (int64_t kb = 0; kb < n_kblocks; kb++) {
            TCT(tm_ex);
            TC_WAIT_EXP(&a_empty[sa], a_par);                                // MMAs that read this TMEM stage are done
            TCT(tm_ae);
            tc_fence_after();
            if (!TC_SW(1)) {
#pragma unroll
            for (int b = 0; b < kHalfCols / 32; b++)
                tmem_st32(lane_base + kColA + sa * (8 * kKS) + kHalfCols * half + 32 * b, *reinterpret_cast<const uint32_t (*)[32]>(&x[32 * b]));
            }
            TCT(tm_st);
            if (kb + 1 < n_kblocks) expand(kb + 1);                          // overlaps the store latency
            TCT(tm_ex);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            TCT(tm_ws);
            tc_fence_before();
            bar_arrive(&a_full[sa]);
            if (++sa == kStagesA) { sa = 0; a_par ^= 1u; }
        }
#ifdef TC_TIMERS
        if (tm_on) printf("expander: per k-block clk: wait a_empty %lld, st issue %lld, wait g_full %lld, lds+alu %lld, wait::st %lld\n",
                          tm_ae / n_kblocks, tm_st / n_kblocks, tm_gf / n_kblocks, tm_ex / n_kblocks, tm_ws / n_kblocks);
#endif
        if (my_list && rok) mcnt[v * 2 + half] = mc;
        // ===================================================================== epilogue: TMEM -> 64-bit recombination
        // half 0 reads the accumulators of digit slices 0-3 (the "hi" part), half 1 those of slices 4-6
        bar_wait(d_full, 0u);
        tc_fence_after();
        int acc[64];
        {
            int t32[32];
#pragma unroll
            for (int b = 0; b < 2; b++) {
                if (half == 0 || b == 0) {
                    tmem_ld32(lane_base + kColD + 64 * half + 32 * b, t32);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int e = 0; e < 32; e++) acc[32 * b + e] = t32[e];
                } else {
                    int t16[16];
                    tmem_ld16(lane_base + kColD + 96, t16);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int e = 0; e < 16; e++) acc[32 + e] = t16[e];
                }
            }
        }
        if (rok) {
            if (half == 0) {
#pragma unroll
                for (int c = 0; c < 16; c++) {                               // accumulator column = 16 * slice + c
                    long long hi = 0;
#pragma unroll
                    for (int sl = 0; sl < 4; sl++) hi = hi * 256 + (long long)acc[16 * sl + c];
                    out_hi[v * 16 + c] = hi;
                }
            } else {
#pragma unroll
                for (int c = 0; c < 16; c++) {
                    long long lo = 0;
#pragma unroll
                    for (int sl = 4; sl < kSlices; sl++) lo = lo * 256 + (long long)acc[16 * (sl - 4) + c];
                    out_lo[v * 16 + c] = lo;
                }
            }
        }
        tc_fence_before();
    } else if (warp == kMmaWarp) {
        // ===================================================================== MMA issuer (one thread)
        // (Tried: two threads alternating k-blocks, accumulators zeroed by tcgen05.st so that no order is needed between
        // them.  It removes the ~250 clk the issuer spends in mbarrier waits per k-block — the tensor pipe queues hardly
        // more than one instruction, tools/pipe_peaks "delay" rows — and reaches the 56 clk/MMA peak when the operand
        // streams are switched off, but with them on the SM's ingress (4.5 KB per K = 32 instruction) is the bound and
        // the whole kernel was not faster: 2.47 vs 2.38 ms per 250 k variants.  DESIGN.md section 4.)
        if (lane == 0) {
            constexpr uint32_t kIdescMain = idesc_i8(kTcRows, 112);
            int sa = 0, sb = 0;
            uint32_t a_par = 0u, b_par = 0u;
#ifdef TC_TIMERS
            long long ti_b = 0, ti_a = 0, ti_i = 0, ti_t0 = clock64();
            const bool ti_on = blockIdx.x == 500;
#define TCI(acc) do { if (ti_on) { const long long t_ = clock64(); acc += t_ - ti_t0; ti_t0 = t_; } } while (0)
#else
#define TCI(acc) do { } while (0)
#endif
            for (int64_t kb = 0; kb < n_kblocks; kb++) {
                TCI(ti_i);
                bar_wait(&b_full[sb], b_par);
                TCI(ti_b);
                bar_wait(&a_full[sa], a_par);
                TCI(ti_a);
                tc_fence_after();
                const uint64_t bdesc = smem_desc(s_addr(sB + (size_t)sb * kTcBlockBytes), 14 * 128, 128);
                const uint32_t a1 = tmem + kColA + sa * (8 * kKS);
#pragma unroll
                for (int ks = 0; ks < kKS; ks++)     // descriptor start address advances by one k-step (3584 B >> 4)
                    if (!TC_SW(5))
                    umma_i8_ts(tmem + kColD, a1 + 8 * ks, bdesc + (uint64_t)(TC_SW(4) ? 0 : ks * (kStepBytes >> 4)), kIdescMain,
                               (kb > 0 || ks > 0) ? 1u : 0u);
                tc_commit(&a_empty[sa]);                                    // frees the TMEM A stage when these MMAs retire
                tc_commit(&b_empty[sb]);                                    // frees the shared-memory digit stage
                if (++sa == kStagesA) { sa = 0; a_par ^= 1u; }
                if (++sb == kStagesB) { sb = 0; b_par ^= 1u; }
            }
            tc_commit(d_full);
#ifdef TC_TIMERS
            if (ti_on) printf("issuer: per k-block clk: wait b_full %lld, wait a_full %lld, issue %lld\n", ti_b / n_kblocks, ti_a / n_kblocks, ti_i / n_kblocks);
#endif
        }
        __syncwarp();
    } else if (warp == kProdWarp) {
        // ===================================================================== digit producer (one thread)
        if (lane == 0) {
            int sb = 0;
            uint32_t b_par = 1u;
            for (int64_t kb = 0; kb < n_kblocks; kb++) {
                bar_wait_relaxed(&b_empty[sb], b_par);
                if (TC_SW(2)) bar_arrive(&b_full[sb]);
                else {
                bar_expect_tx(&b_full[sb], kTcBlockBytes);
                bulk_load(sB + (size_t)sb * kTcBlockBytes, Ytc + (size_t)kb * kTcBlockBytes, kTcBlockBytes, &b_full[sb]);
                }
                if (++sb == kStagesB) { sb = 0; b_par ^= 1u; }
            }
        }
        __syncwarp();
    } else {
        // ===================================================================== genotype producer (one thread)
        if (lane == 0) {
            int sg = 0;
            uint32_t g_par = 1u;
            const int32_t row0 = (int32_t)(blockIdx.x * kTcRows);
            for (int64_t kb = 0; kb < n_kblocks; kb++) {
                bar_wait(&g_empty[sg], g_par);                               // raw genotype tile: rows row0.., bytes kb * 128..
                if (TC_SW(3)) bar_arrive(&g_full[sg]);
                else {
                bar_expect_tx(&g_full[sg], kTileBytes);
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(s_addr(sG + (size_t)sg * kTileBytes)), "l"(&gmap), "r"((int32_t)(kb * kRowBytes)), "r"(row0),
                               "r"(s_addr(&g_full[sg])) : "memory");
                }
                if (++sg == kRing) { sg = 0; g_par ^= 1u; }
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


// ---------------------------------------------------------------------------------------------------------------------
// CTA-pair version (tcgen05 cta_group::2): two CTAs of a cluster own 256 consecutive variants (128 TMEM lanes each) and
// ONE digit stream: each CTA stages half of the N = 112 operand rows (56 rows, 28 KiB per k-block) and the leader's
// M = 256 instruction reads both halves, so per variant half as many digit bytes leave L2 and enter shared memory.
// Hand-offs: the peer's expanders and its bulk copies complete on the peer's own barriers; one relay thread there forwards
// "k-block kb is staged" to the leader (remote mbarrier arrive); tcgen05.commit multicasts every release to both CTAs.
constexpr int kStagesB2 = 4;                  // half-size digit stages
constexpr int kStepBytes2 = 2 * 7 * 128;      // one k-step of one CTA: 2 k-chunks x 7 row groups x (8 rows x 16 B)
constexpr int kPairBlockBytes = kKS * kStepBytes2;
constexpr int kRelayRing = 4;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
contract_tc_pair_kernel(const __grid_constant__ CUtensorMap gmap, int64_t p, const int8_t *__restrict__ Ypair,
                        int64_t n_kblocks, long long *__restrict__ out_hi, long long *__restrict__ out_lo)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sG = smem;                                                       // kRing x kTileBytes
    uint8_t *sB = smem + (size_t)kRing * kTileBytes;                           // kStagesB2 x kPairBlockBytes
    uint64_t *bars = reinterpret_cast<uint64_t *>(sB + (size_t)kStagesB2 * kPairBlockBytes);
    uint64_t *a_full = bars, *a_empty = bars + kStagesA;
    uint64_t *b_full = bars + 2 * kStagesA, *b_empty = b_full + kStagesB2;
    uint64_t *g_full = b_empty + kStagesB2, *g_empty = g_full + kRing;
    uint64_t *peer_ready = g_empty + kRing;                                    // leader: k-block staged in the peer CTA
    uint64_t *d_full = peer_ready + kRelayRing;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(d_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kMmaWarp = kExpThreads / 32, kProdWarp = kMmaWarp + 1, kMma2Warp = kMmaWarp + 3;
    const uint32_t rank = cluster_rank();

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStagesA; s++) { bar_init(&a_full[s], kExpThreads); bar_init(&a_empty[s], 1); }
        for (int s = 0; s < kStagesB2; s++) { bar_init(&b_full[s], 1); bar_init(&b_empty[s], 1); }
        for (int s = 0; s < kRing; s++) { bar_init(&g_full[s], 1); bar_init(&g_empty[s], kExpThreads); }
        for (int s = 0; s < kRelayRing; s++) bar_init(&peer_ready[s], 1);
        bar_init(d_full, 2);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {                                                   // one warp of each CTA
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_addr(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();                                                       // barriers of both CTAs initialised, TMEM allocated
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < kMmaWarp) {
        // ===================================================================== expanders (as in the single-CTA kernel)
        const int row = threadIdx.x & (kTcRows - 1), half = threadIdx.x >> 7;
        const int64_t v = (int64_t)blockIdx.x * kTcRows + row;
        const bool rok = v < p;
        const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t x[kHalfCols];
        int ring_slot = 0;
        uint32_t g_par = 0u;
        const uint8_t *my_row = sG + (size_t)row * kRowBytes;
        auto expand = [&](int64_t) {
            bar_wait(&g_full[ring_slot], g_par);
            const uint8_t *rowp = my_row + (size_t)ring_slot * kTileBytes;
            uint32_t w[kHalfBytes / 4];
#pragma unroll
            for (int c = 0; c < kHalfBytes / 16; c++) {
                const int chunk = (half * (kHalfBytes / 16) + c) ^ (row & 7);
                const uint4 q = *reinterpret_cast<const uint4 *>(rowp + 16 * chunk);
                w[4 * c] = q.x; w[4 * c + 1] = q.y; w[4 * c + 2] = q.z; w[4 * c + 3] = q.w;
            }
            bar_arrive(&g_empty[ring_slot]);
            if (++ring_slot == kRing) { ring_slot = 0; g_par ^= 1u; }
#pragma unroll
            for (int c = 0; c < kHalfCols; c++)
                x[c] = ((c & 3) == 0 ? w[c >> 2] : __umulhi(w[c >> 2], 1u << (32 - 2 * (c & 3)))) & 0x03030303u;
        };
        {
            uint32_t z[32];                                                  // accumulators start at zero (see the single-CTA kernel)
#pragma unroll
            for (int e = 0; e < 32; e++) z[e] = 0u;
            tmem_st32(lane_base + kColD + 64 * half, z);
            tmem_st32(lane_base + kColD + 64 * half + 32, z);
        }
        if (n_kblocks > 0) expand(0);
        int sa = 0;
        uint32_t a_par = 1u;
        for (int64_t kb = 0; kb < n_kblocks; kb++) {
            bar_wait(&a_empty[sa], a_par);                                   // the pair's MMAs that read this stage are done
            tc_fence_after();
#pragma unroll
            for (int b = 0; b < kHalfCols / 32; b++)
                tmem_st32(lane_base + kColA + sa * (8 * kKS) + kHalfCols * half + 32 * b, *reinterpret_cast<const uint32_t (*)[32]>(&x[32 * b]));
            if (kb + 1 < n_kblocks) expand(kb + 1);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            bar_arrive(&a_full[sa]);
            if (++sa == kStagesA) { sa = 0; a_par ^= 1u; }
        }
        // ===================================================================== epilogue: this CTA's 128 accumulator lanes
        bar_wait(d_full, 0u);
        tc_fence_after();
        int acc[64];
        {
            int t32[32];
#pragma unroll
            for (int b = 0; b < 2; b++) {
                if (half == 0 || b == 0) {
                    tmem_ld32(lane_base + kColD + 64 * half + 32 * b, t32);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int e = 0; e < 32; e++) acc[32 * b + e] = t32[e];
                } else {
                    int t16[16];
                    tmem_ld16(lane_base + kColD + 96, t16);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int e = 0; e < 16; e++) acc[32 + e] = t16[e];
                }
            }
        }
        if (rok) {
            if (half == 0) {
#pragma unroll
                for (int c = 0; c < 16; c++) {
                    long long hi = 0;
#pragma unroll
                    for (int sl = 0; sl < 4; sl++) hi = hi * 256 + (long long)acc[16 * sl + c];
                    out_hi[v * 16 + c] = hi;
                }
            } else {
#pragma unroll
                for (int c = 0; c < 16; c++) {
                    long long lo = 0;
#pragma unroll
                    for (int sl = 4; sl < kSlices; sl++) lo = lo * 256 + (long long)acc[16 * (sl - 4) + c];
                    out_lo[v * 16 + c] = lo;
                }
            }
        }
        tc_fence_before();
    } else if (warp == kMmaWarp || warp == kMma2Warp) {
        if (lane == 0) {
            static_assert(kStagesB2 == 4 && kRelayRing == 4 && kStagesA == 3, "stage bookkeeping of the alternating issuers");
            if (rank == 0) {
                // ================================================================= MMA issuers of the pair: two threads, alternating k-blocks
                constexpr uint32_t kIdescPair = idesc_i8(2 * kTcRows, 112);
                const int t = warp == kMmaWarp ? 0 : 1;
                int sa = t, sb = t;
                uint32_t a_par = 0u, b_par = 0u;
                for (int64_t kb = t; kb < n_kblocks; kb += 2) {
                    bar_wait(&b_full[sb], b_par);
                    bar_wait(&a_full[sa], a_par);
                    bar_wait_cluster(&peer_ready[sb], b_par);                 // the peer's A stage and digit half (same ring as b)
                    tc_fence_after();
                    const uint64_t bdesc = smem_desc(s_addr(sB + (size_t)sb * kPairBlockBytes), 7 * 128, 128);
                    const uint32_t a1 = tmem + kColA + sa * (8 * kKS);
#pragma unroll
                    for (int ks = 0; ks < kKS; ks++)
                        umma_i8_ts_pair(tmem + kColD, a1 + 8 * ks, bdesc + (uint64_t)(ks * (kStepBytes2 >> 4)), kIdescPair, 1u);
                    tc_commit_pair(&a_empty[sa]);
                    tc_commit_pair(&b_empty[sb]);
                    sa += 2;
                    if (sa >= kStagesA) { sa -= kStagesA; a_par ^= 1u; }
                    sb += 2;
                    if (sb >= kStagesB2) { sb -= kStagesB2; b_par ^= 1u; }
                }
                tc_commit_pair(d_full);
            } else if (warp == kMmaWarp) {
                // ================================================================= relay: "k-block staged here" -> leader
                int sa = 0, sb = 0;
                uint32_t a_par = 0u, b_par = 0u;
                for (int64_t kb = 0; kb < n_kblocks; kb++) {
                    bar_wait(&b_full[sb], b_par);
                    bar_wait(&a_full[sa], a_par);
                    tc_fence_after();
                    tc_fence_before();
                    bar_arrive_remote(&peer_ready[sb], 0u);
                    if (++sa == kStagesA) { sa = 0; a_par ^= 1u; }
                    if (++sb == kStagesB2) { sb = 0; b_par ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp == kProdWarp) {
        // ===================================================================== digit producer: this CTA's half of N
        if (lane == 0) {
            int sb = 0;
            uint32_t b_par = 1u;
            const int8_t *src = Ypair + (size_t)rank * kPairBlockBytes;
            for (int64_t kb = 0; kb < n_kblocks; kb++) {
                bar_wait_relaxed(&b_empty[sb], b_par);
                bar_expect_tx(&b_full[sb], kPairBlockBytes);
                bulk_load(sB + (size_t)sb * kPairBlockBytes, src + (size_t)kb * (2 * kPairBlockBytes), kPairBlockBytes, &b_full[sb]);
                if (++sb == kStagesB2) { sb = 0; b_par ^= 1u; }
            }
        }
        __syncwarp();
    } else {
        // ===================================================================== genotype producer (one thread)
        if (lane == 0) {
            int sg = 0;
            uint32_t g_par = 1u;
            const int32_t row0 = (int32_t)(blockIdx.x * kTcRows);
            for (int64_t kb = 0; kb < n_kblocks; kb++) {
                bar_wait(&g_empty[sg], g_par);
                bar_expect_tx(&g_full[sg], kTileBytes);
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(s_addr(sG + (size_t)sg * kTileBytes)), "l"(&gmap), "r"((int32_t)(kb * kRowBytes)), "r"(row0),
                               "r"(s_addr(&g_full[sg])) : "memory");
                if (++sg == kRing) { sg = 0; g_par ^= 1u; }
            }
        }
        __syncwarp();
    }
    tc_fence_before();
    cluster_sync_all();                                                       // both CTAs are done with TMEM and each other's smem
    if (warp == kMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

}  // namespace

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda)
static int encode_genotype_map(CUtensorMap *map, const uint8_t *d_packed, size_t stride, int64_t p)
{
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    STAAR_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled is not available in this driver"); return STAAR_ERR_CUDA; }
    const EncodeFn encode = reinterpret_cast<EncodeFn>(fn);
    // rank-2 byte tensor: dim 0 = bytes of a column (stride), dim 1 = variants; box = 128 bytes x 128 variants;
    // reads beyond either extent are zero-filled (missing rows of the last tile, the tail of the last k-block)
    const cuuint64_t dims[2] = {(cuuint64_t)stride, (cuuint64_t)p};
    const cuuint64_t strides[1] = {(cuuint64_t)stride};
    const cuuint32_t box[2] = {(cuuint32_t)kRowBytes, (cuuint32_t)kTcRows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t *>(d_packed), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (CUresult %d): packed base must be 16-byte aligned", (int)r); return STAAR_ERR_CUDA; }
    return STAAR_OK;
}

int launch_contract_tc(staar_ctx *ctx, const uint8_t *d_packed, size_t stride, int64_t n, int64_t p, long long *d_hi,
                       long long *d_lo, uint2 *d_mlist, uint32_t *d_mcnt, int mcap)
{
    if (p == 0) return STAAR_OK;
    const NullSparse &ns = ctx->ns;
    if (!ns.sliced || !ns.Ytc) { set_error("packed contraction: sliced operand not built (q > 13?)"); return STAAR_ERR_STATE; }
    const int64_t n_kblocks = (n + kTcSamples - 1) / kTcSamples;
    if (ctx->use_tc_pair && ns.Ypair && !d_mlist) {
        const size_t smem2 = (size_t)kRing * kTileBytes + (size_t)kStagesB2 * kPairBlockBytes + 40 * sizeof(uint64_t);
        if (!ctx->attr_contract_pair) {
            STAAR_CUDA(cudaFuncSetAttribute(contract_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            ctx->attr_contract_pair = true;
        }
        CUtensorMap gmap2;
        STAAR_TRY(encode_genotype_map(&gmap2, d_packed, stride, p));
        const int64_t pairs = (p + 2 * kTcRows - 1) / (2 * kTcRows);             // rows past p are zero-filled by the tensor load
        contract_tc_pair_kernel<<<(unsigned)(2 * pairs), kPairThreads, smem2, ctx->stream>>>(gmap2, p, ns.Ypair, n_kblocks, d_hi, d_lo);
        ctx->launches++;
        STAAR_CUDA(cudaGetLastError());
        return STAAR_OK;
    }
    const size_t smem = (size_t)kRing * kTileBytes + (size_t)kStagesB * kTcBlockBytes + 32 * sizeof(uint64_t);
#ifdef TC_DBG
    {
        const int dbg = getenv("STAAR_B200_TC_DBG") ? atoi(getenv("STAAR_B200_TC_DBG")) : 0;
        STAAR_CUDA(cudaMemcpyToSymbolAsync(g_tc_dbg, &dbg, sizeof(int), 0, cudaMemcpyHostToDevice, ctx->stream));
    }
#endif
    if (!ctx->attr_contract_tc) {               // per context: the attribute belongs to the context's device
        STAAR_CUDA(cudaFuncSetAttribute(contract_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->attr_contract_tc = true;
    }
    CUtensorMap gmap;
    STAAR_TRY(encode_genotype_map(&gmap, d_packed, stride, p));
    const int64_t tiles = (p + kTcRows - 1) / kTcRows;
    contract_tc_kernel<<<(unsigned)tiles, kTcThreads, smem, ctx->stream>>>(gmap, p, ns.Ytc, n_kblocks, d_hi, d_lo, d_mlist, d_mcnt, mcap);
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

// Digit operand in the canonical layout of the tcgen05 shared-memory descriptors used above.  One k-block:
//   [k-step kKS][k-chunk 2][row group 14][row 8][16 B]   row = 16 * slice + column   (N = 112); k-blocks are kTcBlockBytes apart
// K index kk = 16 * k-chunk + byte of a k-step <-> sample kTcSamples kb + 32 ks + 16 ((kk >> 2) >> 2) + 4 (kk & 3) + ((kk >> 2) & 3).
int build_tc_operand(staar_ctx *ctx, const int8_t *digits /* n x kColsPerSlice x kSlices, position order */, int ncol)
{
    NullSparse &ns = ctx->ns;
    const int64_t n = ns.n;
    const int64_t n_kblocks = (n + kTcSamples - 1) / kTcSamples;
    std::vector<int8_t> img((size_t)n_kblocks * kTcBlockBytes, 0);
    for (int64_t kb = 0; kb < n_kblocks; kb++)
        for (int ks = 0; ks < kKS; ks++)
            for (int kk = 0; kk < 32; kk++) {
                const int c = kk >> 2, y = kk & 3;
                const int64_t smp = kb * kTcSamples + 32 * ks + 16 * (c >> 2) + 4 * y + (c & 3);
                if (smp >= n) continue;
                const int8_t *dg = digits + (size_t)smp * kColsPerSlice * kSlices;
                int8_t *base = img.data() + (size_t)kb * kTcBlockBytes;
                const size_t koff = (size_t)(kk >> 4) * (14 * 128) + (kk & 15);
                for (int sl = 0; sl < kSlices; sl++)
                    for (int col = 0; col < ncol; col++) {
                        const int row = 16 * sl + col;
                        base[(size_t)ks * kStepBytes + koff + (size_t)(row >> 3) * 128 + (size_t)(row & 7) * 16] = dg[col * kSlices + sl];
                    }
            }
    cudaFree(ns.Ytc);
    ns.Ytc = nullptr;
    STAAR_CUDA(cudaMalloc(&ns.Ytc, img.size()));
    ns.reg(&ns.Ytc, img.size());
    STAAR_CUDA(cudaMemcpyAsync(ns.Ytc, img.data(), img.size(), cudaMemcpyHostToDevice, ctx->stream));
    STAAR_CUDA(cudaStreamSynchronize(ctx->stream));
    // the CTA-pair kernel's image: per k-block [CTA 2][k-step kKS][k-chunk 2][row group 7][row 8][16 B], CTA r holding
    // operand rows [56 r, 56 r + 56)
    std::vector<int8_t> img2(img.size(), 0);
    for (int64_t kb = 0; kb < n_kblocks; kb++)
        for (int ks = 0; ks < kKS; ks++)
            for (int kc = 0; kc < 2; kc++)
                for (int row = 0; row < 112; row++) {
                    const int8_t *from = img.data() + (size_t)kb * kTcBlockBytes + (size_t)ks * kStepBytes + (size_t)kc * (14 * 128) +
                                         (size_t)(row >> 3) * 128 + (size_t)(row & 7) * 16;
                    const int r2 = row % 56;
                    int8_t *to = img2.data() + (size_t)kb * kTcBlockBytes + (size_t)(row / 56) * kPairBlockBytes +
                                 (size_t)ks * kStepBytes2 + (size_t)kc * (7 * 128) + (size_t)(r2 >> 3) * 128 + (size_t)(r2 & 7) * 16;
                    memcpy(to, from, 16);
                }
    cudaFree(ns.Ypair);
    ns.Ypair = nullptr;
    STAAR_CUDA(cudaMalloc(&ns.Ypair, img2.size()));
    ns.reg(&ns.Ypair, img2.size());
    STAAR_CUDA(cudaMemcpyAsync(ns.Ypair, img2.data(), img2.size(), cudaMemcpyHostToDevice, ctx->stream));
    STAAR_CUDA(cudaStreamSynchronize(ctx->stream));
    return STAAR_OK;
}

}  // namespace staar
