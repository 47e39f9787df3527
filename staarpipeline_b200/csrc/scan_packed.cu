// scan_packed.cu — the sparse, per-variant terms of the packed path (K1 + K3) and its finaliser (K5).
//
// What the exact tensor-core contraction (contract_tc.cu) cannot deliver are the terms that are not linear in the raw
// 2-bit codes; scan_block_kernel computes them, one CTA per variant with the column staged in shared memory:
//   * flip / impute decisions, AF and MAF — bit-exact with matrix_flip_mean / matrix_flip_minor
//     (src/matrix_flip_mean.cpp:23-70, src/matrix_flip_minor.cpp:23-95) from exact integers: the sum of codes comes from
//     the contraction's constant-1 column, the missing count from this kernel;
//   * column sums of Y over the missing samples (imputation correction of G'[r | Sigma_iX]);
//   * g' Sigma_i g beyond its linear part (the (trans(G)*Sigma_i)*G factor of src/Individual_Score_Test.cpp:43, diagonal
//     entry only): the sum of Sigma_i[j,j] over hom-minor carriers (g^2 - g = 2 [g == 2]) and the kinship off-diagonal
//     part, which thanks to the null-model sample order (common.cuh) is a table look-up per 4-sample group.
// finalize_packed_kernel (one thread per variant): flip and imputation applied analytically to the exact integer sums,
//   Cov_ii = g' Sigma_i g - t' cov t, and the statistics loop of src/Individual_Score_Test.cpp:45-64.
//
// Integer/byte work: 128-bit shared-memory reads, popc/ffs bit tricks, warp-shuffle reductions with fixed trees and
// per-warp queues filled in a fixed order (no floating-point atomics => 1/2/4/8-GPU runs are bit-identical).
#include <climits>
#include "common.cuh"
#include "packed.cuh"

namespace staar {

namespace {

struct ScanParams {
    const uint8_t *packed; size_t stride; int64_t n, p; int impute;
    size_t sbytes;                         // bytes of a column staged in shared memory (list mode: the kinship prefix only)
    // generic kinship path (positions in null-model order)
    const uint8_t *relmask;                // 2-bit mask, 11 = sample has kinship off-diagonals with a larger position
    const NullSparse::CarrierRec *rel;     // per-sample record: first two such entries inline (32 B)
    const NullSparse::OffEntry *up_ent;    // all upper-triangle entries (rows with more than two)
    // table kinship path: families inside one 16-sample word
    int64_t tab_words, kin_words, x_words;
    const uint32_t *kmask;                 // per table word: both bits of every related slot
    const NullSparse::WordDesc *wdesc;
    const NullSparse::KinTerm *xterms;
    const double *lut;
    const NullSparse::SampleInfo *sinfo;   // full kinship rows per position (missing related samples)
    const double *diag;                    // Sigma_i[j,j] by position
    int uvec; double dval;                 // vectors >= uvec: every sample has Sigma_i[j,j] == dval (uvec = INT_MAX: none)
    const NullSparse::OffEntry *off_ent;
    const double *Y; int ldY, ncol;        // Y rows in null-model order
    const long long *codesum;              // contraction output `lo` (p x 16): slot 15 = exact sum of the codes
    // list mode (redo list of the fused path): work item k is variant vlist[k], k < *vcount (both device-resident)
    const long long *vlist; const unsigned long long *vcount;
    // missing-sample lists (kLists instantiation): the contraction kernel writes segment records, expand_lists_kernel
    // turns them into positions: variant i owns plist[i * pcap ..], pcnt[i] = number of missing samples, or kTruncated
    // when the lists did not fit — such a variant is left to the detection path (list mode of the other instantiation)
    const uint32_t *plist; const uint32_t *pcnt; int pcap;
    double *AF, *MAF; int32_t *flip; double *mu, *quad, *Sm;
};

__device__ __forceinline__ double code_value(uint32_t code, double mu)
{
    return code == 3u ? mu : (double)code;
}

// ---------------------------------------------------------------------------------------------------------
// Block-per-variant scan.  (A warp-per-variant kernel reading global memory kept ~10^4 columns in flight, 240 MB at
// n = 1e5, so every second sweep missed L2: ncu showed 134 KB/variant of DRAM traffic.)  A CTA owns one variant at a
// time: its 2-bit column is brought into shared memory ONCE by a 1-D bulk copy (cp.async.bulk completing on an
// mbarrier; optionally the next column is prefetched into a second stage), both sweeps and every relative's genotype
// look-up read shared memory, and only kinship tables / rows of Y (L2-resident, a few MB) are gathered from global
// memory.  Variants are handed out by an atomic work counter (their cost varies >20x with allele frequency); each
// variant's arithmetic is independent of which CTA runs it, so results stay bit-reproducible.
// ---------------------------------------------------------------------------------------------------------
constexpr uint32_t kTruncated = 0xffffffffu;   // pcnt value of a variant whose missing-sample lists did not fit
constexpr int kQueueCap = 96;            // per-warp queue of missing-sample positions; flushed when nearly full
constexpr int kRelQueueCap = 64;         // per-warp queue of missing positions inside the table region (kept per variant)

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// post-flip 2-bit code of the sample at position k of a column held in shared memory
__device__ __forceinline__ uint32_t code_smem(const uint8_t *col, int64_t k, bool flip)
{
    uint32_t code = ((uint32_t)col[k >> 2] >> (2 * (k & 3))) & 3u;
    if (flip && !(code & 1u)) code ^= 2u;
    return code;
}

template <int kBlkWarps, bool kLists>
__global__ void __launch_bounds__(kBlkWarps * 32) scan_block_kernel(ScanParams P, int nstage, unsigned long long *work, int l2pf)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    const size_t stride = P.stride, sbytes = P.sbytes;
    uint8_t *cols = smem;
    constexpr int kQueueWords = kLists ? 0 : kQueueCap + kRelQueueCap;            // the list instantiation keeps no queues
    uint32_t *wq = reinterpret_cast<uint32_t *>(smem + (size_t)nstage * sbytes) + wib * kQueueWords;   // this warp's queues
    uint32_t *rq = wq + kQueueCap;
    double *s_sm = reinterpret_cast<double *>(smem + (size_t)nstage * sbytes + sizeof(uint32_t) * kBlkWarps * kQueueWords);
    double *s_quad = s_sm + kBlkWarps * kColsPerSlice;                           // [2][kBlkWarps]: kinship part, hom-minor diag part
    int *s_cnt = reinterpret_cast<int *>(s_quad + 2 * kBlkWarps);                // [4][kBlkWarps]
    unsigned long long *s_work = reinterpret_cast<unsigned long long *>(s_cnt + 4 * kBlkWarps);   // [2]
    uint64_t *full = reinterpret_cast<uint64_t *>(s_work + 2);                   // [2]

    if (tid == 0) {
        for (int s = 0; s < 2; s++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&full[s])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](unsigned long long v, int stage) {          // thread 0 only
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&full[stage])),
                     "r"((uint32_t)sbytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_addr(cols + (size_t)stage * sbytes)),
                     "l"(P.packed + (size_t)v * stride), "r"((uint32_t)sbytes), "r"(smem_addr(&full[stage]))
                     : "memory");
    };
    auto wait_full = [&](int stage, uint32_t parity) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "SCAN_WAIT:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra SCAN_DONE;\n"
            "bra SCAN_WAIT;\n"
            "SCAN_DONE:\n"
            "}\n" ::"r"(smem_addr(&full[stage])), "r"(parity) : "memory");
    };

    const unsigned long long pmax = P.vlist ? *P.vcount : (unsigned long long)P.p;
    auto variant_of = [&](unsigned long long k) { return P.vlist ? (unsigned long long)P.vlist[k] : k; };
    if (tid == 0) {
        const unsigned long long v = atomicAdd(work, 1ull);
        s_work[0] = v;
        if (v < pmax) issue(variant_of(v), 0);
    }
    __syncthreads();
    unsigned long long cur = s_work[0];
    int stage = 0;
    uint32_t parity = 0u;                                        // bit s = phase parity of full[s]
    const int nvec16 = (int)(stride >> 4);
    const uint32_t *rmask32 = reinterpret_cast<const uint32_t *>(P.relmask);
    const int kin_words = (int)P.kin_words, tab_words = (int)P.tab_words, x_words = (int)P.x_words;
    const int n32 = (int)P.n;
    const int tail_word = n32 >> 4;                              // first word with padding positions
    const int tab_vec = (tab_words + 3) >> 2;                   // 64-sample vectors covering the table words
    const uint32_t tail_mask = (1u << (2 * (n32 & 15))) - 1u;    // valid codes of that word
    const int ycol = lane & 15;
    const uint32_t tab_samples = (uint32_t)tab_words * 16u;

    for (unsigned it = 1; cur < pmax; it++) {
        if (nstage == 2 && tid == 0) {                           // prefetch the next variant's column
            const unsigned long long v = atomicAdd(work, 1ull);
            s_work[it & 1] = v;
            if (v < pmax) issue(variant_of(v), stage ^ 1);
        }
        if (nstage == 1 && l2pf && tid == 0) {
            // one stage: draw the next variant now and pull its column into L2 while this one is scanned — the bulk copy
            // into shared memory issued at the end of the iteration then waits an L2 round trip, not an HBM one
            const unsigned long long v = atomicAdd(work, 1ull);
            s_work[it & 1] = v;
            if (v < pmax)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(P.packed + (size_t)variant_of(v) * stride),
                             "r"((uint32_t)stride) : "memory");
        }
        const int64_t i = (int64_t)variant_of(cur);
        // kLists: the missing samples were listed by the contraction, so their count — and with it the flip — is known
        // before the column is touched: one orientation, no detection, no queues.
        bool skip = false;
        int cm = 0;
        double sm = 0.0;
        if constexpr (kLists) {
            const uint32_t c = P.pcnt[i];
            skip = c == kTruncated;
            cm = skip ? 0 : (int)c;
            // column sums of Y over the listed missing samples (two rows per warp instruction, four in flight): they do
            // not need the column, so they are gathered while its bulk copy is still in flight
            if (ycol < P.ncol) {
                const uint32_t *L = P.plist + (size_t)i * (size_t)P.pcap;
#pragma unroll 4
                for (int e = 2 * wib + (lane >> 4); e < cm; e += 2 * kBlkWarps)
                    sm += __ldg(P.Y + (int64_t)__ldg(L + e) * P.ldY + ycol);
            }
        }
        wait_full(stage, (parity >> stage) & 1u);
        parity ^= 1u << stage;
        const uint8_t *col = cols + (size_t)stage * sbytes;

        // ---- sweep 1 (whole column, 64 samples per lane and step).  The sum of the variant's codes S = c1 + 2 c2 + 3 cm
        //      comes EXACTLY from the tensor-core contraction (its constant-1 column), so no genotype popcounts are
        //      needed here: the sweep finds the missing samples (cm; column sums of Y over them — positions go to the
        //      warp's queue, a flush reads two rows of Y per warp instruction (half-warp h: entries h, h+2, ...; lane c:
        //      column c) and keeps the positions inside the table region in a second queue for the kinship
        //      correction) and the sum of Sigma_i[j,j] over hom-minor carriers (g^2 - g = 2 [g == 2], the part of
        //      g' diag(Sigma_i) g that is not linear in the codes).  The latter needs the flip, which needs cm: it is
        //      computed under the guess cm = 0 and redone in the rare case that the missing count changes the decision.
        const long long S = P.codesum[i * 16 + 15];
        double hom = 0.0;
        int qn = 0, rn = 0;
        bool rq_over = false;
        const uint4 *v4 = reinterpret_cast<const uint4 *>(col);
        auto flush_missing = [&]() {
            __syncwarp();
            if (ycol < P.ncol) {
#pragma unroll 4
                for (int e = lane >> 4; e < qn; e += 2) sm += __ldg(P.Y + (int64_t)wq[e] * P.ldY + ycol);
            }
            for (int e0 = 0; e0 < qn; e0 += 32) {                // warp-uniform trip count
                const uint32_t pos = e0 + lane < qn ? wq[e0 + lane] : 0xffffffffu;
                const bool keep = pos < tab_samples;
                const uint32_t bal = __ballot_sync(0xffffffffu, keep);
                if (rn + __popc(bal) > kRelQueueCap) { rq_over = true; break; }
                if (keep) rq[rn + __popc(bal & ((1u << lane) - 1u))] = pos;
                rn += __popc(bal);
            }
            __syncwarp();
            qn = 0;
        };
        // hom-minor carriers of one 64-sample vector under orientation `fl`: raw code 0 (flip) or raw code 2 (no flip);
        // the four words are walked together so that four independent gathers of Sigma_i[j,j] are in flight
        int hom_cnt = 0;                                             // hom-minor carriers in the uniform-diagonal region
        auto hom_vec = [&](const uint4 &q4, int k, bool fl, double &acc) {
            const uint32_t gm = fl ? 0xffffffffu : 0u;
            uint32_t h0 = ((q4.x >> 1) ^ gm) & ~q4.x & 0x55555555u, h1 = ((q4.y >> 1) ^ gm) & ~q4.y & 0x55555555u;
            uint32_t h2 = ((q4.z >> 1) ^ gm) & ~q4.z & 0x55555555u, h3 = ((q4.w >> 1) ^ gm) & ~q4.w & 0x55555555u;
            const int kw = 4 * k;
            if (kw + 3 >= tail_word) {                               // padding positions hold code 0
                h0 &= kw > tail_word ? 0u : (kw == tail_word ? tail_mask : 0xffffffffu);
                h1 &= kw + 1 > tail_word ? 0u : (kw + 1 == tail_word ? tail_mask : 0xffffffffu);
                h2 &= kw + 2 > tail_word ? 0u : (kw + 2 == tail_word ? tail_mask : 0xffffffffu);
                h3 &= kw + 3 > tail_word ? 0u : (kw + 3 == tail_word ? tail_mask : 0xffffffffu);
            }
            if (k >= P.uvec) {                                       // one diagonal value from here on: count, do not gather
                hom_cnt += (__popc(h0) + __popc(h1)) + (__popc(h2) + __popc(h3));
                return;
            }
            const double *dg = P.diag + (int64_t)kw * 16;
            while (h0 | h1 | h2 | h3) {
                double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
                if (h0) { v0 = __ldg(dg + ((__ffs(h0) - 1) >> 1)); h0 &= h0 - 1u; }
                if (h1) { v1 = __ldg(dg + 16 + ((__ffs(h1) - 1) >> 1)); h1 &= h1 - 1u; }
                if (h2) { v2 = __ldg(dg + 32 + ((__ffs(h2) - 1) >> 1)); h2 &= h2 - 1u; }
                if (h3) { v3 = __ldg(dg + 48 + ((__ffs(h3) - 1) >> 1)); h3 &= h3 - 1u; }
                acc += (v0 + v1) + (v2 + v3);
            }
        };
        // both imputation modes flip iff the mean dosage exceeds 1 when nothing is missing
        bool guess = S > (long long)P.n;
        if constexpr (kLists) {
            const long long gs = S - 3ll * cm, gn = (long long)P.n - cm;
            if (P.impute == STAAR_IMPUTE_MEAN) guess = gs > gn;
            else guess = gs + (gs > gn ? 2 : 0) * (long long)cm > (long long)P.n;
            // hom-minor term beyond the table words: read straight from global memory (only the kinship prefix of the
            // column is staged in shared memory); the table words get it inside the kinship sweep below (one read)
            const int nv = skip ? 0 : nvec16;
            const uint4 *g4 = reinterpret_cast<const uint4 *>(P.packed + (size_t)i * stride);
            int k = tab_vec + tid;
            if (sbytes == stride) {                                                  // short columns: everything is staged
                for (; k < nv; k += kBlkWarps * 32) hom_vec(v4[k], k, guess, hom);
            }
            for (; k + 3 * kBlkWarps * 32 < nv; k += 4 * kBlkWarps * 32) {           // four 16-byte loads in flight
                const uint4 qa = __ldg(g4 + k), qb = __ldg(g4 + k + kBlkWarps * 32), qc = __ldg(g4 + k + 2 * kBlkWarps * 32),
                            qd = __ldg(g4 + k + 3 * kBlkWarps * 32);
                hom_vec(qa, k, guess, hom);
                hom_vec(qb, k + kBlkWarps * 32, guess, hom);
                hom_vec(qc, k + 2 * kBlkWarps * 32, guess, hom);
                hom_vec(qd, k + 3 * kBlkWarps * 32, guess, hom);
            }
            for (; k < nv; k += kBlkWarps * 32) hom_vec(__ldg(g4 + k), k, guess, hom);
            sm += __shfl_xor_sync(0xffffffffu, sm, 16);
            if (lane < kColsPerSlice) s_sm[wib * kColsPerSlice + lane] = sm;
        } else {
            for (int k0 = wib * 32; k0 < nvec16; k0 += kBlkWarps * 32) {
                const int k = k0 + lane;
                uint4 q4 = make_uint4(0u, 0u, 0u, 0u);
                if (k < nvec16) { q4 = v4[k]; hom_vec(q4, k, guess, hom); }
                const uint32_t m0 = q4.x & (q4.x >> 1) & 0x55555555u, m1 = q4.y & (q4.y >> 1) & 0x55555555u;
                const uint32_t m2 = q4.z & (q4.z >> 1) & 0x55555555u, m3 = q4.w & (q4.w >> 1) & 0x55555555u;
                // bit 2f + w of a half = sample f of word w (low half: words 0,1; high half: words 2,3)
                uint32_t blo = m0 | (m1 << 1), bhi = m2 | (m3 << 1);
                uint32_t bal = __ballot_sync(0xffffffffu, (blo | bhi) != 0u);
                if (bal) {
                    cm += __popc(blo) + __popc(bhi);
                    const uint32_t first = (uint32_t)k * 64u;
                    do {                                          // round r: the r-th missing sample of every lane
                        if (blo | bhi) {
                            int b, half;
                            if (blo) { b = __ffs(blo) - 1; blo &= blo - 1u; half = 0; }
                            else { b = __ffs(bhi) - 1; bhi &= bhi - 1u; half = 2; }
                            wq[qn + __popc(bal & ((1u << lane) - 1u))] = first + (uint32_t)((half + (b & 1)) * 16 + (b >> 1));
                        }
                        qn += __popc(bal);
                        if (qn > kQueueCap - 32) flush_missing();
                        bal = __ballot_sync(0xffffffffu, (blo | bhi) != 0u);
                    } while (bal);
                }
            }
            if (qn > 0) flush_missing();
            sm += __shfl_xor_sync(0xffffffffu, sm, 16);            // the two half-warps' partial sums (fixed order)
            cm = __reduce_add_sync(0xffffffffu, cm);
            if (lane == 0) { s_cnt[wib] = cm; s_cnt[kBlkWarps + wib] = rq_over; }
            if (lane < kColsPerSlice) s_sm[wib * kColsPerSlice + lane] = sm;
            __syncthreads();
            cm = 0;
            int over = 0;
#pragma unroll
            for (int w = 0; w < kBlkWarps; w++) { cm += s_cnt[w]; over |= s_cnt[kBlkWarps + w]; }
            rq_over = over != 0;                                  // block-wide: some warp could not keep its positions
        }
        // flip decision in exact integers (every thread; the FP64 divisions stay off the critical path):
        //   sum = c1 + 2 c2 = S - 3 cm,  num = n - cm
        //   mean : AF = sum/2/num > 0.5  <=>  sum > num          (|AF - 0.5| >= 1/(4 num) >> ulp, division is monotone)
        //   minor: fill = AF > 0.5 ? 2 : 0;  AF' = (sum + fill cm)/2/n > 0.5  <=>  sum + fill cm > n
        const long long gsum = S - 3ll * cm, gnum = (long long)P.n - cm;
        const PackedCounts pc{(int)gsum, 0, cm};                   // flip_from_counts only uses c1 + 2 c2 and cm
        bool flip, mu_zero;
        if (P.impute == STAAR_IMPUTE_MEAN) {
            flip = gsum > gnum;
            mu_zero = flip ? (gsum == 2 * gnum) : (gsum == 0);      // mu = 2 AF or 2 - 2 AF
        } else {
            const long long fill = gsum > gnum ? 2 : 0;
            flip = gsum + fill * cm > (long long)P.n;
            mu_zero = flip ? (fill == 2) : (fill == 0);            // mu = fill or 2 - fill
        }
        auto mu_of = [&]() { return flip_from_counts(pc, P.n, P.impute).mu; };   // rare paths only
        if (guess != flip) {                                       // block-uniform and rare: the missing count decided
            hom = 0.0;
            hom_cnt = 0;
            for (int k0 = wib * 32; k0 < nvec16; k0 += kBlkWarps * 32)
                if (k0 + lane < nvec16) hom_vec(v4[k0 + lane], k0 + lane, flip, hom);
        }

        // ---- sweep 2: kinship off-diagonal part of g' Sigma_i g over the related prefix of the column.
        double quad = 0.0;
        const uint32_t *col32 = reinterpret_cast<const uint32_t *>(col);
        const uint32_t flipmask = flip ? 0x55555555u : 0u;
        // correction for a missing related sample j of a table word (imputed to mu; the tables see it as 0)
        auto related_missing = [&](int64_t j, double mu) {
            const NullSparse::SampleInfo si = P.sinfo[j];
            double sacc = 0.0;
            for (uint32_t t = 0; t < si.off_count; t++) {
                const NullSparse::OffEntry oe = P.off_ent[si.off_start + t];
                const uint32_t c = code_smem(col, oe.col, flip);
                // a pair of two missing samples is counted once, from its smaller position
                const double v = c == 3u ? ((int64_t)oe.col > j ? mu : 0.0) : (double)c;
                sacc += oe.val * v;
            }
            quad += 2.0 * mu * sacc;
        };
        // (a) table words [0, tab_words): full words of real samples, four words per lane and step.  A lane only
        //     works on the words in which a group holds two related carriers (or, in a cross word, two anywhere):
        //     the codes of a 4-sample group index that group's table.
        //     The kernel is latency-bound, so the look-ups of two words (8 groups) are issued together: the table of a
        //     group sits at a fixed stride (no descriptor fetch in front of it) and the related-slot masks of a lane's four
        //     words are one 128-bit load issued before anything else.
        const int tab_vec_eff = skip ? 0 : tab_vec;
        for (int k0 = wib * 32; k0 < tab_vec_eff; k0 += kBlkWarps * 32) {
            const int k = k0 + lane;
            uint4 q4 = make_uint4(0u, 0u, 0u, 0u), km4 = make_uint4(0u, 0u, 0u, 0u);
            if (k < tab_vec) {
                km4 = __ldg(reinterpret_cast<const uint4 *>(P.kmask) + k);
                q4 = v4[k];
                if constexpr (kLists) hom_vec(q4, k, flip, hom);
            }
            const uint32_t wv[4] = {q4.x, q4.y, q4.z, q4.w}, kmv[4] = {km4.x, km4.y, km4.z, km4.w};
            uint32_t wzv[4], twov[4];
            bool crossv[4];
#pragma unroll
            for (int wi = 0; wi < 4; wi++) {
                const int kw = 4 * k + wi;
                const uint32_t w = wv[wi], km = kmv[wi];             // km == 0 beyond the table words
                const uint32_t lo = w & 0x55555555u;
                const uint32_t mis = lo & (w >> 1);
                const uint32_t wf = w ^ ((~lo & flipmask) << 1);     // g -> 2 - g on the codes 0 and 2
                const uint32_t wz = wf & ~(mis * 3u) & km;           // related, non-missing codes (tables hold such pairs)
                const uint32_t hit = (wz | (wz >> 1)) & 0x55555555u; // carriers: bits 0,2,4,6 of every group byte
                // per byte: clear the lowest carrier bit (bit 7 guards the borrow): non-zero byte <=> >= 2 carriers in the group
                twov[wi] = hit & ((hit | 0x80808080u) - 0x01010101u);
                crossv[wi] = kw < x_words && (hit & (hit - 1u)) != 0u;   // >= 2 carriers anywhere in a cross word
                wzv[wi] = wz;
                if (rq_over && !mu_zero) {                           // slow path: the position queues overflowed
                    uint32_t mk = mis & km;
                    if (mk) {
                        const double mu = mu_of();
                        while (mk) {
                            const int b = __ffs(mk) - 1;
                            mk &= mk - 1u;
                            related_missing((int64_t)kw * 16 + (b >> 1), mu);
                        }
                    }
                }
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                if (!__any_sync(0xffffffffu, (twov[2 * h] | twov[2 * h + 1]) != 0u)) continue;
                double v[2][4];
#pragma unroll
                for (int wj = 0; wj < 2; wj++) {
                    const int wi = 2 * h + wj;
                    const uint32_t gbase = (uint32_t)(4 * (4 * k + wi)) << 8;
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        v[wj][g] = 0.0;
                        if (twov[wi] & (0xffu << (8 * g))) v[wj][g] = __ldg(P.lut + gbase + (g << 8) + ((wzv[wi] >> (8 * g)) & 0xffu));
                    }
                }
                quad += ((v[0][0] + v[0][1]) + (v[0][2] + v[0][3])) + ((v[1][0] + v[1][1]) + (v[1][2] + v[1][3]));
            }
#pragma unroll
            for (int wi = 0; wi < 4; wi++)
                if (crossv[wi]) {                                    // families of 5..16 members: member x other chunk
                    const int kw = 4 * k + wi;
                    const uint32_t wz = wzv[wi];
                    const uint2 wd = __ldg(reinterpret_cast<const uint2 *>(P.wdesc) + kw);
                    const uint2 *tp = reinterpret_cast<const uint2 *>(P.xterms) + wd.x;
                    for (uint32_t t = 0; t < wd.y; t++) {
                        const uint2 tr = __ldg(tp + t);
                        const uint32_t idx = (wz >> (tr.y & 0xffu)) & ((tr.y >> 8) & 0xffu);
                        const uint32_t gj = (wz >> ((tr.y >> 16) & 0xffu)) & 3u;
                        if (gj != 0u && idx != 0u) quad += (double)gj * __ldg(P.lut + tr.x + idx);
                    }
                }
        }
        if constexpr (kLists) {
            if (!mu_zero && cm > 0) {                                // listed missing samples inside the table region
                const double mu = mu_of();
                const uint32_t *L = P.plist + (size_t)i * (size_t)P.pcap;
                for (int e = tid; e < cm; e += kBlkWarps * 32) {
                    const uint32_t pos = __ldg(L + e);
                    if (pos < tab_samples) related_missing((int64_t)pos, mu);
                }
            }
        } else if (!rq_over && !mu_zero && rn > 0) {                 // missing samples queued in sweep 1 (this warp's)
            const double mu = mu_of();
            for (int e = lane; e < rn; e += 32) related_missing((int64_t)rq[e], mu);
        }
        // (b) generic words [tab_words, kin_words): per-carrier records (families of more than 16 members)
        for (int k = tab_words + tid; k < (skip ? 0 : kin_words); k += kBlkWarps * 32) {
            const uint32_t w = col32[k];
            const uint32_t lo = w & 0x55555555u;
            const uint32_t mis = lo & (w >> 1);
            uint32_t wf = w ^ ((~lo & flipmask) << 1);
            if (k >= tail_word) wf &= (k == tail_word) ? tail_mask : 0u;   // padding codes would flip to 2
            uint32_t sel = (wf | (wf >> 1)) & 0x55555555u & __ldg(rmask32 + k);
            if (mu_zero) sel &= ~mis;
            const double mu = sel ? mu_of() : 0.0;
            const int64_t base = (int64_t)k * 16;
            while (sel) {
                const int b = __ffs(sel) - 1;
                sel &= sel - 1;
                const NullSparse::CarrierRec rr = P.rel[base + (b >> 1)];
                double sacc = 0.0;
                for (uint32_t t = 0; t < rr.count; t++) {
                    double sv; int32_t sc;
                    if (t < 2u) { sv = rr.val[t]; sc = rr.col[t]; }
                    else { const NullSparse::OffEntry oe = P.up_ent[rr.start + t]; sv = oe.val; sc = oe.col; }
                    sacc += sv * code_value(code_smem(col, sc, flip), mu);
                }
                quad += 2.0 * code_value((wf >> b) & 3u, mu) * sacc;
            }
        }
        hom += P.dval * (double)hom_cnt;
        quad = warp_sum(quad);
        hom = warp_sum(hom);
        if (lane == 0) { s_quad[wib] = quad; s_quad[kBlkWarps + wib] = hom; }
        __syncthreads();                      // also: every read of this stage's column is complete
        if (skip) {
            // nothing to write: the detection path fills this variant's outputs
        } else if (tid < kColsPerSlice) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < kBlkWarps; w++) t += s_sm[w * kColsPerSlice + tid];
            if (tid == kColsPerSlice - 1) {                              // slot 15: hom-minor diag sum
                t = 0.0;
#pragma unroll
                for (int w = 0; w < kBlkWarps; w++) t += s_quad[kBlkWarps + w];
                P.Sm[i * kColsPerSlice + tid] = t;
            } else {
                P.Sm[i * kColsPerSlice + tid] = (tid < P.ncol) ? t : 0.0;
            }
        } else if (tid == (kBlkWarps > 1 ? 32 : kColsPerSlice)) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < kBlkWarps; w++) t += s_quad[w];
            const VariantFlip vf = flip_from_counts(pc, P.n, P.impute);
            // bit 1: every genotype is 0 after flip and imputation (the reference's exact Cov == 0 branch):
            // all non-missing codes are 0 (sum == 0) or all are 2 (sum == 2 num)
            const int mono = (gsum == 0 || gsum == 2 * gnum) ? 2 : 0;
            P.AF[i] = vf.af; P.MAF[i] = vf.maf; P.flip[i] = (flip ? 1 : 0) | mono; P.mu[i] = vf.mu; P.quad[i] = t;
        }
        if (nstage == 2) {
            cur = s_work[it & 1];
            stage ^= 1;
            __syncthreads();                  // s_sm / s_quad / s_cnt are rewritten by the next variant
        } else {
            __syncthreads();
            if (tid == 0) {
                unsigned long long v;
                if (l2pf) v = s_work[it & 1];
                else { v = atomicAdd(work, 1ull); s_work[it & 1] = v; }
                if (v < pmax) issue(variant_of(v), 0);
            }
            __syncthreads();
            cur = s_work[it & 1];
        }
    }
}

struct FinParams {
    int64_t p; int q;
    const int32_t *flip;
    const double *mu, *quad, *Sm;
    const long long *hi, *lo, *tot;
    const double *scale, *cov;
    double *Score, *Score_se, *pl, *Est, *Est_se;
    const long long *vlist; const unsigned long long *vcount;     // list mode (see ScanParams)
};

// The contraction delivers, per variant and column c, the exact integers R_c = sum_j code_j X_jc over the RAW codes
// (missing = 3); the scan delivers the FP64 column
// sums Sm_c over missing samples, the flip decision, the imputation value mu and the kinship off-diagonal part.
// With T_c = sum_j X_jc (exact, set-up time) the genotype the reference tests (flipped, imputed) gives
//   no flip: G'y_c = R_c - 3 M_c + mu M_c
//   flip   : G'y_c = 2 T_c - R_c + M_c + mu M_c
// (integer parts combined exactly in 64-bit, M_c = Sm_c).  A variant whose genotypes are all 0 after flip and
// imputation takes the reference's exact Cov == 0 branch from the integer counts (bit 1 of flip[]).
__global__ void __launch_bounds__(128) finalize_packed_kernel(FinParams P)
{
    const int64_t count = P.vlist ? (int64_t)*P.vcount : P.p;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = P.vlist ? (int64_t)P.vlist[k] : k;
        const int fl = P.flip[i];
        const bool flip = (fl & 1) != 0, mono = (fl & 2) != 0;
        const double mu = P.mu[i];
        if (mono) {
            P.Score[i] = 0.0;
            single_trait_stats(0.0, 0.0, P.Score_se[i], P.pl[i], P.Est[i], P.Est_se[i]);
            continue;
        }
        double L[kColsPerSlice];
        const double wm = flip ? mu + 1.0 : mu - 3.0;
        for (int c = 0; c <= P.q; c++) {
            long long hi = P.hi[i * 16 + c], lo = P.lo[i * 16 + c];
            if (flip) { hi = 2 * P.tot[c] - hi; lo = 2 * P.tot[kColsPerSlice + c] - lo; }
            L[c] = ((double)hi * 16777216.0 + (double)lo) * P.scale[c] + wm * P.Sm[i * kColsPerSlice + c];
        }
        const double U = L[0];
        double tct = 0.0;
        for (int b = 0; b < P.q; b++) {
            double xb = 0.0;
            for (int a = 0; a < P.q; a++) xb += L[1 + a] * P.cov[a + b * P.q];
            tct += xb * L[1 + b];
        }
        // g' Sigma_i g = sum_j s_jj g_j (non-missing; the contraction's diag column, same algebra as L_c above)
        //              + 2 sum over hom-minor carriers of s_jj   (g^2 - g = 2 [g == 2]; scan, slot 15 of Sm)
        //              + mu^2 sum over missing of s_jj + kinship off-diagonal part (scan)
        const int d = P.q + 1;
        long long dhi = P.hi[i * 16 + d], dlo = P.lo[i * 16 + d];
        if (flip) { dhi = 2 * P.tot[d] - dhi; dlo = 2 * P.tot[kColsPerSlice + d] - dlo; }
        const double smd = P.Sm[i * kColsPerSlice + d];
        const double lin = ((double)dhi * 16777216.0 + (double)dlo) * P.scale[d] + (flip ? 1.0 : -3.0) * smd;
        const double quad = lin + 2.0 * P.Sm[i * kColsPerSlice + 15] + mu * mu * smd + P.quad[i];
        const double C = quad - tct;
        P.Score[i] = U;
        single_trait_stats(U, C, P.Score_se[i], P.pl[i], P.Est[i], P.Est_se[i]);
    }
}

}  // namespace

int launch_scan_packed(staar_ctx *ctx, const uint8_t *d_packed, size_t stride, int64_t n, int64_t p, int impute,
                       const long long *d_lo, double *dAF, double *dMAF, int32_t *dflip, double *d_mu, double *d_quad,
                       double *d_Sm, const long long *d_vlist, const unsigned long long *d_vcount, const uint32_t *d_plist,
                       const uint32_t *d_pcnt, int pcap)
{
    if (p == 0) return STAAR_OK;
    const NullSparse &ns = ctx->ns;
    const bool lists = d_plist != nullptr;
    if (n >= (1ll << 30)) { set_error("packed scan: n too large for 30-bit sample indices"); return STAAR_ERR_ARG; }
    ScanParams P;
    P.packed = d_packed; P.stride = stride; P.n = n; P.p = p; P.impute = impute;
    // list mode stages only the kinship prefix of a column (table words, generic kinship words; at least one vector)
    const size_t kin_bytes = (size_t)std::max<int64_t>(((ns.tab_words + 3) / 4) * 16, ((ns.kin_words * 4 + 15) / 16) * 16);
    // Shorter columns are staged whole: the gain is occupancy (n = 1e5: 25 -> 14.5 KB per block), and below ~20 KB a whole
    // column already leaves room for 10+ blocks per SM while the direct loads of a short tail are latency-bound
    // (sparse GRM, prefix vs whole column: n = 2e4 5.15 vs 4.06 ms per 1 M variants, n = 6e4 4.95 vs 4.20 ms per 400 k,
    // n = 1e5 4.52 vs 5.03 ms per 250 k, n = 4e5 6.18 vs 8.92 ms per 60 k).
    const size_t sbytes = (lists && stride > 20000) ? std::min(stride, std::max<size_t>(kin_bytes, 16)) : stride;
    P.sbytes = sbytes;
    P.relmask = ns.relmask; P.rel = ns.rel; P.up_ent = ns.up_ent;
    P.tab_words = ns.tab_words; P.kin_words = ns.kin_words; P.x_words = ns.x_words; P.kmask = ns.kmask; P.wdesc = ns.wdesc;
    P.xterms = ns.xterms; P.lut = ns.lut; P.sinfo = ns.sinfo_p; P.off_ent = ns.off_ent_p; P.diag = ns.diag_p;
    P.uvec = ns.diag_uniform_vec >= 0 ? (int)ns.diag_uniform_vec : INT_MAX; P.dval = ns.diag_uniform_val;
    P.Y = ns.Yp; P.ldY = ns.ldY; P.ncol = (int)ns.q + 2;
    P.codesum = d_lo; P.vlist = d_vlist; P.vcount = d_vcount;
    P.plist = d_plist; P.pcnt = d_pcnt; P.pcap = pcap;
    P.AF = dAF; P.MAF = dMAF; P.flip = dflip; P.mu = d_mu; P.quad = d_quad; P.Sm = d_Sm;
    // block-per-variant kernel, column in shared memory.  CTA width: 4 warps put more variants in flight per SM (7 vs 4 at
    // n = 1e5: 6.75 -> 6.63 ms per 250 k variants); when long columns leave room for few CTAs (n = 4e5: 2 per SM), 8 warps
    // per CTA keep more warps resident (6.1 M -> 8.5 M variants/s).
    auto extra_for = [lists](int warps) {
        return sizeof(uint32_t) * warps * (lists ? 0 : kQueueCap + kRelQueueCap) + sizeof(double) * (warps * kColsPerSlice + 2 * warps) +
               sizeof(int) * 4 * warps + 64;
    };
    const size_t smem_max = 227 * 1024;
    // Short columns (e.g. BASELINE config 1, n = 1e4): per-variant block overheads dominate, so the block shrinks to one
    // or two warps and up to 32 blocks share an SM (measured per 400 k variants, sparse GRM: n = 2e4: 2.37 / 3.15 / 3.78 ms
    // with 1 / 2 / 4 warps; n = 4e4: 5.09 / 4.60 / 5.22 ms; n = 1e5: 27.9 / 13.8 / 10.5 ms).
    static const int force_warps = getenv("STAAR_B200_SCAN_WARPS") ? atoi(getenv("STAAR_B200_SCAN_WARPS")) : 0;
    int warps = (6 * (sbytes + extra_for(4) + 1024) <= 228 * 1024) ? 4 : 8;
    if (stride <= 6144) warps = 1;
    else if (stride <= 14336) warps = 2;
    if (force_warps == 1 || force_warps == 2 || force_warps == 4 || force_warps == 8) warps = force_warps;
    const size_t extra = extra_for(warps);
    // The kernel is latency-bound (dependent gathers between block barriers), so resident warps matter more than
    // overlapping the column copy: one column stage (measured at n = 1e5: 6.76 ms vs 7.13 ms with two stages per
    // 250 k variants); STAAR_B200_SCAN_STAGES=2 forces the double-buffered variant for experiments.
    static const int force = getenv("STAAR_B200_SCAN_STAGES") ? atoi(getenv("STAAR_B200_SCAN_STAGES")) : 0;
    int nstage = (sbytes + extra <= smem_max) ? 1 : 0;
    if (nstage == 1 && force == 2 && 2 * sbytes + extra <= smem_max) nstage = 2;
    if (nstage > 0) {
        const size_t smem = (size_t)nstage * sbytes + extra;
        auto kern = lists ? (warps == 1 ? scan_block_kernel<1, true> : warps == 2 ? scan_block_kernel<2, true> :
                             warps == 4 ? scan_block_kernel<4, true> : scan_block_kernel<8, true>)
                          : (warps == 1 ? scan_block_kernel<1, false> : warps == 2 ? scan_block_kernel<2, false> :
                             warps == 4 ? scan_block_kernel<4, false> : scan_block_kernel<8, false>);
        STAAR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        STAAR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, warps * 32, smem));
        if (per_sm < 1) { set_error("packed scan: kernel does not fit on an SM (smem %zu B)", smem); return STAAR_ERR_CUDA; }
        STAAR_TRY(ctx->ctr.reserve(sizeof(unsigned long long)));
        STAAR_CUDA(cudaMemsetAsync(ctx->ctr.p, 0, sizeof(unsigned long long), ctx->stream));
        const int64_t blocks = std::min<int64_t>(p, (int64_t)ctx->sm_count * per_sm);
        // STAAR_B200_SCAN_L2PF=1 switches an L2 prefetch of the next column on (experiments)
        // (measured: no gain — 6.85 vs 6.69 ms per 250 k variants at n = 1e5 — so it is off unless asked for)
        static const int l2pf = getenv("STAAR_B200_SCAN_L2PF") ? atoi(getenv("STAAR_B200_SCAN_L2PF")) : 0;
        kern<<<(int)blocks, warps * 32, smem, ctx->stream>>>(P, nstage, ctx->ctr.as<unsigned long long>(), l2pf);
    } else {
        set_error("packed scan: a column of %zu bytes does not fit in shared memory (n > ~9e5): use the FP64 entry points", stride);
        return STAAR_ERR_ARG;
    }
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

namespace {
// Segment records of the contraction kernel -> positions.  One warp per variant: records are read 32 at a time, every
// lane pops the bits of its record, positions are written in record order (exclusive prefix sum of the popcounts).
// A variant whose lists were truncated, or whose positions do not fit, is marked kTruncated and appended to the work
// list of the detection path.
__global__ void __launch_bounds__(256) expand_lists_kernel(const uint2 *mlist, const uint32_t *mcnt, int mcap, int64_t p, uint32_t *plist,
                                                           uint32_t *pcnt, int pcap, long long *vlist, unsigned long long *vcount)
{
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= p) return;
    const uint32_t c0 = mcnt[2 * i], c1 = mcnt[2 * i + 1];
    bool fits = c0 <= (uint32_t)mcap && c1 <= (uint32_t)mcap;
    const uint2 *L0 = mlist + (size_t)(2 * i) * (size_t)mcap, *L1 = L0 + mcap;
    uint32_t *out = plist + (size_t)i * (size_t)pcap;
    const int nrec = fits ? (int)(c0 + c1) : 0;
    uint32_t done = 0u;
    for (int base = 0; base < nrec && fits; base += 32) {              // warp-uniform
        const int e = base + lane;
        uint2 rec = make_uint2(0u, 0u);
        if (e < nrec) rec = (uint32_t)e < c0 ? L0[e] : L1[e - (int)c0];
        const uint32_t k = __popc(rec.y);
        uint32_t incl = k;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (done + total > (uint32_t)pcap) { fits = false; break; }    // more missing samples than the position list holds
        uint32_t at = done + incl - k;
        while (rec.y) {
            const int b = __ffs(rec.y) - 1;
            rec.y &= rec.y - 1u;
            out[at++] = rec.x * 32u + (uint32_t)((b & 1) * 16 + (b >> 1));
        }
        done += total;
    }
    if (lane == 0) {
        pcnt[i] = fits ? done : kTruncated;
        if (!fits) vlist[atomicAdd(vcount, 1ull)] = i;
    }
}
}  // namespace

int launch_expand_lists(staar_ctx *ctx, const uint2 *d_mlist, const uint32_t *d_mcnt, int mcap, int64_t p, uint32_t *d_plist,
                        uint32_t *d_pcnt, int pcap, long long *d_vlist, unsigned long long *d_vcount)
{
    if (p == 0) return STAAR_OK;
    STAAR_CUDA(cudaMemsetAsync(d_vcount, 0, sizeof(unsigned long long), ctx->stream));
    expand_lists_kernel<<<(unsigned)((p * 32 + 255) / 256), 256, 0, ctx->stream>>>(d_mlist, d_mcnt, mcap, p, d_plist, d_pcnt, pcap,
                                                                                  d_vlist, d_vcount);
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

int launch_finalize_packed(staar_ctx *ctx, int64_t p, const int32_t *d_flip, const double *d_mu, const double *d_quad,
                           const double *d_Sm, const long long *d_hi, const long long *d_lo, double *dScore,
                           double *dScore_se, double *dpl, double *dEst, double *dEst_se, const long long *d_vlist,
                           const unsigned long long *d_vcount)
{
    if (p == 0) return STAAR_OK;
    const NullSparse &ns = ctx->ns;
    FinParams P;
    P.p = p; P.q = (int)ns.q; P.flip = d_flip; P.mu = d_mu; P.quad = d_quad; P.Sm = d_Sm;
    P.hi = d_hi; P.lo = d_lo; P.tot = ns.col_total; P.scale = ns.col_scale; P.cov = ns.cov;
    P.Score = dScore; P.Score_se = dScore_se; P.pl = dpl; P.Est = dEst; P.Est_se = dEst_se;
    P.vlist = d_vlist; P.vcount = d_vcount;
    int blocks = (int)std::min<int64_t>((p + 127) / 128, (int64_t)ctx->sm_count * 8);
    finalize_packed_kernel<<<blocks, 128, 0, ctx->stream>>>(P);
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

}  // namespace staar
