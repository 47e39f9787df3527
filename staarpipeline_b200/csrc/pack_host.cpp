// pack_host.cpp — host side of the genotype ingest: a `$dosage` block as R / SeqArray deliver it
// (FP64 matrix at the Rcpp boundary, R integer matrix, or raw bytes) -> 2-bit packed columns (include/staar_b200.h).
//
// The Rcpp boundary hands over 8 bytes per genotype; at n = 1e5 a 5 000-variant chunk is 4 GB, so this pass — not the
// GPU — bounds the end-to-end rate.  It is therefore threaded over variants and vectorised: with AVX-512 a vector of
// values is classified by four compares into k-masks, the low/high code bits of 16 (or 64) samples are two mask
// registers and BMI2 PDEP interleaves them into the 2-bit layout; a scalar version covers other CPUs.  Compiled by
// g++ (not nvcc) so that the target attributes and intrinsics are available; chosen at run time by CPUID.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <immintrin.h>
#include "../../include/staar_b200.h"

namespace staar {
void set_error(const char *fmt, ...);
}

namespace {

// ---------------------------------------------------------------------------------------------- scalar versions
template <typename T> inline unsigned code_of(T v, unsigned &ok);
template <> inline unsigned code_of<double>(double v, unsigned &ok)
{
    // code = [v==1] + 2[v==2] + 3[!(v>-1)]; NaN and any value <= -1 are missing (src/matrix_flip_mean.cpp:26)
    const unsigned is1 = v == 1.0, is2 = v == 2.0, is0 = v == 0.0, miss = !(v > -1.0);
    ok &= (is0 | is1 | is2 | miss);
    return is1 + 2u * is2 + 3u * miss;
}
template <> inline unsigned code_of<int32_t>(int32_t v, unsigned &ok)
{
    const unsigned is1 = v == 1, is2 = v == 2, is0 = v == 0, miss = v <= -1;      // NA_integer_ = INT_MIN
    ok &= (is0 | is1 | is2 | miss);
    return is1 + 2u * is2 + 3u * miss;
}
template <> inline unsigned code_of<uint8_t>(uint8_t v, unsigned &ok)
{
    const unsigned is1 = v == 1, is2 = v == 2, is0 = v == 0, miss = v == 0xFF;    // SeqArray raw dosage: 0xFF = NA
    ok &= (is0 | is1 | is2 | miss);
    return is1 + 2u * is2 + 3u * miss;
}

template <typename T> bool pack_column_scalar(const T *g, int64_t n, uint8_t *dst, size_t stride, int64_t from_sample = 0)
{
    unsigned ok = 1;
    const int64_t b0 = from_sample / 4, n4 = n / 4;
    for (int64_t b = b0; b < n4; b++) {
        const T *s = g + 4 * b;
        dst[b] = (uint8_t)(code_of<T>(s[0], ok) | (code_of<T>(s[1], ok) << 2) | (code_of<T>(s[2], ok) << 4) | (code_of<T>(s[3], ok) << 6));
    }
    unsigned last = 0;
    for (int64_t j = 4 * n4; j < n; j++) last |= code_of<T>(g[j], ok) << (2 * (j & 3));
    memset(dst + n4, 0, stride - n4);
    if (n & 3) dst[n4] = (uint8_t)last;
    return ok != 0;
}

// ---------------------------------------------------------------------------------------------- AVX-512 versions
#if defined(__x86_64__)
#define STAAR_AVX512 __attribute__((target("avx512f,avx512bw,avx512vl,bmi2")))

STAAR_AVX512 bool pack_column_avx512(const double *g, int64_t n, uint8_t *dst, size_t stride)
{
    const __m512d one = _mm512_set1_pd(1.0), two = _mm512_set1_pd(2.0), zero = _mm512_setzero_pd(), neg1 = _mm512_set1_pd(-1.0);
    const int64_t n16 = n / 16;
    unsigned bad = 0;
    uint32_t *out = reinterpret_cast<uint32_t *>(dst);
    for (int64_t w = 0; w < n16; w++) {
        const __m512d a = _mm512_loadu_pd(g + 16 * w), b = _mm512_loadu_pd(g + 16 * w + 8);
        const unsigned m1 = _mm512_cmp_pd_mask(a, one, _CMP_EQ_OQ) | ((unsigned)_mm512_cmp_pd_mask(b, one, _CMP_EQ_OQ) << 8);
        const unsigned m2 = _mm512_cmp_pd_mask(a, two, _CMP_EQ_OQ) | ((unsigned)_mm512_cmp_pd_mask(b, two, _CMP_EQ_OQ) << 8);
        const unsigned m0 = _mm512_cmp_pd_mask(a, zero, _CMP_EQ_OQ) | ((unsigned)_mm512_cmp_pd_mask(b, zero, _CMP_EQ_OQ) << 8);
        const unsigned gt = _mm512_cmp_pd_mask(a, neg1, _CMP_GT_OQ) | ((unsigned)_mm512_cmp_pd_mask(b, neg1, _CMP_GT_OQ) << 8);
        const unsigned mm = ~gt & 0xFFFFu;                                       // NaN or <= -1
        bad |= ~(m0 | m1 | m2 | mm) & 0xFFFFu;
        out[w] = _pdep_u32(m1 | mm, 0x55555555u) | _pdep_u32(m2 | mm, 0xAAAAAAAAu);
    }
    // samples beyond the last full 16-sample word, and the zero padding of the stride
    const int64_t done = 16 * n16;
    std::vector<uint8_t> tail(stride - (size_t)(done / 4), 0);
    const bool ok_tail = pack_column_scalar<double>(g + done, n - done, tail.data(), tail.size());
    memcpy(dst + done / 4, tail.data(), tail.size());
    return bad == 0 && ok_tail;
}

STAAR_AVX512 bool pack_column_avx512(const int32_t *g, int64_t n, uint8_t *dst, size_t stride)
{
    const __m512i one = _mm512_set1_epi32(1), two = _mm512_set1_epi32(2), zero = _mm512_setzero_si512(), neg1 = _mm512_set1_epi32(-1);
    const int64_t n16 = n / 16;
    unsigned bad = 0;
    uint32_t *out = reinterpret_cast<uint32_t *>(dst);
    for (int64_t w = 0; w < n16; w++) {
        const __m512i a = _mm512_loadu_si512(g + 16 * w);
        const unsigned m1 = _mm512_cmpeq_epi32_mask(a, one), m2 = _mm512_cmpeq_epi32_mask(a, two), m0 = _mm512_cmpeq_epi32_mask(a, zero);
        const unsigned mm = _mm512_cmple_epi32_mask(a, neg1);
        bad |= ~(m0 | m1 | m2 | mm) & 0xFFFFu;
        out[w] = _pdep_u32(m1 | mm, 0x55555555u) | _pdep_u32(m2 | mm, 0xAAAAAAAAu);
    }
    const int64_t done = 16 * n16;
    std::vector<uint8_t> tail(stride - (size_t)(done / 4), 0);
    const bool ok_tail = pack_column_scalar<int32_t>(g + done, n - done, tail.data(), tail.size());
    memcpy(dst + done / 4, tail.data(), tail.size());
    return bad == 0 && ok_tail;
}

STAAR_AVX512 bool pack_column_avx512(const uint8_t *g, int64_t n, uint8_t *dst, size_t stride)
{
    const __m512i one = _mm512_set1_epi8(1), two = _mm512_set1_epi8(2), zero = _mm512_setzero_si512(), na = _mm512_set1_epi8((char)0xFF);
    const int64_t n64 = n / 64;
    uint64_t bad = 0;
    for (int64_t w = 0; w < n64; w++) {
        const __m512i a = _mm512_loadu_si512(g + 64 * w);
        const uint64_t m1 = _mm512_cmpeq_epi8_mask(a, one), m2 = _mm512_cmpeq_epi8_mask(a, two), m0 = _mm512_cmpeq_epi8_mask(a, zero);
        const uint64_t mm = _mm512_cmpeq_epi8_mask(a, na);
        bad |= ~(m0 | m1 | m2 | mm);
        const uint64_t lo = m1 | mm, hi = m2 | mm;
        uint64_t o0 = _pdep_u64(lo & 0xFFFFFFFFull, 0x5555555555555555ull) | _pdep_u64(hi & 0xFFFFFFFFull, 0xAAAAAAAAAAAAAAAAull);
        uint64_t o1 = _pdep_u64(lo >> 32, 0x5555555555555555ull) | _pdep_u64(hi >> 32, 0xAAAAAAAAAAAAAAAAull);
        memcpy(dst + 16 * w, &o0, 8);
        memcpy(dst + 16 * w + 8, &o1, 8);
    }
    const int64_t done = 64 * n64;
    std::vector<uint8_t> tail(stride - (size_t)(done / 4), 0);
    const bool ok_tail = pack_column_scalar<uint8_t>(g + done, n - done, tail.data(), tail.size());
    memcpy(dst + done / 4, tail.data(), tail.size());
    return bad == 0 && ok_tail;
}

bool cpu_has_avx512()
{
    static const bool has = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
                            __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("bmi2");
    return has;
}
#else
bool cpu_has_avx512() { return false; }
template <typename T> bool pack_column_avx512(const T *, int64_t, uint8_t *, size_t) { return false; }
#endif

template <typename T>
int pack_block(const T *G, int64_t n, int64_t p, uint8_t *packed, size_t stride, int threads)
{
    if (!G || !packed || n <= 0 || p < 0 || stride < (size_t)((n + 3) / 4) || stride % 16 != 0) {
        staar::set_error("pack_dosage: bad arguments");
        return STAAR_ERR_ARG;
    }
    if (threads < 1) {
        // one process per GPU (torchrun sets LOCAL_WORLD_SIZE): the ranks of a box share its cores
        static const int ranks = getenv("LOCAL_WORLD_SIZE") ? std::max(1, atoi(getenv("LOCAL_WORLD_SIZE"))) : 1;
        threads = std::max(1, (int)std::thread::hardware_concurrency() / ranks);
    }
    threads = (int)std::min<int64_t>(threads, std::max<int64_t>(p, 1));
    const bool simd = cpu_has_avx512() && !getenv("STAAR_B200_NO_AVX512");
    std::vector<int> bad(threads, 0);
    auto work = [&](int tid) {
        // contiguous variant ranges per thread: sequential reads of a contiguous slab of the column-major block
        const int64_t i0 = p * tid / threads, i1 = p * (tid + 1) / threads;
        for (int64_t i = i0; i < i1; i++) {
            const T *g = G + i * n;
            uint8_t *dst = packed + (size_t)i * stride;
            const bool ok = simd ? pack_column_avx512(g, n, dst, stride) : pack_column_scalar<T>(g, n, dst, stride);
            if (!ok) bad[tid] = 1;
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; t++) th.emplace_back(work, t);
    work(0);
    for (auto &t : th) t.join();
    for (int b : bad)
        if (b) { staar::set_error("pack_dosage: a non-missing genotype is not 0, 1 or 2"); return STAAR_ERR_ARG; }
    return STAAR_OK;
}

}  // namespace

extern "C" {

int staar_pack_dosage_host(const double *G, int64_t n, int64_t p, uint8_t *packed, size_t stride, int threads)
{
    return pack_block<double>(G, n, p, packed, stride, threads);
}
int staar_pack_dosage_host_i32(const int32_t *G, int64_t n, int64_t p, uint8_t *packed, size_t stride, int threads)
{
    return pack_block<int32_t>(G, n, p, packed, stride, threads);
}
int staar_pack_dosage_host_u8(const uint8_t *G, int64_t n, int64_t p, uint8_t *packed, size_t stride, int threads)
{
    return pack_block<uint8_t>(G, n, p, packed, stride, threads);
}

}  // extern "C"
