// packed.cuh — device helpers for 2-bit packed dosage columns.
// Layout (include/staar_b200.h): variant-major, sample j in byte j/4, bits 2*(j%4)..+1,
// codes 0/1/2 = dosage, 3 = missing, tail bits beyond sample n-1 are zero, stride multiple of 16 B.
#pragma once
#include "common.cuh"

namespace staar {

struct PackedCounts { int c1, c2, cm; };          // #het, #hom(2), #missing  (c0 = n - c1 - c2 - cm)
struct VariantFlip { double af, maf, mu; int flip; };

#ifdef __CUDACC__
__device__ __forceinline__ uint4 ldg_nc_u4(const void *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ void count_word(uint32_t w, int &c1, int &c2, int &cm)
{
    const uint32_t lo = w & 0x55555555u, hi = (w >> 1) & 0x55555555u;
    c1 += __popc(lo & ~hi);
    c2 += __popc(hi & ~lo);
    cm += __popc(lo & hi);
}

// Warp-cooperative counts over one packed column (all 32 lanes call; all lanes get the totals).
__device__ __forceinline__ PackedCounts count_packed_column(const uint8_t *col, size_t stride, int lane)
{
    int c1 = 0, c2 = 0, cm = 0;
    const uint4 *v = reinterpret_cast<const uint4 *>(col);
    const int nvec = (int)(stride >> 4);
    for (int k = lane; k < nvec; k += 32) {
        uint4 w = __ldg(v + k);
        count_word(w.x, c1, c2, cm); count_word(w.y, c1, c2, cm);
        count_word(w.z, c1, c2, cm); count_word(w.w, c1, c2, cm);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c1 += __shfl_xor_sync(0xffffffffu, c1, o);
        c2 += __shfl_xor_sync(0xffffffffu, c2, o);
        cm += __shfl_xor_sync(0xffffffffu, cm, o);
    }
    return PackedCounts{c1, c2, cm};
}

// AF / MAF / flip / imputation value from integer counts — the arithmetic of
// matrix_flip_mean.cpp:36-41,56-70 and matrix_flip_minor.cpp:36-41,44-95 on exact integer sums.
//   mean : AF = sum/2/num (0 if all missing); missing <- 2*AF; flip iff AF > 0.5
//   minor: missing <- 0 (AF<=0.5) or 2; AF recomputed over all n; flip iff that AF > 0.5
// mu = value a missing entry holds AFTER the flip (what the score test sees).
__device__ __forceinline__ VariantFlip flip_from_counts(PackedCounts pc, int64_t n, int impute)
{
    const double sum = (double)pc.c1 + 2.0 * (double)pc.c2;
    const double num = (double)(n - pc.cm);
    double af = sum;
    if (num > 0.0) af = sum / 2.0 / num;
    double fill;
    if (impute == STAAR_IMPUTE_MEAN) {
        fill = af * 2.0;
    } else {
        fill = (af <= 0.5) ? 0.0 : 2.0;
        af = (sum + fill * (double)pc.cm) / 2.0 / (double)n;
    }
    VariantFlip vf;
    vf.flip = !(af <= 0.5);
    vf.af = af;
    vf.maf = vf.flip ? 1.0 - af : af;
    vf.mu = vf.flip ? 2.0 - fill : fill;
    return vf;
}

// 2-bit flip g -> 2-g on a whole word: 00<->10, 01 and 11 (missing) unchanged.
__device__ __forceinline__ uint32_t flip_word(uint32_t w)
{
    const uint32_t lo = w & 0x55555555u;
    return w ^ ((~lo & 0x55555555u) << 1);
}
#endif

}  // namespace staar
