// synth.cu — synthetic packed dosage columns generated on the device (bench.py / parity subsets).
// Same counter-based hash and integer thresholds as staarpipeline_b200/synth.py (hash_u32, dosage_codes), so a
// device-resident batch can be regenerated bit-for-bit on the host for the oracle.  Not on the measured path.
#include "common.cuh"

namespace staar {

namespace {

__device__ __forceinline__ uint32_t hash_u32(uint64_t seed, uint64_t variant, uint64_t sample, uint64_t stream)
{
    uint64_t z = seed + stream * 0xD6E8FEB86659FD93ull + variant * 0x9E3779B97F4A7C15ull + sample * 0xC2B2AE3D27D4EB4Full;
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)(z >> 32);
}

__global__ void synth_packed_kernel(uint8_t *__restrict__ packed, size_t stride, int64_t n, int64_t first_variant,
                                    int64_t p, uint64_t seed, const uint64_t *__restrict__ thr_hom,
                                    const uint64_t *__restrict__ thr_het, const uint64_t *__restrict__ thr_miss,
                                    const uint8_t *__restrict__ store_alt, const int32_t *__restrict__ order)
{
    const int64_t words_per_col = (int64_t)(stride >> 2);
    const int64_t total = p * words_per_col;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = idx / words_per_col, wi = idx - i * words_per_col;
        const uint64_t v = (uint64_t)(first_variant + i);
        const uint64_t th = thr_hom[i], te = thr_het[i], tm = thr_miss[i];
        const bool alt_stored = store_alt[i] != 0;
        uint32_t w = 0;
        for (int f = 0; f < 16; f++) {
            const int64_t pos = wi * 16 + f;
            if (pos >= n) break;
            const int64_t j = order ? (int64_t)order[pos] : pos;     // the sample stored at this position
            const uint64_t u = hash_u32(seed, v, (uint64_t)j, 0), u2 = hash_u32(seed, v, (uint64_t)j, 4);
            uint32_t alt = (u < th ? 1u : 0u) + (u < te ? 1u : 0u);
            uint32_t code = alt_stored ? alt : 2u - alt;
            if (u2 < tm) code = 3u;
            w |= code << (2 * f);
        }
        reinterpret_cast<uint32_t *>(packed + (size_t)i * stride)[wi] = w;
    }
}

}  // namespace

int launch_synth_packed(staar_ctx *ctx, uint8_t *d_packed, size_t stride, int64_t n, int64_t first_variant, int64_t p,
                        uint64_t seed, const uint64_t *thr_hom, const uint64_t *thr_het, const uint64_t *thr_miss,
                        const uint8_t *store_alt, const int32_t *sample_order)
{
    if (p == 0) return STAAR_OK;
    const int64_t total = p * (int64_t)(stride >> 2);
    int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)ctx->sm_count * 32);
    synth_packed_kernel<<<blocks, 256, 0, ctx->stream>>>(d_packed, stride, n, first_variant, p, seed, thr_hom, thr_het,
                                                         thr_miss, store_alt, sample_order);
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

}  // namespace staar
