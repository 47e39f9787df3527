// tools/pipe_peaks.cu — measures the pipe peaks this path's rooflines are quoted against
// (MEASURED_PEAKS.json has HBM and bf16 only): FP64 FMA, FP64 mma.sync (DMMA), INT8 mma.sync (IMMA), and the
// tcgen05.mma.kind::i8 rate at the contraction's instruction shape (M = 128, K = 32, A in TMEM, N = 112).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_peaks tools/pipe_peaks.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;

__global__ void k_dfma(double *out, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < ITERS; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void k_dmma884(double *out, double a, double b) {
    double c[4][2] = {};
    double av = a + threadIdx.x, bv = b;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int t = 0; t < 4; t++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[t][0]), "+d"(c[t][1]) : "d"(av), "d"(bv));
    }
    double s = 0; for (int t = 0; t < 4; t++) s += c[t][0] + c[t][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma16816(double *out, double a, double b) {
    double c[4][4] = {};
    double av[8], bv[4];
    for (int i = 0; i < 8; i++) av[i] = a + i + threadIdx.x;
    for (int i = 0; i < 4; i++) bv[i] = b + i;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int t = 0; t < 4; t++)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[t][0]), "+d"(c[t][1]), "+d"(c[t][2]), "+d"(c[t][3])
                         : "d"(av[0]), "d"(av[1]), "d"(av[2]), "d"(av[3]), "d"(av[4]), "d"(av[5]), "d"(av[6]), "d"(av[7]),
                           "d"(bv[0]), "d"(bv[1]), "d"(bv[2]), "d"(bv[3]));
    }
    double s = 0; for (int t = 0; t < 4; t++) for (int j = 0; j < 4; j++) s += c[t][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_imma16832(int *out, int a, int b) {
    int c[8][4] = {};
    int av[4] = {a + (int)threadIdx.x, a + 1, a + 2, a + 3}, bv[2] = {b, b + 1};
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int t = 0; t < 8; t++)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(c[t][0]), "+r"(c[t][1]), "+r"(c[t][2]), "+r"(c[t][3])
                         : "r"(av[0]), "r"(av[1]), "r"(av[2]), "r"(av[3]), "r"(bv[0]), "r"(bv[1]));
    }
    int s = 0; for (int t = 0; t < 8; t++) for (int j = 0; j < 4; j++) s += c[t][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_hmma16816(float *out, unsigned a, unsigned b) {
    float c[8][4] = {};
    unsigned av[4] = {a + threadIdx.x, a + 1, a + 2, a + 3}, bv[2] = {b, b + 1};
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int t = 0; t < 8; t++)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[t][0]), "+f"(c[t][1]), "+f"(c[t][2]), "+f"(c[t][3])
                         : "r"(av[0]), "r"(av[1]), "r"(av[2]), "r"(av[3]), "r"(bv[0]), "r"(bv[1]));
    }
    float s = 0; for (int t = 0; t < 8; t++) for (int j = 0; j < 4; j++) s += c[t][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- tcgen05.mma.cta_group::1.kind::i8: one CTA per SM, one thread issues `reps` x 16 back-to-back MMAs (M = 128,
// K = 32) against the same operands — A in TMEM (".ts", as the contraction uses it) or in shared memory (".ss"),
// B in shared memory (K-major, no swizzle; contents do not matter) — then one commit and a wait.  No operand streams:
// this is the tensor pipe's own rate for the instruction shape, the denominator of the contraction's roofline.
__device__ __forceinline__ uint32_t saddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t sdesc(uint32_t addr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
template <int N, bool A_TMEM, int COMMITS = 0, int DELAY = 0>
__global__ void __launch_bounds__(128, 1) k_tcgen05_i8(int reps, int *out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    uint32_t *slot = reinterpret_cast<uint32_t *>(smem + 8);
    uint8_t *sA = smem + 1024, *sB = sA + 128 * 32;              // A: 128 rows x 32 B, B: N rows x 32 B
    for (int i = threadIdx.x; i < (128 + N) * 32 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(sA)[i] = 0x01010101u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(saddr(bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(saddr(bar + 2)));   // intermediate commits (COMMITS > 0): nobody waits
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(saddr(bar + 3)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(saddr(slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t bdesc = sdesc(saddr(sB), (N / 8) * 128, 128), adesc = sdesc(saddr(sA), 16 * 128, 128);
        for (int r = 0; r < reps; r++) {
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if (A_TMEM)
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                                 "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n}\n"
                                 ::"r"(tmem), "r"(tmem + 256 + 8 * (k & 7)), "l"(bdesc), "r"(idesc), "r"(1u), "r"(0u) : "memory");
                else
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n"
                                 ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u), "r"(0u) : "memory");
            }
            // the contraction commits twice per 16 instructions (A stage and B stage released separately)
            if (COMMITS >= 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(saddr(bar + 2)) : "memory");
            if (COMMITS >= 2) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(saddr(bar + 3)) : "memory");
            if (DELAY > 0) {           // the issuing thread is away for DELAY clocks (barrier waits between k-blocks): does the queue cover it?
                const long long t0 = clock64();
                while (clock64() - t0 < DELAY) { }
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(saddr(bar)) : "memory");
        asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n"
                     ::"r"(saddr(bar)) : "memory");
        out[blockIdx.x] = reps;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

template <int N, bool A_TMEM, int COMMITS = 0, int DELAY = 0> double tcgen05_tops(int sms, int *out)
{
    const int reps = 2048;
    const size_t smem = 1024 + (128 + N) * 32 + 1024;
    cudaFuncSetAttribute(k_tcgen05_i8<N, A_TMEM, COMMITS, DELAY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_tcgen05_i8<N, A_TMEM, COMMITS, DELAY><<<sms, 128, smem>>>(64, out); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); k_tcgen05_i8<N, A_TMEM, COMMITS, DELAY><<<sms, 128, smem>>>(reps, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return (double)sms * reps * 16 * (2.0 * 128 * N * 32) / best / 1e9;     // TOP/s
}

template <typename F> float time_ms(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);
    void *buf; CK(cudaMalloc(&buf, (size_t)sms * 16 * 1024 * 8));
    const int blocks = sms * 8, threads = 512;
    double warps = (double)blocks * threads / 32;
    float ms;
    ms = time_ms([&] { k_dfma<<<blocks, threads>>>((double *)buf, 1.0000001, 1e-9); });
    printf(", \"fp64_fma_tflops\": %.2f", (double)blocks * threads * ITERS * 8 * 2 / ms / 1e9);
    ms = time_ms([&] { k_dmma884<<<blocks, threads>>>((double *)buf, 1.0, 1e-9); });
    printf(", \"fp64_mma_m8n8k4_tflops\": %.2f", warps * ITERS * 4 * (8.0 * 8 * 4 * 2) / ms / 1e9);
    ms = time_ms([&] { k_dmma16816<<<blocks, threads>>>((double *)buf, 1.0, 1e-9); });
    printf(", \"fp64_mma_m16n8k16_tflops\": %.2f", warps * ITERS * 4 * (16.0 * 8 * 16 * 2) / ms / 1e9);
    ms = time_ms([&] { k_imma16832<<<blocks, threads>>>((int *)buf, 1, 2); });
    printf(", \"int8_mma_m16n8k32_tops\": %.2f", warps * ITERS * 8 * (16.0 * 8 * 32 * 2) / ms / 1e9);
    ms = time_ms([&] { k_hmma16816<<<blocks, threads>>>((float *)buf, 1, 2); });
    printf(", \"bf16_mma_m16n8k16_tflops\": %.2f", warps * ITERS * 8 * (16.0 * 8 * 16 * 2) / ms / 1e9);
    CK(cudaGetLastError());
    // tcgen05 kind::i8, M = 128, K = 32: A from TMEM (.ts) at N = 112 (the contraction's shape), 128, 256; A from smem at 112, 256
    printf(", \"tcgen05_i8_ts_n112_tops\": %.1f", tcgen05_tops<112, true>(sms, (int *)buf));
    printf(", \"tcgen05_i8_ts_n112_commit1_tops\": %.1f", tcgen05_tops<112, true, 1>(sms, (int *)buf));
    printf(", \"tcgen05_i8_ts_n112_commit2_tops\": %.1f", tcgen05_tops<112, true, 2>(sms, (int *)buf));
    printf(", \"tcgen05_i8_ts_n112_commit2_delay100_tops\": %.1f", tcgen05_tops<112, true, 2, 100>(sms, (int *)buf));
    printf(", \"tcgen05_i8_ts_n112_commit2_delay200_tops\": %.1f", tcgen05_tops<112, true, 2, 200>(sms, (int *)buf));
    printf(", \"tcgen05_i8_ts_n112_commit2_delay400_tops\": %.1f", tcgen05_tops<112, true, 2, 400>(sms, (int *)buf));
    printf(", \"tcgen05_i8_ts_n112_commit2_delay800_tops\": %.1f", tcgen05_tops<112, true, 2, 800>(sms, (int *)buf));
    printf(", \"tcgen05_i8_ts_n128_tops\": %.1f", tcgen05_tops<128, true>(sms, (int *)buf));
    printf(", \"tcgen05_i8_ts_n256_tops\": %.1f", tcgen05_tops<256, true>(sms, (int *)buf));
    printf(", \"tcgen05_i8_ts_n64_tops\": %.1f", tcgen05_tops<64, true>(sms, (int *)buf));
    printf(", \"tcgen05_i8_ss_n112_tops\": %.1f", tcgen05_tops<112, false>(sms, (int *)buf));
    printf(", \"tcgen05_i8_ss_n256_tops\": %.1f", tcgen05_tops<256, false>(sms, (int *)buf));
    CK(cudaGetLastError());
    printf("}\n");
    return 0;
}
