// tools/pipe_peaks.cu — measures the pipe peaks this path's rooflines are quoted against
// (MEASURED_PEAKS.json has HBM and bf16 only): FP64 FMA, FP64 mma.sync (DMMA), INT8 mma.sync (IMMA).
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
    printf("}\n");
    return 0;
}
