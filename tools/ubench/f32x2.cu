// Micro-benchmark: issue cost of packed fp32 (FFMA2) against scalar FFMA on sm_100a.
// Each thread runs NACC independent FMA chains; both kernels do the same number of fp32 FMAs.
#include <cuda_runtime.h>
#include <cstdio>
struct f2 { unsigned long long v; };
__device__ __forceinline__ f2 mk(float a, float b) { f2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
constexpr int NACC = 8, ITERS = 4096;
__global__ void k_scalar(float* out, float a, float b) {
    float acc[2 * NACC];
    for (int i = 0; i < 2 * NACC; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < 2 * NACC; ++i) acc[i] = fmaf(acc[i], a, b);
    float s = 0; for (int i = 0; i < 2 * NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float* out, float a, float b) {
    f2 acc[NACC];
    for (int i = 0; i < NACC; ++i) acc[i] = mk(threadIdx.x * 1e-3f + i, threadIdx.x * 1e-3f + i + NACC);
    const f2 A = mk(a, a), B = mk(b, b);
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fma2(acc[i], A, B);
    float s = 0;
    for (int i = 0; i < NACC; ++i) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i].v)); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps = 4; warps <= 32; warps *= 2) {
        const int threads = 32 * warps > 1024 ? 1024 : 32 * warps, blocks = 148 * (32 * warps / threads);
        float ms[2];
        for (int v = 0; v < 2; ++v) {
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                if (v == 0) k_scalar<<<blocks, threads>>>(out, 1.0001f, 0.5f); else k_packed<<<blocks, threads>>>(out, 1.0001f, 0.5f);
                cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms[v], e0, e1);
            }
        }
        const double fma = (double)blocks * threads * ITERS * 2 * NACC;
        printf("warps/SM %2d: scalar %.3f ms (%.1f TFLOP/s), packed %.3f ms (%.1f TFLOP/s)\n", warps, ms[0], 2 * fma / ms[0] / 1e9,
               ms[1], 2 * fma / ms[1] / 1e9);
    }
    return 0;
}
