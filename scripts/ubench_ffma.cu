// Microbenchmark: FFMA (3-reg) vs FFMA2 (packed fp32x2) issue throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int MODE> __global__ void k(float* out, int iters, float x) {
    float a[16]; u64 p[8];
    for (int i = 0; i < 16; i++) a[i] = x + i + threadIdx.x;
    for (int i = 0; i < 8; i++) asm("mov.b64 %0, {%1,%2};" : "=l"(p[i]) : "f"(a[2*i]), "f"(a[2*i+1]));
    u64 m; asm("mov.b64 %0, {%1,%2};" : "=l"(m) : "f"(x), "f"(x));
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], x, a[(i + 1) & 15] * 0.0f + x);
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = ffma2(p[i], m, m);
        }
    }
    float s = 0;
    if (MODE == 0) { for (int i = 0; i < 16; i++) s += a[i]; }
    else { for (int i = 0; i < 8; i++) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i])); s += lo + hi; } }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma(float* out, int iters, float x) {
    float a[16];
    for (int i = 0; i < 16; i++) a[i] = x + i + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(x));
    }
    float s = 0; for (int i = 0; i < 16; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000; float ms;
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0); k_ffma<<<148 * 4, 512>>>(out, iters, 1.0001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("FFMA : %.3f ms  %.1f TFLOP/s\n", ms, 2.0 * 16 * iters * 148 * 4 * 512 / ms / 1e9);
        cudaEventRecord(e0); k<1><<<148 * 4, 512>>>(out, iters, 1.0001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("FFMA2: %.3f ms  %.1f TFLOP/s\n", ms, 2.0 * 16 * iters * 148 * 4 * 512 / ms / 1e9);
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
