// Latency microbenchmarks on sm_100a: dependent chains of SHFL, FFMA2, FFMA, FADD, LDS, SEL; 1 warp.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int MODE> __global__ void k(float* out, long long* cyc, int iters) {
    __shared__ float sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (float)((i * 7) & 1023);
    __syncthreads();
    float x = threadIdx.x * 0.001f + 1.0f, y = 0.999f;
    u64 p; asm("mov.b64 %0, {%1,%2};" : "=l"(p) : "f"(x), "f"(y));
    u64 m; asm("mov.b64 %0, {%1,%2};" : "=l"(m) : "f"(y), "f"(y));
    int idx = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (MODE == 0) x = __shfl_xor_sync(0xffffffffu, x, 1 + (u & 3));
            if (MODE == 1) p = ffma2(p, m, m);
            if (MODE == 2) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x) : "f"(y));
            if (MODE == 3) asm volatile("add.f32 %0, %0, %1;" : "+f"(x) : "f"(y));
            if (MODE == 4) { idx = (int)sm[idx & 1023]; }
            if (MODE == 5) { x = __shfl_xor_sync(0xffffffffu, x, 1 + (u & 3)) + y; }
            if (MODE == 6) { float s = (threadIdx.x & 4) ? x : y; x = __shfl_xor_sync(0xffffffffu, s, 4) + ((threadIdx.x & 4) ? y : x); }
        }
    }
    long long t1 = clock64();
    float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p));
    out[threadIdx.x] = x + lo + hi + idx;
    if (threadIdx.x == 0) cyc[MODE] = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 4096); cudaMallocManaged(&cyc, 64 * 8);
    const char* names[] = {"SHFL chain", "FFMA2 chain", "FFMA chain", "FADD chain", "LDS chain (+cvt)", "SHFL+FADD chain", "SEL+SHFL+SEL+FADD"};
    int iters = 2000;
    for (int rep = 0; rep < 2; rep++) {
        k<0><<<1, 32>>>(out, cyc, iters); k<1><<<1, 32>>>(out, cyc, iters); k<2><<<1, 32>>>(out, cyc, iters);
        k<3><<<1, 32>>>(out, cyc, iters); k<4><<<1, 32>>>(out, cyc, iters); k<5><<<1, 32>>>(out, cyc, iters); k<6><<<1, 32>>>(out, cyc, iters);
        cudaDeviceSynchronize();
    }
    for (int m = 0; m < 7; m++) printf("%-22s %.1f cycles/op\n", names[m], (double)cyc[m] / (iters * 16.0));
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
