// TMEM -> register read throughput: W warps each issue REP x tcgen05.ld.32x32b.x32 (4 KB per instruction and warp).
#include <cstdio>
#include <cuda_runtime.h>
#include "../visualrwkv_b200/csrc/umma.cuh"
using namespace vrwkv;
__global__ void k(int rep, int mode, long long* out, float* sink) {
    __shared__ uint32_t base;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc<512>(&base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t t = base + ((uint32_t)(32 * (warp & 3)) << 16);
    float acc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < rep; i++) {
        if (mode == 0) {
            uint32_t v[32];
            tmem_ld32_nowait(t + 32 * ((i + (warp >> 2)) & 15), v);
            tmem_ld_wait();
            acc += __uint_as_float(v[0]) + __uint_as_float(v[31]);
        } else if (mode == 1) {
            uint32_t v[16];
            tmem_ld16_nowait(t + 16 * ((i + (warp >> 2)) & 31), v);
            tmem_ld_wait();
            acc += __uint_as_float(v[0]) + __uint_as_float(v[15]);
        } else {  // four x16 loads in flight before the wait
            uint32_t a[16], b[16], c[16], d[16];
            tmem_ld16_nowait(t + 64 * (i & 7), a);
            tmem_ld16_nowait(t + 64 * (i & 7) + 16, b);
            tmem_ld16_nowait(t + 64 * (i & 7) + 32, c);
            tmem_ld16_nowait(t + 64 * (i & 7) + 48, d);
            tmem_ld_wait();
            acc += __uint_as_float(a[0]) + __uint_as_float(b[1]) + __uint_as_float(c[2]) + __uint_as_float(d[3]);
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (acc == 123.456f) sink[0] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(base);
}
int main() {
    long long* d; float* s;
    cudaMalloc(&d, 8); cudaMalloc(&s, 4);
    for (int mode = 0; mode < 3; mode++)
        for (int warps : {1, 4, 8, 16}) {
            const int rep = 256;
            k<<<1, warps * 32>>>(rep, mode, d, s);
            k<<<1, warps * 32>>>(rep, mode, d, s);
            cudaDeviceSynchronize();
            long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
            const double bytes = (double)warps * rep * 32 * 4 * (mode == 0 ? 32 : mode == 1 ? 16 : 64);
            printf("mode %d (%s) warps %2d: %lld cycles, %.1f B/clk per SM, %.1f cycles per instruction-group\n", mode,
                   mode == 0 ? "x32" : mode == 1 ? "x16" : "4 x x16", warps, c, bytes / c, (double)c / rep);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
