// tcgen05.mma issue / completion cost: one thread (or one lane of each of W warps) issues REP MMAs of shape
// M=128 x N x (K=8 tf32 | K=16 bf16) from shared-memory descriptors, then commits and waits.
#include <cstdio>
#include <cuda_runtime.h>
#include "../visualrwkv_b200/csrc/umma.cuh"
using namespace vrwkv;
__global__ void k(int rep, int N, int bf16, int nissuer, long long* out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar, nissuer); fence_mbar_init(); }
    for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 0.f;
    if (warp == 0) tmem_alloc<512>(&base);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t idesc = bf16 ? umma_idesc_bf16(128, N) : umma_idesc_tf32(128, N);
    long long t0 = 0, t1 = 0;
    if (lane == 0 && warp < nissuer) {
        const uint64_t da = umma_desc_sw128(sm), db = umma_desc_sw128(sm + 32768);
        const uint32_t d = base + (uint32_t)(warp * (N <= 128 ? 128 : 256));
        t0 = clock64();
        for (int i = 0; i < rep; i++) {
            if (bf16) umma_bf16(d, umma_desc_advance(da, (i & 3) * 32), umma_desc_advance(db, (i & 3) * 32), idesc, i > 0);
            else umma_tf32(d, umma_desc_advance(da, (i & 3) * 32), umma_desc_advance(db, (i & 3) * 32), idesc, i > 0);
        }
        t1 = clock64();
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(base);
}
int main() {
    long long* d;
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int rep = 64;
    for (int bf16 = 0; bf16 < 2; bf16++)
        for (int N : {64, 128, 256})
            for (int ni : {1, 2, 4}) {
                if (ni * (N <= 128 ? 128 : 256) > 512) continue;
                k<<<1, 128, 100 * 1024>>>(rep, N, bf16, ni, d);
                k<<<1, 128, 100 * 1024>>>(rep, N, bf16, ni, d);
                cudaDeviceSynchronize();
                long long c[2];
                cudaMemcpy(c, d, 16, cudaMemcpyDeviceToHost);
                printf("%s N=%3d issuers=%d: issue %.1f cyc/MMA, issue+complete %.1f cyc/MMA (per issuer)\n", bf16 ? "bf16 K16" : "tf32 K8 ", N, ni,
                       (double)c[0] / rep, (double)c[1] / rep);
            }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
