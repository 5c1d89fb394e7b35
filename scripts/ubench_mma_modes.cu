// Two questions for the chunked WKV7 kernels (answers in DESIGN.md 2.2b; NOTE: (1) passes here, 50/50 exact, yet the same
// pattern inside wkv7_chunk_bwd_kernel gave wrong gradients — do not rely on it; (2) is used by the product kernels):
//  (1) may two threads accumulate into the SAME TMEM accumulator concurrently (K range split over two issuers)?
//  (2) does tcgen05.mma accept its A operand from TMEM for kind::tf32 (fp32 accumulator columns re-used as A)?
// D[128x64] = A[128x64] * B[64x64]^T with small-integer data (exact in tf32); checked against a CPU product.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../visualrwkv_b200/csrc/umma.cuh"
using namespace vrwkv;

__global__ void k(const float* A, const float* B, float* out, int mode) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t base;
    uint8_t* sa = sm;            // A: K-major SW128, 2 atoms x 128 rows
    uint8_t* sb = sm + 32768;    // B: K-major SW128, 2 atoms x 64 rows
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(&bar, 2); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<256>(&base);
    for (int i = tid; i < 128 * 64; i += blockDim.x) {
        const int r = i >> 6, c = i & 63;
        *reinterpret_cast<float*>(sa + (c >> 5) * 16384 + sw128_off(r, c & 31)) = A[i];
        if (r < 64) *reinterpret_cast<float*>(sb + (c >> 5) * 8192 + sw128_off(r, c & 31)) = B[i];
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t trow = base + ((uint32_t)(32 * warp) << 16);
    {   // zero the accumulator (columns 0-63) and, for mode 2, put A into TMEM columns 64-127 (row = lane, column = k)
        uint32_t z[32];
        for (int e = 0; e < 32; e++) z[e] = 0u;
        tmem_st32(trow, z);
        tmem_st32(trow + 32, z);
        if (mode == 2) {
            uint32_t v[32];
            const int r = 32 * warp + lane;
            for (int h = 0; h < 2; h++) {
                for (int e = 0; e < 32; e++) v[e] = __float_as_uint(A[r * 64 + 32 * h + e]);
                tmem_st32(trow + 64 + 32 * h, v);
            }
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t idesc = umma_idesc_tf32(128, 64);
    if (lane == 0 && warp < 2) {
        if (mode == 0) {          // one issuer, K = 64
            if (warp == 0)
                for (int kk = 0; kk < 8; kk++)
                    umma_tf32(base, umma_desc_advance(umma_desc_sw128(sa + (kk >> 2) * 16384), (kk & 3) * 32),
                              umma_desc_advance(umma_desc_sw128(sb + (kk >> 2) * 8192), (kk & 3) * 32), idesc, 1);
        } else if (mode == 1) {   // two issuers, K halves, same accumulator
            for (int kk = 4 * warp; kk < 4 * warp + 4; kk++)
                umma_tf32(base, umma_desc_advance(umma_desc_sw128(sa + (kk >> 2) * 16384), (kk & 3) * 32),
                          umma_desc_advance(umma_desc_sw128(sb + (kk >> 2) * 8192), (kk & 3) * 32), idesc, 1);
        } else {                  // A from TMEM (columns 64 + 8 kk), B from shared memory
            if (warp == 0)
                for (int kk = 0; kk < 8; kk++) {
                    const uint64_t db = umma_desc_advance(umma_desc_sw128(sb + (kk >> 2) * 8192), (kk & 3) * 32);
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(base), "r"(base + 64 + 8 * kk), "l"(db),
                                 "r"(idesc), "r"(1u)
                                 : "memory");
                }
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    uint32_t v[32];
    for (int h = 0; h < 2; h++) {
        tmem_ld32(trow + 32 * h, v);
        for (int e = 0; e < 32; e++) out[(32 * warp + lane) * 64 + 32 * h + e] = __uint_as_float(v[e]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(base);
}

int main() {
    std::vector<float> A(128 * 64), B(64 * 64), ref(128 * 64), got(128 * 64);
    srand(1);
    for (auto& x : A) x = (float)(rand() % 7 - 3);
    for (auto& x : B) x = (float)(rand() % 5 - 2);
    for (int r = 0; r < 128; r++)
        for (int n = 0; n < 64; n++) {
            float s = 0.f;
            for (int kk = 0; kk < 64; kk++) s += A[r * 64 + kk] * B[n * 64 + kk];
            ref[r * 64 + n] = s;
        }
    float *dA, *dB, *dO;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, got.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const char* names[3] = {"one issuer", "two issuers, same accumulator", "A operand from TMEM"};
    for (int mode = 0; mode < 3; mode++) {
        int bad_runs = 0; double worst = 0;
        for (int rep = 0; rep < 50; rep++) {
            k<<<1, 128, 64 * 1024>>>(dA, dB, dO, mode);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 0; }
            cudaMemcpy(got.data(), dO, got.size() * 4, cudaMemcpyDeviceToHost);
            double m = 0;
            for (size_t i = 0; i < got.size(); i++) m = std::max(m, (double)fabsf(got[i] - ref[i]));
            if (m > 0) bad_runs++;
            worst = std::max(worst, m);
        }
        printf("mode %d (%s): %d of 50 runs differ from the exact product, max |err| %.1f\n", mode, names[mode], bad_runs, worst);
    }
    return 0;
}
