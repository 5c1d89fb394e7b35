// Round-2 micro-tests for the bf16 "x6" chunk kernels (answers recorded in DESIGN.md):
//  (1) one SWIZZLE_128B bf16 tile read K-major and MN-major, for the A and the B operand (integer data, exact);
//  (2) A operand from TMEM for kind::f16 (two bf16 per 32-bit column);
//  (3) M=64 instruction: which accumulator lanes it writes;
//  (4) accuracy of the 3-way bf16 split with 6 product terms on fp32 data;
//  (5) issue/complete cost of 24 dependent bf16 MMAs (N=64 / 128; K-major / MN-major B).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../visualrwkv_b200/csrc/umma.cuh"
using namespace vrwkv;

// mode bits: 1 = A MN-major, 2 = B MN-major, 4 = A from TMEM, 8 = M=64 instruction
// A is given as fp32 [128][64] (m, k), B as fp32 [64][64] (n, k); nsplit = 1 (values are bf16-exact) or 3
__global__ void k(const float* A, const float* B, float* out, int mode, int nsplit, long long* clk) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t base;
    // A: 3 components x 2 tiles (rows m 0-63, 64-127 for K-major; M blocks for MN-major); B: 3 components x 1 tile
    uint8_t* sa = sm;
    uint8_t* sb = sm + 3 * 2 * BT_BYTES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool a_mn = mode & 1, b_mn = mode & 2, a_tm = mode & 4, m64 = mode & 8;
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<256>(&base);
    for (int i = tid; i < 128 * 64; i += blockDim.x) {
        const int m = i >> 6, kk = i & 63;
        uint16_t s[3];
        split3(A[i], s[0], s[1], s[2]);
        for (int c = 0; c < 3; c++) {
            uint8_t* t = sa + c * 2 * BT_BYTES + (m >> 6) * BT_BYTES;
            const uint32_t off = a_mn ? bt_off(kk, m & 63) : bt_off(m & 63, kk);
            *reinterpret_cast<uint16_t*>(t + off) = s[c];
        }
        if (m < 64) {
            split3(B[i], s[0], s[1], s[2]);
            for (int c = 0; c < 3; c++)
                *reinterpret_cast<uint16_t*>(sb + c * BT_BYTES + (b_mn ? bt_off(kk, m) : bt_off(m, kk))) = s[c];
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t trow = base + ((uint32_t)(32 * warp) << 16);
    {
        uint32_t z[32];
        for (int e = 0; e < 32; e++) z[e] = 0x7fc00000u;  // NaN pattern: untouched lanes stay visible
        tmem_st32(trow, z);
        tmem_st32(trow + 32, z);
        if (a_tm) {  // A components into TMEM columns 64 + 32 c .. : column j holds k = 2j, 2j+1
            const int r = 32 * warp + lane;
            for (int c = 0; c < 3; c++) {
                uint32_t v[32];
                for (int e = 0; e < 32; e++) {
                    uint16_t s0[3], s1[3];
                    split3(A[r * 64 + 2 * e], s0[0], s0[1], s0[2]);
                    split3(A[r * 64 + 2 * e + 1], s1[0], s1[1], s1[2]);
                    v[e] = (uint32_t)s0[c] | ((uint32_t)s1[c] << 16);
                }
                tmem_st32(trow + 64 + 32 * c, v);
            }
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t b4 = smem_u32(sm) >> 4;
    const uint32_t idesc = umma_idesc_bf16_mj(m64 ? 64 : 128, 64, a_mn, b_mn);
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        t0 = clock64();
        bool first = true;
        for (int ca = 0; ca < nsplit; ca++)
            for (int cb = 0; cb + ca < nsplit; cb++)
                for (int ks = 0; ks < 4; ks++) {
                    const uint64_t db = b_mn ? bdesc_mn(b4, 3 * 2 * BT_BYTES + cb * BT_BYTES + ks * 2048)
                                             : bdesc_k(b4, 3 * 2 * BT_BYTES + cb * BT_BYTES + ks * 32);
                    if (a_tm) {
                        umma_bf16_ts(base, base + 64 + 32 * ca + 8 * ks, db, idesc, first ? 0u : 1u);
                    } else {
                        const uint64_t da = a_mn ? bdesc_mn(b4, ca * 2 * BT_BYTES + ks * 2048) : bdesc_k(b4, ca * 2 * BT_BYTES + ks * 32);
                        umma_bf16(base, da, db, idesc, first ? 0u : 1u);
                    }
                    first = false;
                }
        t1 = clock64();
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (tid == 0 && clk) { clk[0] = t1 - t0; clk[1] = t2 - t0; }
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
    std::vector<float> A(128 * 64), B(64 * 64), got(128 * 64);
    std::vector<double> ref(128 * 64);
    float *dA, *dB, *dO; long long* dC;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, got.size() * 4); cudaMalloc(&dC, 16);
    const int SM = 3 * 2 * BT_BYTES + 3 * BT_BYTES + 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SM);
    auto run = [&](int mode, int nsplit, bool integer) {
        srand(7 + mode);
        for (auto& x : A) x = integer ? (float)(rand() % 7 - 3) : (float)((rand() / (double)RAND_MAX * 2 - 1) * exp((rand() % 9) - 4.0));
        for (auto& x : B) x = integer ? (float)(rand() % 5 - 2) : (float)((rand() / (double)RAND_MAX * 2 - 1) * exp((rand() % 9) - 4.0));
        for (int r = 0; r < 128; r++)
            for (int n = 0; n < 64; n++) {
                double s = 0;
                for (int kk = 0; kk < 64; kk++) s += (double)A[r * 64 + kk] * (double)B[n * 64 + kk];
                ref[r * 64 + n] = s;
            }
        cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
        k<<<1, 128, SM>>>(dA, dB, dO, mode, nsplit, dC);
        k<<<1, 128, SM>>>(dA, dB, dO, mode, nsplit, dC);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); exit(0); }
        cudaMemcpy(got.data(), dO, got.size() * 4, cudaMemcpyDeviceToHost);
        long long c[2];
        cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost);
        double maxabs = 0, num = 0, den = 0; int nan_rows = 0, bad = 0;
        const int rows = 128;
        for (int r = 0; r < rows; r++) {
            bool rownan = false;
            for (int n = 0; n < 64; n++) {
                const float g = got[r * 64 + n];
                if (g != g) { rownan = true; continue; }
                const double d = fabs(g - ref[r * 64 + n]);
                maxabs = std::max(maxabs, d); num += d * d; den += ref[r * 64 + n] * ref[r * 64 + n];
                if (d > 0) bad++;
            }
            nan_rows += rownan;
        }
        printf("mode %2d (A %s%s, B %s%s) nsplit %d %s: untouched rows %d, mismatches %d, max|err| %.3g, rms rel %.3g; %d MMAs issue %.0f cyc, issue+complete %.0f cyc\n",
               mode, (mode & 4) ? "TMEM" : ((mode & 1) ? "MN" : "K"), "", (mode & 2) ? "MN" : "K", (mode & 8) ? ", M=64" : "", nsplit,
               integer ? "int" : "fp32", nan_rows, bad, maxabs, sqrt(num / std::max(den, 1e-300)), nsplit == 1 ? 4 : 24, (double)c[0], (double)c[1]);
        if (mode & 8) {
            printf("   rows written by M=64:");
            for (int r = 0; r < 128; r++) if (got[r * 64] == got[r * 64]) printf(" %d", r);
            printf("\n");
        }
    };
    for (int mode : {0, 1, 2, 3, 4, 6, 8, 9}) run(mode, 1, true);
    for (int mode : {0, 3, 4}) run(mode, 3, false);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
