// gemm_sm100.cu — bf16 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands staged
// by TMA into 128B-swizzled shared memory), with the RWKV block's element-wise neighbours fused into the epilogue.
//
//   C[M,N] = epilogue( A[M,K] . B[N,K]^T )        A, B, C bf16 row-major (torch.nn.Linear layout), fp32 accumulate
//
//   epilogue NONE     : C = acc                                   receptance / key / value projections (model.py:175-178)
//   epilogue RELU_SQ  : C = relu(acc)^2                           channel-mix key + activation          (model.py:225)
//   epilogue ADD      : C = R + acc  (R [M,N] bf16)               output / value projections + residual (model.py:194,227,251-252)
//
// Structure (persistent: one CTA per SM loops over 128 x BN output tiles, n fastest so that the CTAs running
// concurrently share A tiles in L2; 192 threads; the fp32 accumulator is double-buffered in TMEM so that the epilogue
// of tile i overlaps the TMA/MMA main loop of tile i+1):
//   warp 0   : TMA producer — one elected lane issues cp.async.bulk.tensor.2d for the A (128x64) and B (BNx64) k-blocks
//              into a NSTAGE ring; completion via mbarrier complete_tx;
//   warp 1   : allocates BN TMEM columns, then one elected lane issues 4 x tcgen05.mma.cta_group::1.kind::f16
//              (M=128, N=BN, K=16) per k-block and releases the stage with tcgen05.commit -> mbarrier;
//   warps 2-5: epilogue — each warp owns the 32 TMEM lanes (= output rows) of its quadrant (warp id mod 4), reads the
//              accumulator with tcgen05.ld.32x32b.x32, applies the epilogue in fp32, converts to bf16 and stores
//              64 contiguous bytes per row and 32-column chunk.
// Shared-memory operand layout: K-major, 128-byte swizzle (CU_TENSOR_MAP_SWIZZLE_128B <-> UMMA LayoutType SWIZZLE_128B,
// descriptor SBO = 1024 B, k-advance of 32 B inside the swizzle atom).
#include <cudaTypedefs.h>

#include "common.cuh"
#include "host_util.h"
#include "umma.cuh"

namespace vrwkv {

constexpr int GM_BM = 128, GM_BK = 64, GM_STAGES = 4;

__device__ __forceinline__ bool elect_one_lane() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
enum { GM_EPI_NONE = 0, GM_EPI_RELU_SQ = 1, GM_EPI_ADD = 2 };

struct GemmArgs {
    int M, N, K;
    uint16_t* C;
    const uint16_t* R;  // residual (EPI_ADD)
};

template <int BN>
struct alignas(1024) GemmSmem {
    uint16_t a[GM_STAGES][GM_BM * GM_BK];  // 16 KB per stage
    uint16_t b[GM_STAGES][BN * GM_BK];     // BN*128 B per stage
    uint64_t full[GM_STAGES], empty[GM_STAGES], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};

template <int BN, int EPI>
__global__ void __launch_bounds__(192, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const GemmArgs p) {
    extern __shared__ __align__(1024) uint8_t smem_bytes[];
    GemmSmem<BN>& sm = *reinterpret_cast<GemmSmem<BN>*>(smem_bytes);
    constexpr int NACC = (2 * BN <= 512) ? 2 : 1;  // accumulator buffers that fit the 512 TMEM columns
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int nkb = p.K / GM_BK;
    const int ntn = p.N / BN, ntm = (p.M + GM_BM - 1) / GM_BM;
    const int ntiles = ntn * ntm;

    if (tid == 0) {
        for (int i = 0; i < GM_STAGES; i++) {
            mbar_init(&sm.full[i], 1);
            mbar_init(&sm.empty[i], 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&sm.tmem_full[i], 1);
            mbar_init(&sm.tmem_empty[i], 4);  // one arrive per epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == 1) {  // TMEM allocation is a warp-wide operation; NACC x BN fp32 columns x 128 lanes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "n"(NACC * BN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_c = sm.tmem_base;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one_lane()) {
            tma_prefetch_desc(&tm_a);
            tma_prefetch_desc(&tm_b);
            int it = 0;  // running k-block counter across tiles (ring position)
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int m0 = (tile / ntn) * GM_BM, n0 = (tile % ntn) * BN;
                for (int kb = 0; kb < nkb; kb++, it++) {
                    const int s = it % GM_STAGES;
                    if (it >= GM_STAGES) mbar_wait(&sm.empty[s], ((it / GM_STAGES) - 1) & 1);
                    mbar_arrive_expect_tx(&sm.full[s], (GM_BM + BN) * GM_BK * 2);
                    tma_load_2d(&sm.a[s][0], &tm_a, kb * GM_BK, m0, &sm.full[s]);
                    tma_load_2d(&sm.b[s][0], &tm_b, kb * GM_BK, n0, &sm.full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // (elect.sync, not `lane == 0`: only then does the compiler treat the descriptor / TMEM-address operands as
        //  warp-uniform and issue tcgen05.mma back to back instead of inside a vote/elect/R2UR loop of ~120 cycles)
        if (elect_one_lane()) {
            // instruction descriptor (kind::f16): D=f32, A=B=bf16, both K-major, N>>3 at [17,23), M>>4 at [24,29)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(GM_BM >> 4) << 24);
            int it = 0, lt = 0;  // lt: local tile counter
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, lt++) {
                const int acc = lt % NACC;
                if (lt >= NACC) mbar_wait(&sm.tmem_empty[acc], ((lt / NACC) - 1) & 1);  // epilogue drained this buffer
                tc_fence_after();
                const uint32_t tc = tmem_c + (uint32_t)(acc * BN);
                for (int kb = 0; kb < nkb; kb++, it++) {
                    const int s = it % GM_STAGES;
                    mbar_wait(&sm.full[s], (it / GM_STAGES) & 1);
                    tc_fence_after();
                    const uint64_t da = umma_desc_sw128(&sm.a[s][0]), db = umma_desc_sw128(&sm.b[s][0]);
#pragma unroll
                    for (int k = 0; k < GM_BK / 16; k++)  // 16 bf16 = 32 B along K inside the 128 B swizzle atom: +2 in the address field
                        umma_bf16(tc, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    umma_commit(&sm.empty[s]);  // arrives when the MMAs above have consumed the stage
                }
                umma_commit(&sm.tmem_full[acc]);
            }
        }
    } else {
        // ===================== epilogue warps (2..5) =====================
        const int q = warp & 3;  // TMEM lane quadrant this warp may access
        int lt = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, lt++) {
            const int acc = lt % NACC;
            const int m0 = (tile / ntn) * GM_BM, n0 = (tile % ntn) * BN;
            const int row = m0 + 32 * q + lane;
            mbar_wait(&sm.tmem_full[acc], (lt / NACC) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_c + ((uint32_t)(32 * q) << 16) + (uint32_t)(acc * BN + c), r);
                if (row < p.M) {
                    const size_t off = (size_t)row * p.N + n0 + c;
                    uint4 out[4];
                    uint32_t* o = reinterpret_cast<uint32_t*>(out);
                    uint4 res[4];
                    if (EPI == GM_EPI_ADD) {
#pragma unroll
                        for (int i = 0; i < 4; i++) res[i] = *reinterpret_cast<const uint4*>(p.R + off + 8 * i);
                    }
                    const uint32_t* rr = reinterpret_cast<const uint32_t*>(res);
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        float x0 = __uint_as_float(r[2 * i]), x1 = __uint_as_float(r[2 * i + 1]);
                        if (EPI == GM_EPI_RELU_SQ) {
                            // eager graph: key() -> bf16, relu, **2 -> bf16
                            x0 = __bfloat162float(__float2bfloat16_rn(fmaxf(x0, 0.f)));
                            x1 = __bfloat162float(__float2bfloat16_rn(fmaxf(x1, 0.f)));
                            x0 *= x0;
                            x1 *= x1;
                        } else if (EPI == GM_EPI_ADD) {
                            x0 += bf16lo_to_f32(rr[i]);
                            x1 += bf16hi_to_f32(rr[i]);
                        }
                        o[i] = pack_bf16x2(x0, x1);
                    }
#pragma unroll
                    for (int i = 0; i < 4; i++) *reinterpret_cast<uint4*>(p.C + off + 8 * i) = out[i];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.tmem_empty[acc]);
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_c), "n"(NACC * BN) : "memory");
    }
}

}  // namespace vrwkv

using namespace vrwkv;

template <int BN, int EPI>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& a, cudaStream_t st) {
    auto kern = gemm_bf16_tn_kernel<BN, EPI>;
    const size_t smem = sizeof(GemmSmem<BN>) + 1024;
    VRWKV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ntiles = (a.N / BN) * ((a.M + GM_BM - 1) / GM_BM);
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    dim3 grid(ntiles < nsm ? ntiles : nsm), block(192);
    kern<<<grid, block, smem, st>>>(ta, tb, a);
    VRWKV_CUDA(cudaGetLastError());
    vrwkv_count_launch(1);
    return VRWKV_OK;
}

extern "C" int vrwkv_gemm_bf16_tn(int M, int N, int K, const uint16_t* A, const uint16_t* B, uint16_t* C, int epilogue,
                                  const uint16_t* R, void* stream) {
    if (M <= 0 || N <= 0 || K <= 0) return vrwkv_fail(VRWKV_EINVAL, "gemm: bad shape (%d,%d,%d)", M, N, K);
    if (K % GM_BK) return vrwkv_fail(VRWKV_EUNSUP, "gemm: K=%d must be a multiple of %d", K, GM_BK);
    if (N % 128) return vrwkv_fail(VRWKV_EUNSUP, "gemm: N=%d must be a multiple of 128", N);
    if (!A || !B || !C || (epilogue == GM_EPI_ADD && !R)) return vrwkv_fail(VRWKV_EINVAL, "gemm: null pointer");
    if ((((uintptr_t)A) | ((uintptr_t)B) | ((uintptr_t)C) | ((uintptr_t)R)) & 15) return vrwkv_fail(VRWKV_EINVAL, "gemm: pointers must be 16-byte aligned");
    const int BN = (N % 256 == 0) ? 256 : 128;
    CUtensorMap ta, tb;
    int rc;
    if ((rc = vrwkv_encode_2d(&ta, A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)K, (uint64_t)M, (uint64_t)K * 2, GM_BK, GM_BM,
                              CU_TENSOR_MAP_SWIZZLE_128B)))
        return rc;
    if ((rc = vrwkv_encode_2d(&tb, B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)K, (uint64_t)N, (uint64_t)K * 2, GM_BK, BN,
                              CU_TENSOR_MAP_SWIZZLE_128B)))
        return rc;
    GemmArgs a{M, N, K, C, R};
    cudaStream_t st = (cudaStream_t)stream;
#define VRWKV_GEMM_CASE(bn, epi) \
    if (BN == bn && epilogue == epi) return launch_gemm<bn, epi>(ta, tb, a, st);
    VRWKV_GEMM_CASE(256, GM_EPI_NONE) VRWKV_GEMM_CASE(256, GM_EPI_RELU_SQ) VRWKV_GEMM_CASE(256, GM_EPI_ADD)
    VRWKV_GEMM_CASE(128, GM_EPI_NONE) VRWKV_GEMM_CASE(128, GM_EPI_RELU_SQ) VRWKV_GEMM_CASE(128, GM_EPI_ADD)
#undef VRWKV_GEMM_CASE
    return vrwkv_fail(VRWKV_EINVAL, "gemm: unknown epilogue %d", epilogue);
}
