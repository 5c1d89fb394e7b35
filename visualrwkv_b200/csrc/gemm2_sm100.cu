// gemm2_sm100.cu — bf16 GEMM on CTA PAIRS (tcgen05.mma.cta_group::2, M = 256 across two SMs), every operand layout the
// RWKV blocks need in forward and backward, element-wise neighbours fused into the epilogue.
//
//   C[M,N] = epilogue( op(A) . op(B) )     fp32 accumulation in TMEM, bf16 (or fp32-atomic) output, row-major C
//     A_MN = 0: A is [M,K] row-major (K contiguous)      A_MN = 1: A is [K,M] row-major (M contiguous)  -> A^T . B
//     B_MN = 0: B is [N,K] row-major (nn.Linear weight)  B_MN = 1: B is [K,N] row-major (N contiguous)
//   forward  y = x W^T          : A_MN 0, B_MN 0      (VisualRWKV-v7/v7.00/src/model.py:175-178,194,225-227,325)
//   dgrad    dx = dy W          : A_MN 0, B_MN 1      (W is [N_out, K_in]: the contraction index is its row)
//   wgrad    dW = dy^T x        : A_MN 1, B_MN 1      (contraction over the 16384 token rows; optional split along it)
//   epilogues: NONE, RELU_SQ (C = relu(acc)^2, model.py:225), ADD (C = R + acc: residual / accumulate), RELUSQ_BWD
//   (C = acc * 2 sqrt(R): the backward of relu()^2 from the saved activation R, fused into the dgrad that produces acc),
//   ATOMIC_F32 (fp32 red.add partial sums of a split contraction; the last slice writes the bf16 result).
//
// Why pairs: the 1-CTA kernel (gemm_sm100.cu) streams 48 KB per 64-wide k-block per SM and is L2-bandwidth bound at ~40 %
// tensor-pipe utilisation (148 SMs x 48 KB / 512 cycles = 14 KB/clk vs ~6 KB/clk the L2 delivers).  With cta_group::2
// the two CTAs of a cluster each stage their own 128 rows of A and HALF of the 256-wide B tile; the pair's MMA reads B
// halves from both shared memories: 32 KB per k-block per SM.
//
// Structure per CTA (320 threads), persistent over 256 x BN tiles:
//   warp 0   : TMA producer — its own A rows and B half into a STAGES-deep ring; completion bytes are signalled on the
//              LEADER CTA's full barrier (mbarrier address with the peer bit cleared);
//   warp 1   : TMEM allocation (cta_group::2, both CTAs); in the leader CTA one elected lane issues
//              tcgen05.mma.cta_group::2.kind::f16 (M=256, N=BN, K=16) and releases ring stages / publishes accumulators
//              with tcgen05.commit.cta_group::2 multicast to both CTAs' barriers;
//   warps 2-9: epilogue of this CTA's 128 accumulator rows (TMEM lanes; two warps per lane quadrant, half of the columns
//              each), double-buffered against the main loop.  With K = 768 a tile is only 6 k cycles of tensor-core work,
//              so the epilogue (TMEM -> registers -> fp32 math -> bf16 -> global) must run at >= that pace.
// Operand tiles are bf16 64x64 boxes (8 KB, SWIZZLE_128B): K-major operands take [rows][64 k] boxes, MN-major operands
// [64 k-lines][64 m/n] boxes — the same TMA box shape, only the UMMA descriptor differs (umma.cuh).
#include <cudaTypedefs.h>
#include <cstdlib>

#include "common.cuh"
#include "host_util.h"
#include "umma.cuh"

namespace vrwkv {

constexpr int G2_BM = 128;     // rows per CTA (256 per pair)
constexpr int G2_BK = 64;
enum { G2_EPI_NONE = 0, G2_EPI_RELU_SQ = 1, G2_EPI_ADD = 2, G2_EPI_ATOMIC_F32 = 3, G2_EPI_RELUSQ_BWD = 4,
       G2_EPI_BIAS = 5, G2_EPI_BIAS_GELU = 6, G2_EPI_BIAS_ADD = 7,    // the SigLIP tower's Linear layers (bias; tanh-GELU; + residual)
       G2_EPI_ACT = 8, G2_EPI_ACT_BWD = 9 };   // per-group activation (none / tanh / sigmoid) and its backward from the saved output:
                                               // the LoRA branches of RWKV_Tmix_x070 (model.py:176,181-184)

constexpr int G2_MAXG = 4;   // problems of identical shape in one launch (r/k/v projections, the four C x C weight gradients, ...)
struct Gemm2Args {
    int M, N, K;        // C is [M,N]; K = contraction length handled by this launch (all splits together)
    int ksplit;         // number of slices of the contraction (grid covers groups x tiles x ksplit); > 1 only with ATOMIC_F32
    int ngroups;
    uint16_t* C[G2_MAXG];
    const uint16_t* R[G2_MAXG];  // residual (EPI_ADD)
    int ct[G2_MAXG];             // 1: store this group's result transposed (C[g] is [N,M]); EPI_NONE only
    const uint16_t* bias[G2_MAXG];  // [N] (EPI_BIAS*)
    int act[G2_MAXG];            // EPI_ACT / EPI_ACT_BWD: 0 none, 1 tanh, 2 sigmoid
    int mv[G2_MAXG], nv[G2_MAXG];   // this group's real result extents (<= M, N): its C / R have nv columns; TMA clips the rest
    int r_rows;                  // rows of R (EPI_BIAS_ADD: R row = row % r_rows — a position table shared by all images); 0 = M
    int dbg_mode;                // development (VRWKV_GEMM2_DBG): 1 = no TMA traffic after the ring is primed (results wrong: tensor-core
                                 // ceiling), 2 = no MMAs (load ceiling)
    float* Cf;          // ATOMIC_F32: fp32 partial sums [ngroups][M][N]; must be zero on entry, is zero again on exit
    int* tickets;       // ATOMIC_F32: one counter per (group, tile, CTA of the pair); zero on entry and on exit
};
struct Gemm2Maps {
    CUtensorMap a[G2_MAXG], b[G2_MAXG];
    CUtensorMap c[G2_MAXG];   // C as [M rows][N cols], box [32 rows][64 cols], SWIZZLE_128B: the epilogue's TMA stores
};

__device__ __forceinline__ bool g2_elect() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
        "r"(rank)
        : "memory");
}
// 2-D TMA tile into THIS CTA's shared memory, completion bytes on the LEADER CTA's barrier (peer bit cleared)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* m, int x, int y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(x), "r"(y)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_c),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {   // arrives on `bar` in both CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}

template <int BN, int STAGES>
struct alignas(1024) Gemm2Smem {
    uint8_t a[STAGES][G2_BM * G2_BK * 2];         // 16 KB per stage: this CTA's 128 rows of A (two 8 KB boxes when MN-major)
    uint8_t b[STAGES][(BN / 2) * G2_BK * 2];      // this CTA's half of the B tile
    uint8_t stg[8][2][4096];                      // per epilogue warp: two [32 rows][64 cols] bf16 staging tiles for the TMA stores
    uint64_t full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ float g2_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

template <int BN, int EPI, int A_MN, int B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
gemm2_kernel(const __grid_constant__ Gemm2Maps maps, const Gemm2Args p) {
    constexpr int STAGES = (BN == 256) ? 5 : 6;
    constexpr int NACC = 2;
    constexpr uint32_t STAGE_BYTES = G2_BM * G2_BK * 2 + (BN / 2) * G2_BK * 2;
    extern __shared__ __align__(1024) uint8_t g2_smem[];
    Gemm2Smem<BN, STAGES>& sm = *reinterpret_cast<Gemm2Smem<BN, STAGES>*>((reinterpret_cast<uintptr_t>(g2_smem) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int npairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
    const int ntn = p.N / BN, ntm = (p.M + 2 * G2_BM - 1) / (2 * G2_BM);
    const int nkb_total = p.K / G2_BK, nkb = nkb_total / p.ksplit;   // k-blocks per work item
    const int ntiles = ntn * ntm;
    const int nwork = p.ngroups * ntiles * p.ksplit;   // work item = (group, tile, k-slice), k-slice fastest

    if (tid == 0) {
        for (int i = 0; i < STAGES; i++) {
            mbar_init(&sm.full[i], 1);
            mbar_init(&sm.empty[i], 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&sm.tmem_full[i], 1);
            mbar_init(&sm.tmem_empty[i], 16);  // 8 epilogue warps of each CTA (only the leader's copy is waited on)
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "n"(NACC * BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_c = sm.tmem_base;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (g2_elect()) {
            for (int g = 0; g < p.ngroups; g++) {
                tma_prefetch_desc(&maps.a[g]);
                tma_prefetch_desc(&maps.b[g]);
            }
            int it = 0;
            for (int wk = pair; wk < nwork; wk += npairs) {
                const int gt = wk / p.ksplit, ks = wk - gt * p.ksplit, g = gt / ntiles, tile = gt - g * ntiles;
                const CUtensorMap* tm_a_ = &maps.a[g];
                const CUtensorMap* tm_b_ = &maps.b[g];
                const int m0 = (tile / ntn) * 2 * G2_BM + (int)rank * G2_BM, n0 = (tile % ntn) * BN + (int)rank * (BN / 2);
                for (int kb = 0; kb < nkb; kb++, it++) {
                    const int s = it % STAGES, k0 = (ks * nkb + kb) * G2_BK;
                    if (it >= STAGES) mbar_wait(&sm.empty[s], ((it / STAGES) - 1) & 1);
                    if (p.dbg_mode == 1 && it >= STAGES) {
                        if (leader) mbar_arrive(&sm.full[s]);
                        continue;
                    }
                    if (leader) mbar_arrive_expect_tx(&sm.full[s], 2 * STAGE_BYTES);
                    if (A_MN) {
                        tma_load_2d_pair(&sm.a[s][0], tm_a_, m0, k0, &sm.full[s]);
                        tma_load_2d_pair(&sm.a[s][BT_BYTES], tm_a_, m0 + 64, k0, &sm.full[s]);
                    } else {
                        tma_load_2d_pair(&sm.a[s][0], tm_a_, k0, m0, &sm.full[s]);
                    }
                    if (B_MN) {
#pragma unroll
                        for (int i = 0; i < BN / 128; i++) tma_load_2d_pair(&sm.b[s][i * BT_BYTES], tm_b_, n0 + 64 * i, k0, &sm.full[s]);
                    } else {
                        tma_load_2d_pair(&sm.b[s][0], tm_b_, k0, n0, &sm.full[s]);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader && g2_elect()) {
            constexpr uint32_t idesc = umma_idesc_bf16_mj(2 * G2_BM, BN, A_MN, B_MN);
            const uint32_t b4 = smem_u32(&sm) >> 4;
            int it = 0, lt = 0;
            for (int wk = pair; wk < nwork; wk += npairs, lt++) {
                const int acc = lt % NACC;
                if (lt >= NACC) mbar_wait(&sm.tmem_empty[acc], ((lt / NACC) - 1) & 1);   // both CTAs' epilogues drained this buffer
                tc_fence_after();
                const uint32_t tc = tmem_c + (uint32_t)(acc * BN);
                for (int kb = 0; kb < nkb; kb++, it++) {
                    const int s = it % STAGES;
                    mbar_wait(&sm.full[s], (it / STAGES) & 1);
                    tc_fence_after();
                    const uint32_t oa = (uint32_t)(&sm.a[s][0] - (uint8_t*)&sm), ob = (uint32_t)(&sm.b[s][0] - (uint8_t*)&sm);
#pragma unroll
                    for (int k = 0; k < G2_BK / 16; k++) {
                        if (p.dbg_mode == 2) break;
                        const uint64_t da = A_MN ? bdesc_mn(b4, oa + k * 2048) : bdesc_k(b4, oa + k * 32);
                        const uint64_t db = B_MN ? bdesc_mn(b4, ob + k * 2048) : bdesc_k(b4, ob + k * 32);
                        umma_bf16_pair(tc, da, db, idesc, (kb | k) != 0);
                    }
                    umma_commit_pair(&sm.empty[s]);
                }
                umma_commit_pair(&sm.tmem_full[acc]);
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue warps (2..9) of both CTAs =====================
        const int q = warp & 3, half = (warp - 2) >> 2;   // TMEM lane quadrant; column half
        int lt = 0;
        // Residual operand of the epilogue (ADD / RELUSQ_BWD / BIAS_ADD / ACT_BWD): this warp's 32 x 64 tiles are fetched with
        // row-contiguous accesses (8 lanes x 16 bytes = one 128-byte line per row, 4 rows per instruction), ONE TILE AHEAD, and
        // pass through the staging tile to the lane that owns the row.  (One 16-byte load per lane from 32 different rows
        // made the relu^2-backward GEMM 3.4x slower than its forward twin; loading at the start of the tile's own epilogue
        // still exposed the round trip: 104 vs 60 us with a plain residual add at N = 3072.)
        constexpr bool NEED_R = EPI == G2_EPI_ADD || EPI == G2_EPI_RELUSQ_BWD || EPI == G2_EPI_BIAS_ADD || EPI == G2_EPI_ACT_BWD;
        constexpr int NCH = BN / 128;   // 64-column chunks per warp per tile
        uint4 rpre[NEED_R ? NCH : 1][8], rnext[NEED_R ? NCH : 1][8];
        auto load_r = [&](int wk_, uint4 (&dst)[NEED_R ? NCH : 1][8]) {
            const int gt_ = wk_ / p.ksplit, g_ = gt_ / ntiles, tile_ = gt_ - g_ * ntiles;
            const int m0_ = (tile_ / ntn) * 2 * G2_BM + (int)rank * G2_BM, n0_ = (tile_ % ntn) * BN;
            const int mvg = p.mv[g_], nvg = p.nv[g_];
            const uint16_t* const Rg_ = p.R[g_];
#pragma unroll
            for (int ch = 0; ch < NCH; ch++)
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int grow = m0_ + 32 * q + 4 * k + (lane >> 3), gcol = n0_ + half * (BN / 2) + 64 * ch + 8 * (lane & 7);
                    const int srow = (EPI == G2_EPI_BIAS_ADD && p.r_rows) ? grow % p.r_rows : grow;
                    dst[ch][k] = (grow < mvg && gcol < nvg) ? *reinterpret_cast<const uint4*>(Rg_ + (size_t)srow * nvg + gcol)
                                                            : make_uint4(0u, 0u, 0u, 0u);
                }
        };
        if constexpr (NEED_R) {
            if (pair < nwork) load_r(pair, rnext);
        }
        int chunk_i = 0;   // staging buffers alternate across tiles too: with BN = 128 a tile is ONE chunk per warp, and the bulk store of
                           // the previous tile may still be reading the other buffer (tma_store_wait_read<1> leaves one in flight)
        for (int wk = pair; wk < nwork; wk += npairs, lt++) {
            const int gt = wk / p.ksplit, g = gt / ntiles, tile = gt - g * ntiles;
            const int acc = lt % NACC;
            const int m0 = (tile / ntn) * 2 * G2_BM + (int)rank * G2_BM, n0 = (tile % ntn) * BN;
            const int row = m0 + 32 * q + lane;
            uint16_t* const Cg = p.C[g];
            const uint16_t* const Rg = p.R[g];
            float* const Cfg = (EPI == G2_EPI_ATOMIC_F32) ? p.Cf + (size_t)g * p.M * p.N : nullptr;
            // Residual operand of the epilogue (ADD / RELUSQ_BWD / BIAS_ADD / ACT_BWD): this warp's 32 x 64 tiles are fetched
            // with row-contiguous accesses (8 lanes x 16 bytes = one 128-byte line per row, 4 rows per instruction) before the
            // wait for the accumulator, and pass through the staging tile to the lane that owns the row.  (One 16-byte load per
            // lane from 32 different rows made the relu^2-backward GEMM 3.4x slower than its forward twin: 198 vs 59 us.)
            if constexpr (NEED_R) {
#pragma unroll
                for (int ch = 0; ch < NCH; ch++)
#pragma unroll
                    for (int k = 0; k < 8; k++) rpre[ch][k] = rnext[ch][k];
                if (wk + npairs < nwork) load_r(wk + npairs, rnext);   // one tile ahead: in flight during this tile's epilogue
            }
            mbar_wait(&sm.tmem_full[acc], (lt / NACC) & 1);
            tc_fence_after();
            // Plain results leave through shared memory and TMA stores (full 128-byte lines per row): 32 lanes writing 16 bytes
            // to 32 different rows each cost one L2 transaction per lane and held the epilogue at ~12 k cycles per tile — twice
            // the tensor-core time of a K = 768 tile.  The atomic (split-K) and transposed epilogues keep register stores.
            const bool via_tma = EPI != G2_EPI_ATOMIC_F32 && !(EPI == G2_EPI_NONE && p.ct[g]);
#pragma unroll 1
            for (int c64 = half * (BN / 2); c64 < (half + 1) * (BN / 2); c64 += 64, chunk_i++) {
                uint8_t* const stile = &sm.stg[warp - 2][chunk_i & 1][0];
                if (via_tma) {
                    if (g2_elect()) tma_store_wait_read<1>();   // the store that read this staging tile two chunks ago is done
                    __syncwarp();
                }
                if constexpr (NEED_R) {
                    const int ch = (c64 - half * (BN / 2)) >> 6;
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const int rt = 4 * k + (lane >> 3);
                        *reinterpret_cast<uint4*>(stile + rt * 128 + (((lane & 7) ^ (rt & 7)) << 4)) = rpre[ch][k];
                    }
                    __syncwarp();
                }
#pragma unroll
                for (int sub = 0; sub < 2; sub++) {
                const int c = c64 + 32 * sub;
                uint32_t r[32];
                tmem_ld32(tmem_c + ((uint32_t)(32 * q) << 16) + (uint32_t)(acc * BN + c), r);
                if (row < p.M || via_tma) {
                    const size_t off = (size_t)row * p.N + n0 + c;
                    const int nvg = p.nv[g];
                    const size_t offr = (size_t)row * nvg + n0 + c;   // residuals have the group's real width
                    if (EPI == G2_EPI_ATOMIC_F32) {
#pragma unroll
                        for (int i = 0; i < 8; i++)
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(Cfg + off + 4 * i), "f"(__uint_as_float(r[4 * i])),
                                         "f"(__uint_as_float(r[4 * i + 1])), "f"(__uint_as_float(r[4 * i + 2])), "f"(__uint_as_float(r[4 * i + 3]))
                                         : "memory");
                    } else {
                        uint4 out[4], res[4];
                        uint32_t* o = reinterpret_cast<uint32_t*>(out);
                        uint4 bia[4];
                        const bool inb = row < p.mv[g];
                        if constexpr (NEED_R) {
#pragma unroll
                            for (int i = 0; i < 4; i++)
                                res[i] = *reinterpret_cast<const uint4*>(stile + lane * 128 + (((4 * sub + i) ^ (lane & 7)) << 4));
                        }
                        if (EPI == G2_EPI_BIAS || EPI == G2_EPI_BIAS_GELU || EPI == G2_EPI_BIAS_ADD) {
#pragma unroll
                            for (int i = 0; i < 4; i++) bia[i] = (n0 + c + 8 * i < nvg) ? __ldg(reinterpret_cast<const uint4*>(p.bias[g] + n0 + c + 8 * i)) : make_uint4(0u, 0u, 0u, 0u);
                        }
                        const uint32_t* bb = reinterpret_cast<const uint32_t*>(bia);
                        const uint32_t* rr = reinterpret_cast<const uint32_t*>(res);
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            float x0 = __uint_as_float(r[2 * i]), x1 = __uint_as_float(r[2 * i + 1]);
                            if (EPI == G2_EPI_RELU_SQ) {
                                // eager graph: key() -> bf16, relu, **2 -> bf16
                                x0 = __bfloat162float(__float2bfloat16_rn(fmaxf(x0, 0.f)));
                                x1 = __bfloat162float(__float2bfloat16_rn(fmaxf(x1, 0.f)));
                                x0 *= x0;
                                x1 *= x1;
                            } else if (EPI == G2_EPI_ADD) {
                                x0 += bf16lo_to_f32(rr[i]);
                                x1 += bf16hi_to_f32(rr[i]);
                            } else if (EPI == G2_EPI_BIAS || EPI == G2_EPI_BIAS_GELU || EPI == G2_EPI_BIAS_ADD) {
                                // nn.Linear in bf16: one rounding after the bias; then the tanh-GELU / residual add of the eager graph
                                x0 = __bfloat162float(__float2bfloat16_rn(x0 + bf16lo_to_f32(bb[i])));
                                x1 = __bfloat162float(__float2bfloat16_rn(x1 + bf16hi_to_f32(bb[i])));
                                if (EPI == G2_EPI_BIAS_GELU) {
                                    x0 = 0.5f * x0 * (1.f + tanhf(0.7978845608028654f * (x0 + 0.044715f * x0 * x0 * x0)));
                                    x1 = 0.5f * x1 * (1.f + tanhf(0.7978845608028654f * (x1 + 0.044715f * x1 * x1 * x1)));
                                } else if (EPI == G2_EPI_BIAS_ADD) {
                                    x0 += bf16lo_to_f32(rr[i]);
                                    x1 += bf16hi_to_f32(rr[i]);
                                }
                            } else if (EPI == G2_EPI_ACT) {
                                // eager graph: the matmul output is rounded to bf16, then tanh / sigmoid in fp32
                                x0 = __bfloat162float(__float2bfloat16_rn(x0));
                                x1 = __bfloat162float(__float2bfloat16_rn(x1));
                                const int ac = p.act[g];
                                if (ac == 1) { x0 = tanhf(x0); x1 = tanhf(x1); }
                                else if (ac == 2) { x0 = 1.f / (1.f + __expf(-x0)); x1 = 1.f / (1.f + __expf(-x1)); }
                            } else if (EPI == G2_EPI_ACT_BWD) {
                                // d/dx tanh = 1 - h^2, d/dx sigmoid = h (1 - h), from the saved output h
                                x0 = __bfloat162float(__float2bfloat16_rn(x0));
                                x1 = __bfloat162float(__float2bfloat16_rn(x1));
                                const float h0 = bf16lo_to_f32(rr[i]), h1 = bf16hi_to_f32(rr[i]);
                                const int ac = p.act[g];
                                if (ac == 1) { x0 *= (1.f - h0 * h0); x1 *= (1.f - h1 * h1); }
                                else if (ac == 2) { x0 *= (1.f - h0) * h0; x1 *= (1.f - h1) * h1; }
                            } else if (EPI == G2_EPI_RELUSQ_BWD) {
                                // eager graph: dact -> bf16, then d/dx relu(x)^2 = 2 relu(x) = 2 sqrt(act)
                                // MUFU sqrt: the IEEE sqrtf of a non-fast-math build is a ~30-instruction routine per element and made
                                // this epilogue 2.5x the tile's tensor-core time (dev_gemm2_epi: 152 vs 60 us)
                                x0 = __bfloat162float(__float2bfloat16_rn(x0)) * 2.f * g2_sqrt(bf16lo_to_f32(rr[i]));
                                x1 = __bfloat162float(__float2bfloat16_rn(x1)) * 2.f * g2_sqrt(bf16hi_to_f32(rr[i]));
                            }
                            o[i] = pack_bf16x2(x0, x1);
                        }
                        if (!via_tma) {
                            // C^T: lanes hold consecutive rows, so each of the 32 columns is one 64-byte run across the warp
                            const uint16_t* o16 = reinterpret_cast<const uint16_t*>(out);
                            if (inb) {
#pragma unroll
                                for (int i = 0; i < 32; i++)
                                    if (n0 + c + i < nvg) Cg[(size_t)(n0 + c + i) * p.mv[g] + row] = o16[i];
                            }
                        } else {
                            // staging tile row = lane, 128 bytes (64 columns), 16-byte chunks XOR-swizzled by (row & 7)
#pragma unroll
                            for (int i = 0; i < 4; i++)
                                *reinterpret_cast<uint4*>(stile + lane * 128 + (((4 * sub + i) ^ (lane & 7)) << 4)) = out[i];
                        }
                    }
                }
                }
                if (via_tma) {
                    fence_proxy_async();
                    __syncwarp();
                    if (g2_elect()) {
                        tma_store_2d(&maps.c[g], stile, n0 + c64, m0 + 32 * q);
                        tma_store_commit();
                    }
                    __syncwarp();
                }
            }
            if (EPI == G2_EPI_ATOMIC_F32) {
                // The last of the ksplit slices to finish this (group, tile, CTA) turns the fp32 sums into the bf16 result and
                // leaves zeros behind, so the workspace needs no memset between launches.
                __threadfence();
                __syncwarp();
                int* ticket = p.tickets + ((size_t)gt * 2 + rank) * 8 + (warp - 2);   // one per epilogue warp (its 32 rows x BN/2 columns)
                int prev = 0;
                if (lane == 0) prev = atomicAdd(ticket, 1);
                prev = __shfl_sync(0xffffffffu, prev, 0);
                if (prev == p.ksplit - 1) {
                    __threadfence();
                    if (lane == 0) *ticket = 0;
                    if (row < p.M) {
                        const int mvg = p.mv[g], nvg = p.nv[g];
                        // 32 columns per trip with all eight loads issued first: one L2 round trip per 128 bytes of a row instead of
                        // one per 32 bytes (this tail is part of the fixed cost of a split launch)
#pragma unroll 1
                        for (int c = half * (BN / 2); c < (half + 1) * (BN / 2); c += 32) {
                            float4* src = reinterpret_cast<float4*>(Cfg + (size_t)row * p.N + n0 + c);
                            float4 x[8];
#pragma unroll
                            for (int i = 0; i < 8; i++) x[i] = __ldcg(src + i);
#pragma unroll
                            for (int i = 0; i < 4; i++) {
                                uint4 o;
                                o.x = pack_bf16x2(x[2 * i].x, x[2 * i].y); o.y = pack_bf16x2(x[2 * i].z, x[2 * i].w);
                                o.z = pack_bf16x2(x[2 * i + 1].x, x[2 * i + 1].y); o.w = pack_bf16x2(x[2 * i + 1].z, x[2 * i + 1].w);
                                if (p.ct[g]) {   // C^T: consecutive lanes (rows) write consecutive bf16 of one output row
                                    const uint16_t* o16 = reinterpret_cast<const uint16_t*>(&o);
#pragma unroll
                                    for (int e = 0; e < 8; e++) {
                                        const int col = n0 + c + 8 * i + e;
                                        if (row < mvg && col < nvg) Cg[(size_t)col * mvg + row] = o16[e];
                                    }
                                } else if (row < mvg && n0 + c + 8 * i < nvg) {
                                    *reinterpret_cast<uint4*>(Cg + (size_t)row * nvg + n0 + c + 8 * i) = o;
                                }
                                src[2 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
                                src[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(&sm.tmem_empty[acc], 0);   // the leader's MMA warp waits for all 8 epilogue warps
        }
        if (g2_elect()) tma_store_wait<0>();   // this warp's bulk stores have landed before the CTA (and its shared memory) goes away
        __syncwarp();
    }
    tc_fence_before();
    cluster_sync_all();   // nobody leaves (or frees TMEM) while the partner may still signal / be read
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_c), "n"(NACC * BN) : "memory");
    }
}

}  // namespace vrwkv

using namespace vrwkv;

template <int BN, int EPI, int A_MN, int B_MN>
static int launch_gemm2(const Gemm2Maps& maps, const Gemm2Args& a, cudaStream_t st) {
    auto kern = gemm2_kernel<BN, EPI, A_MN, B_MN>;
    constexpr int STAGES = (BN == 256) ? 5 : 6;
    const size_t smem = sizeof(Gemm2Smem<BN, STAGES>) + 1024;
    VRWKV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int nwork = a.ngroups * (a.N / BN) * ((a.M + 2 * G2_BM - 1) / (2 * G2_BM)) * a.ksplit;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const int npairs = nwork < nsm / 2 ? nwork : nsm / 2;
    kern<<<dim3(2 * npairs), dim3(320), smem, st>>>(maps, a);
    VRWKV_CUDA(cudaGetLastError());
    vrwkv_count_launch(1);
    return VRWKV_OK;
}

// Split-K workspace (fp32 partial sums + tickets): one per device, grown on demand, zeroed when (re)allocated and left
// zero by every launch (the finalising CTA cleans up behind itself).  All launches that use it must be stream-ordered
// with respect to each other — true for the training step, which runs on one stream.
struct G2Workspace {
    float* cf = nullptr;
    int* tickets = nullptr;
    size_t cf_elems = 0, n_tickets = 0;
};
static G2Workspace g_ws[64];

static int g2_workspace(size_t cf_elems, size_t n_tickets, cudaStream_t st, G2Workspace** out) {
    int dev = 0;
    VRWKV_CUDA(cudaGetDevice(&dev));
    G2Workspace& w = g_ws[dev & 63];
    if (w.cf_elems < cf_elems) {
        if (w.cf) VRWKV_CUDA(cudaFree(w.cf));   // synchronises: nothing is in flight on the old buffer afterwards
        VRWKV_CUDA(cudaMalloc((void**)&w.cf, cf_elems * sizeof(float)));
        VRWKV_CUDA(cudaMemsetAsync(w.cf, 0, cf_elems * sizeof(float), st));
        w.cf_elems = cf_elems;
    }
    if (w.n_tickets < n_tickets) {
        if (w.tickets) VRWKV_CUDA(cudaFree(w.tickets));
        VRWKV_CUDA(cudaMalloc((void**)&w.tickets, n_tickets * sizeof(int)));
        VRWKV_CUDA(cudaMemsetAsync(w.tickets, 0, n_tickets * sizeof(int), st));
        w.n_tickets = n_tickets;
    }
    *out = &w;
    return VRWKV_OK;
}

// layout bit 0: A is [K,M] (M contiguous); bit 1: B is [K,N] (N contiguous).  `ngroups` problems of identical shape
// (A[g], B[g]) -> C[g] share one launch.  ksplit > 1 slices the contraction; the slices meet in fp32 and the last one
// writes the bf16 result (no element-wise epilogue in that mode).
extern "C" int vrwkv_gemm2_bf16_grouped(int M, int N, int K, int ngroups, const uint16_t* const* A, const uint16_t* const* B,
                                        uint16_t* const* C, const uint16_t* const* R, const int* c_transposed, int layout, int epilogue,
                                        int ksplit, const uint16_t* const* bias, int r_rows, const int* act, const int* dims, void* stream) {
    if (M <= 0 || N <= 0 || K <= 0) return vrwkv_fail(VRWKV_EINVAL, "gemm2: bad shape (%d,%d,%d)", M, N, K);
    if (ngroups < 1 || ngroups > G2_MAXG) return vrwkv_fail(VRWKV_EINVAL, "gemm2: 1..%d groups (got %d)", G2_MAXG, ngroups);
    if (ksplit < 1) ksplit = 1;
    if (K % (G2_BK * ksplit)) return vrwkv_fail(VRWKV_EUNSUP, "gemm2: K=%d must be a multiple of %d x ksplit=%d", K, G2_BK, ksplit);
    if (N % 128) return vrwkv_fail(VRWKV_EUNSUP, "gemm2: N=%d must be a multiple of 128", N);
    const int a_mn = layout & 1, b_mn = (layout >> 1) & 1;
    if (a_mn && (M % 8)) return vrwkv_fail(VRWKV_EUNSUP, "gemm2: M=%d must be a multiple of 8 when A is stored [K,M]", M);
    if (ksplit > 1 && epilogue != G2_EPI_NONE) return vrwkv_fail(VRWKV_EINVAL, "gemm2: no element-wise epilogue with ksplit > 1");
    if (epilogue < 0 || epilogue > G2_EPI_ACT_BWD || epilogue == G2_EPI_ATOMIC_F32) return vrwkv_fail(VRWKV_EINVAL, "gemm2: unknown epilogue %d", epilogue);
    const bool need_r = epilogue == G2_EPI_ADD || epilogue == G2_EPI_RELUSQ_BWD || epilogue == G2_EPI_BIAS_ADD || epilogue == G2_EPI_ACT_BWD;
    const bool need_b = epilogue >= G2_EPI_BIAS && epilogue <= G2_EPI_BIAS_ADD;
    if (need_b && (layout != 0 || !bias)) return vrwkv_fail(VRWKV_EINVAL, "gemm2: bias epilogues need the [M,K] x [N,K] layout and a bias per group");
    int BN = (N % 256 == 0) ? 256 : 128;
    // Weight gradients whose 256-wide tiling leaves most CTA pairs idle (the four C x C gradients of a time-mix block: 36
    // tiles for 74 pairs) run better as twice as many 128-wide tiles than as two contraction slices meeting in fp32 atomics:
    // the split pays a fixed ~35 us per launch for the atomics to drain, the ticket and the last slice's conversion
    // (dev_lora.py: T(k-blocks) = 37 us + 0.19 us per k-block).
    if (BN == 256 && ksplit == 1 && a_mn && b_mn) {
        const int tiles256 = ngroups * ((M + 2 * G2_BM - 1) / (2 * G2_BM)) * (N / 256);
        if (tiles256 < 60) BN = 128;
    }
    cudaStream_t st = (cudaStream_t)stream;
    Gemm2Maps maps;
    Gemm2Args a{};
    a.M = M; a.N = N; a.K = K; a.ksplit = ksplit; a.ngroups = ngroups; a.r_rows = r_rows;
    {
        static const int dbg = [] { const char* e = getenv("VRWKV_GEMM2_DBG"); return e ? atoi(e) : 0; }();
        a.dbg_mode = dbg;
    }
    for (int g = 0; g < ngroups; g++) {
        if (!A[g] || !B[g] || !C[g] || (need_r && (!R || !R[g])) || (need_b && !bias[g])) return vrwkv_fail(VRWKV_EINVAL, "gemm2: null pointer (group %d)", g);
        if ((((uintptr_t)A[g]) | ((uintptr_t)B[g]) | ((uintptr_t)C[g]) | (R ? (uintptr_t)R[g] : 0)) & 15)
            return vrwkv_fail(VRWKV_EINVAL, "gemm2: pointers must be 16-byte aligned");
        int rc;
        // The group's real extents: its tensors are exactly [Mg | Ng | Kg] wide; the launch tiles the common (M, N, K) and TMA
        // zero-fills loads / clips stores beyond them (LoRA ranks 32..128 share one launch this way, unpadded).
        const uint64_t Mg = dims ? dims[3 * g] : M, Ng = dims ? dims[3 * g + 1] : N, Kg = dims ? dims[3 * g + 2] : K;
        if (Mg < 1 || Ng < 1 || Kg < 1 || Mg > (uint64_t)M || Ng > (uint64_t)N || Kg > (uint64_t)K || (Ng % 8) || (Kg % 8) || (a_mn && (Mg % 8)))
            return vrwkv_fail(VRWKV_EINVAL, "gemm2: group %d extents (%d,%d,%d) must be multiples of 8 within (%d,%d,%d)", g, (int)Mg, (int)Ng, (int)Kg, M, N, K);
        const bool transposed = c_transposed && c_transposed[g];
        a.mv[g] = (int)Mg; a.nv[g] = (int)Ng;
        // K-major operand: matrix [rows = M|N][cols = K], box [128 | BN/2 rows][64 cols];  MN-major: matrix [rows = K][cols = M|N], box [64][64]
        if (a_mn) rc = vrwkv_encode_2d(&maps.a[g], A[g], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Mg, Kg, Mg * 2, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B);
        else rc = vrwkv_encode_2d(&maps.a[g], A[g], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Kg, Mg, Kg * 2, G2_BK, G2_BM, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
        if (b_mn) rc = vrwkv_encode_2d(&maps.b[g], B[g], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Ng, Kg, Ng * 2, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B);
        else rc = vrwkv_encode_2d(&maps.b[g], B[g], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Kg, Ng, Kg * 2, G2_BK, BN / 2, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
        if ((rc = vrwkv_encode_2d(&maps.c[g], C[g], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, transposed ? Mg : Ng,
                                  transposed ? Ng : Mg, (transposed ? Mg : Ng) * 2, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B)))
            return rc;
        a.C[g] = C[g];
        a.R[g] = R ? R[g] : nullptr;
        a.ct[g] = (c_transposed && c_transposed[g]) ? 1 : 0;
        a.bias[g] = need_b ? bias[g] : nullptr;
        a.act[g] = act ? act[g] : 0;
        if (a.ct[g] && epilogue != G2_EPI_NONE) return vrwkv_fail(VRWKV_EINVAL, "gemm2: transposed store only with the plain epilogue");
    }
    for (int g = ngroups; g < G2_MAXG; g++) { maps.a[g] = maps.a[0]; maps.b[g] = maps.b[0]; maps.c[g] = maps.c[0]; }
    int epi = epilogue;
    if (ksplit > 1) {
        epi = G2_EPI_ATOMIC_F32;
        const size_t ntiles = (size_t)(N / BN) * ((M + 2 * G2_BM - 1) / (2 * G2_BM));
        G2Workspace* w;
        int rc = g2_workspace((size_t)ngroups * M * N, (size_t)ngroups * ntiles * 2 * 8, st, &w);
        if (rc) return rc;
        a.Cf = w->cf;
        a.tickets = w->tickets;
    }
#define G2_CASE(bn, e, am, bm) \
    if (BN == bn && epi == e && a_mn == am && b_mn == bm) return launch_gemm2<bn, e, am, bm>(maps, a, st);
#define G2_LAYOUTS(bn, e) G2_CASE(bn, e, 0, 0) G2_CASE(bn, e, 0, 1) G2_CASE(bn, e, 1, 1)
    G2_LAYOUTS(256, G2_EPI_NONE) G2_LAYOUTS(256, G2_EPI_ADD) G2_LAYOUTS(256, G2_EPI_ATOMIC_F32) G2_CASE(256, G2_EPI_RELU_SQ, 0, 0)
    G2_CASE(256, G2_EPI_RELUSQ_BWD, 0, 1)
    G2_LAYOUTS(128, G2_EPI_NONE) G2_LAYOUTS(128, G2_EPI_ADD) G2_LAYOUTS(128, G2_EPI_ATOMIC_F32) G2_CASE(128, G2_EPI_RELU_SQ, 0, 0)
    G2_CASE(128, G2_EPI_RELUSQ_BWD, 0, 1)
    G2_CASE(256, G2_EPI_ACT, 0, 1) G2_CASE(128, G2_EPI_ACT, 0, 1) G2_CASE(256, G2_EPI_ACT_BWD, 0, 0) G2_CASE(128, G2_EPI_ACT_BWD, 0, 0)
    G2_CASE(256, G2_EPI_BIAS, 0, 0) G2_CASE(256, G2_EPI_BIAS_GELU, 0, 0) G2_CASE(256, G2_EPI_BIAS_ADD, 0, 0)
    G2_CASE(128, G2_EPI_BIAS, 0, 0) G2_CASE(128, G2_EPI_BIAS_GELU, 0, 0) G2_CASE(128, G2_EPI_BIAS_ADD, 0, 0)
#undef G2_LAYOUTS
#undef G2_CASE
    return vrwkv_fail(VRWKV_EUNSUP, "gemm2: unsupported combination (layout %d, epilogue %d)", layout, epilogue);
}

extern "C" int vrwkv_gemm2_bf16(int M, int N, int K, const uint16_t* A, const uint16_t* B, uint16_t* C, int layout, int epilogue,
                                const uint16_t* R, int ksplit, void* stream) {
    return vrwkv_gemm2_bf16_grouped(M, N, K, 1, &A, &B, &C, &R, nullptr, layout, epilogue, ksplit, nullptr, 0, nullptr, nullptr, stream);
}
