// vit_attn.cu — the SigLIP tower's self-attention (forward only: the tower is frozen, VisualRWKV-v7/v7.00/src/model.py:
// 368-369) on the tcgen05 tensor cores, one CTA per (image, head):
//     S = Q K^T / sqrt(64)          softmax over the keys (fp32)          O = P V
// (transformers' SiglipAttention, the arithmetic VisualRWKV-v7/v7.01/src/model.py:347-352,448-454 calls)
// Sequences of up to 256 patches (224x224 / 16 -> 196; 256x256 -> 256) fit one tile: Q, K, V of the head are staged by TMA
// as [64 rows][64 channels] bf16 boxes (SWIZZLE_128B) straight out of the [rows, D] projection outputs — K and V tiles
// double as the K-major B operand of Q K^T and the MN-major B operand of P V, so nothing is transposed.  Per 128-query
// block: 4 MMAs (M128 N256 K16) put the scores in TMEM; each of the 128 threads owns one query row: max and
// exp(x - max) in two sweeps over its 256 TMEM columns, the un-normalised probabilities go to shared memory as the bf16
// A operand of the second product (16 MMAs, M128 N64 K16), and the row sum divides O on its way out.
// Also here: the two element-wise companions of the image path that are not GEMM epilogues — the patch im2col in front
// of the patch-embedding GEMM and the context gate x * sigmoid(g) of the projector (model.py:335-338) with its backward.
#include <cudaTypedefs.h>

#include "common.cuh"
#include "host_util.h"
#include "umma.cuh"

namespace vrwkv {

struct alignas(1024) AttnSmem {
    uint8_t q[4 * BT_BYTES];      // 256 query rows x 64 (four 64-row tiles)
    uint8_t k[4 * BT_BYTES];      // 256 key rows x 64
    uint8_t v[4 * BT_BYTES];      // 256 key rows x 64
    uint8_t p[4 * 2 * BT_BYTES];  // P of one 128-query block: 4 k-tiles (64 keys each) x [128 rows][64 keys]
    uint64_t bar_in, bar_mma;
    uint32_t tmem_base;
};

struct AttnArgs {
    int S, H, D;          // patches per image, heads, model width (= 64 H)
    uint16_t* o;          // [N*S, D]
    float scale;
};

__device__ __forceinline__ bool attn_elect() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

__global__ void __launch_bounds__(128, 1)
vit_attn_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                const AttnArgs p) {
    extern __shared__ __align__(1024) uint8_t attn_smem[];
    AttnSmem& sm = *reinterpret_cast<AttnSmem*>((reinterpret_cast<uintptr_t>(attn_smem) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int head = blockIdx.x, img = blockIdx.y;
    const int S = p.S;
    if (tid == 0) {
        mbar_init(&sm.bar_in, 1);
        mbar_init(&sm.bar_mma, 1);
        fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tmem_alloc<512>(&sm.tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    const int ntile = (S + 63) / 64;   // 64-row tiles that hold real rows (rows past the image are masked / never stored)
    if (warp == 0 && attn_elect()) {
        mbar_arrive_expect_tx(&sm.bar_in, (uint32_t)(3 * ntile) * BT_BYTES);
        for (int i = 0; i < ntile; i++) {
            tma_load_2d(sm.q + i * BT_BYTES, &tm_q, head * 64, img * S + 64 * i, &sm.bar_in);
            tma_load_2d(sm.k + i * BT_BYTES, &tm_k, head * 64, img * S + 64 * i, &sm.bar_in);
            tma_load_2d(sm.v + i * BT_BYTES, &tm_v, head * 64, img * S + 64 * i, &sm.bar_in);
        }
    }
    // tiles that were not loaded must still hold finite numbers: they enter the products with weight zero
    for (int i = ntile; i < 4; i++) {
        for (int o = tid * 16; o < (int)BT_BYTES; o += 128 * 16) {
            *reinterpret_cast<uint4*>(sm.q + i * BT_BYTES + o) = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(sm.k + i * BT_BYTES + o) = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(sm.v + i * BT_BYTES + o) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    fence_proxy_async();
    __syncthreads();
    mbar_wait(&sm.bar_in, 0);
    const uint32_t b4 = smem_u32(&sm) >> 4;
    const uint32_t O_Q = (uint32_t)(sm.q - (uint8_t*)&sm), O_K = (uint32_t)(sm.k - (uint8_t*)&sm), O_V = (uint32_t)(sm.v - (uint8_t*)&sm),
                   O_P = (uint32_t)(sm.p - (uint8_t*)&sm);
    constexpr uint32_t ID_S = umma_idesc_bf16_mj(128, 256, 0, 0), ID_O = umma_idesc_bf16_mj(128, 64, 0, 1);
    constexpr uint32_t C_S = 0, C_O = 256;
    const uint32_t tm_row = tmem + ((uint32_t)(32 * warp) << 16);
    uint32_t ph = 0;
    const int nqb = (S + 127) / 128;
    for (int qb = 0; qb < nqb; qb++) {
        // ---- scores of this query block: [128 x 256] = Q_qb K^T ----
        if (warp == 0) {
            if (attn_elect()) {
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < 4; ks++) umma_bf16(tmem + C_S, bdesc_k(b4, O_Q + qb * 2 * BT_BYTES + ks * 32), bdesc_k(b4, O_K + ks * 32), ID_S, ks > 0);
                umma_commit(&sm.bar_mma);
            }
            __syncwarp();
        }
        mbar_wait(&sm.bar_mma, ph & 1);
        ph++;
        tc_fence_after();
        __syncwarp();
        // ---- softmax of row r = 128 qb + tid over keys < S ----
        const int r = tid;
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < 256; c += 32) {
            uint32_t v[32];
            tmem_ld32(tm_row + C_S + c, v);
#pragma unroll
            for (int e = 0; e < 32; e++)
                if (c + e < S) mx = fmaxf(mx, __uint_as_float(v[e]));
        }
        float sum = 0.f;
        // eager graph: the scaled scores are rounded to bf16 before the fp32 softmax (the maximum of the rounded values is the
        // rounded maximum); exp(x) = exp2(x log2 e)
        const float mb = __bfloat162float(__float2bfloat16_rn(mx * p.scale)) * 1.4426950408889634f;
#pragma unroll 1
        for (int c = 0; c < 256; c += 32) {
            uint32_t v[32];
            tmem_ld32(tm_row + C_S + c, v);
            float pr[32];
#pragma unroll
            for (int e = 0; e < 32; e++) {
                const float x = __bfloat162float(__float2bfloat16_rn(__uint_as_float(v[e]) * p.scale));
                pr[e] = (c + e < S) ? exp2f(x * 1.4426950408889634f - mb) : 0.f;
                sum += pr[e];
            }
            // P as the K-major A operand: k-tile c/64, row r, 16-byte chunks of 8 keys
            uint8_t* dst = sm.p + (c >> 6) * 2 * BT_BYTES + (r >> 6) * BT_BYTES;
#pragma unroll
            for (int cc = 0; cc < 4; cc++) {
                uint4 o;
                o.x = pack_bf16x2(pr[8 * cc], pr[8 * cc + 1]);
                o.y = pack_bf16x2(pr[8 * cc + 2], pr[8 * cc + 3]);
                o.z = pack_bf16x2(pr[8 * cc + 4], pr[8 * cc + 5]);
                o.w = pack_bf16x2(pr[8 * cc + 6], pr[8 * cc + 7]);
                const int ch = ((c & 63) >> 3) + cc;
                *reinterpret_cast<uint4*>(dst + (uint32_t)(r & 63) * 128u + (((uint32_t)ch ^ ((uint32_t)r & 7u)) << 4)) = o;
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        // ---- O = P V ----
        if (warp == 0) {
            if (attn_elect()) {
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < 16; ks++)   // 256 keys: k-tile ks/4 of P (K-major), k-lines of V (MN-major, tiles back to back)
                    umma_bf16(tmem + C_O, bdesc_k(b4, O_P + (ks >> 2) * 2 * BT_BYTES + (ks & 3) * 32), bdesc_mn(b4, O_V + ks * 2048), ID_O, ks > 0);
                umma_commit(&sm.bar_mma);
            }
            __syncwarp();
        }
        mbar_wait(&sm.bar_mma, ph & 1);
        ph++;
        tc_fence_after();
        __syncwarp();
        {
            uint32_t v[32], w[32];
            tmem_ld32_nowait(tm_row + C_O, v);
            tmem_ld32_nowait(tm_row + C_O + 32, w);
            tmem_ld_wait();
            const int row = 128 * qb + r;
            if (row < S) {
                // eager graph: probabilities are normalised in fp32 and rounded to bf16 before P V; dividing the fp32 product by
                // the row sum instead differs by the rounding of P only (within the bf16 tolerance of the parity test)
                const float inv = 1.f / sum;
                uint16_t* dst = p.o + ((size_t)img * S + row) * p.D + head * 64;
#pragma unroll
                for (int cc = 0; cc < 4; cc++) {
                    uint4 o;
                    o.x = pack_bf16x2(__uint_as_float(v[8 * cc]) * inv, __uint_as_float(v[8 * cc + 1]) * inv);
                    o.y = pack_bf16x2(__uint_as_float(v[8 * cc + 2]) * inv, __uint_as_float(v[8 * cc + 3]) * inv);
                    o.z = pack_bf16x2(__uint_as_float(v[8 * cc + 4]) * inv, __uint_as_float(v[8 * cc + 5]) * inv);
                    o.w = pack_bf16x2(__uint_as_float(v[8 * cc + 6]) * inv, __uint_as_float(v[8 * cc + 7]) * inv);
                    *reinterpret_cast<uint4*>(dst + 8 * cc) = o;
                    uint4 o2;
                    o2.x = pack_bf16x2(__uint_as_float(w[8 * cc]) * inv, __uint_as_float(w[8 * cc + 1]) * inv);
                    o2.y = pack_bf16x2(__uint_as_float(w[8 * cc + 2]) * inv, __uint_as_float(w[8 * cc + 3]) * inv);
                    o2.z = pack_bf16x2(__uint_as_float(w[8 * cc + 4]) * inv, __uint_as_float(w[8 * cc + 5]) * inv);
                    o2.w = pack_bf16x2(__uint_as_float(w[8 * cc + 6]) * inv, __uint_as_float(w[8 * cc + 7]) * inv);
                    *reinterpret_cast<uint4*>(dst + 32 + 8 * cc) = o2;
                }
            }
        }
        tc_fence_before();
        __syncthreads();   // P and the TMEM tiles are rewritten by the next query block
    }
    if (warp == 0) tmem_dealloc<512>(tmem);
}

// pixels [N,3,Hp*P,Wp*P] (bf16) -> patches [N*Hp*Wp, 3*P*P] in Conv2d weight order (c, py, px)
__global__ void __launch_bounds__(256) im2col_kernel(const uint16_t* px, uint16_t* out, int N, int Hp, int Wp, int P) {
    const int K = 3 * P * P;
    const size_t total = (size_t)N * Hp * Wp * K / 8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t e = i * 8;
        const int kk = (int)(e % K);
        const size_t patch = e / K;
        const int c = kk / (P * P), py = (kk / P) % P, x0 = kk % P;   // 8 consecutive px inside one patch row (P % 8 == 0)
        const int pw = (int)(patch % Wp), ph = (int)((patch / Wp) % Hp), n = (int)(patch / ((size_t)Wp * Hp));
        const size_t src = (((size_t)n * 3 + c) * (Hp * P) + ph * P + py) * (size_t)(Wp * P) + pw * P + x0;
        *reinterpret_cast<uint4*>(out + e) = __ldg(reinterpret_cast<const uint4*>(px + src));
    }
}

// h = x * sigmoid(g), each product rounded to bf16 like the eager graph (model.py:336-337)
__global__ void __launch_bounds__(256) sigmul_fwd_kernel(const uint16_t* x, const uint16_t* g, uint16_t* h, size_t n8) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 xv = __ldg(reinterpret_cast<const uint4*>(x) + i), gv = __ldg(reinterpret_cast<const uint4*>(g) + i);
        const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const float s0 = __bfloat162float(__float2bfloat16_rn(1.f / (1.f + __expf(-bf16lo_to_f32(gs[e])))));
            const float s1 = __bfloat162float(__float2bfloat16_rn(1.f / (1.f + __expf(-bf16hi_to_f32(gs[e])))));
            o[e] = pack_bf16x2(bf16lo_to_f32(xs[e]) * s0, bf16hi_to_f32(xs[e]) * s1);
        }
        reinterpret_cast<uint4*>(h)[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}
// dx = dh * s, dg = dh * x * s (1 - s), s = sigmoid(g)
__global__ void __launch_bounds__(256) sigmul_bwd_kernel(const uint16_t* x, const uint16_t* g, const uint16_t* dh, uint16_t* dx, uint16_t* dg, size_t n8) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 xv = __ldg(reinterpret_cast<const uint4*>(x) + i), gv = __ldg(reinterpret_cast<const uint4*>(g) + i),
                    dv = __ldg(reinterpret_cast<const uint4*>(dh) + i);
        const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w}, ds[4] = {dv.x, dv.y, dv.z, dv.w};
        uint32_t ox[4], og[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const float s0 = __bfloat162float(__float2bfloat16_rn(1.f / (1.f + __expf(-bf16lo_to_f32(gs[e])))));
            const float s1 = __bfloat162float(__float2bfloat16_rn(1.f / (1.f + __expf(-bf16hi_to_f32(gs[e])))));
            const float d0 = bf16lo_to_f32(ds[e]), d1 = bf16hi_to_f32(ds[e]);
            ox[e] = pack_bf16x2(d0 * s0, d1 * s1);
            const float t0 = __bfloat162float(__float2bfloat16_rn(d0 * bf16lo_to_f32(xs[e]))), t1 = __bfloat162float(__float2bfloat16_rn(d1 * bf16hi_to_f32(xs[e])));
            og[e] = pack_bf16x2(t0 * (1.f - s0) * s0, t1 * (1.f - s1) * s1);
        }
        if (dx) reinterpret_cast<uint4*>(dx)[i] = make_uint4(ox[0], ox[1], ox[2], ox[3]);
        reinterpret_cast<uint4*>(dg)[i] = make_uint4(og[0], og[1], og[2], og[3]);
    }
}

}  // namespace vrwkv

using namespace vrwkv;

extern "C" int vrwkv_vit_attention(int N, int S, int H, const uint16_t* q, const uint16_t* k, const uint16_t* v, uint16_t* o, void* stream) {
    if (N <= 0 || S <= 0 || H <= 0) return vrwkv_fail(VRWKV_EINVAL, "vit_attention: bad shape (%d,%d,%d)", N, S, H);
    if (S > 256) return vrwkv_fail(VRWKV_EUNSUP, "vit_attention: S=%d patches per image (single-tile kernel: S <= 256)", S);
    if (!q || !k || !v || !o) return vrwkv_fail(VRWKV_EINVAL, "vit_attention: null pointer");
    const int D = 64 * H;
    CUtensorMap tq, tk, tv;
    int rc;
    const void* in[3] = {q, k, v};
    CUtensorMap* tm[3] = {&tq, &tk, &tv};
    for (int i = 0; i < 3; i++)
        if ((rc = vrwkv_encode_2d(tm[i], in[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)D, (uint64_t)N * S, (uint64_t)D * 2, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B)))
            return rc;
    AttnArgs a{S, H, D, o, 0.125f};
    const size_t smem = sizeof(AttnSmem) + 1024;
    VRWKV_CUDA(cudaFuncSetAttribute(vit_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vit_attn_kernel<<<dim3(H, N), 128, smem, (cudaStream_t)stream>>>(tq, tk, tv, a);
    VRWKV_CUDA(cudaGetLastError());
    vrwkv_count_launch(1);
    return VRWKV_OK;
}

extern "C" int vrwkv_im2col_patches(int N, int Hp, int Wp, int P, const uint16_t* pixels, uint16_t* out, void* stream) {
    if (N <= 0 || Hp <= 0 || Wp <= 0 || P <= 0 || P % 8) return vrwkv_fail(VRWKV_EINVAL, "im2col: bad shape");
    if (!pixels || !out) return vrwkv_fail(VRWKV_EINVAL, "im2col: null pointer");
    im2col_kernel<<<148 * 4, 256, 0, (cudaStream_t)stream>>>(pixels, out, N, Hp, Wp, P);
    VRWKV_CUDA(cudaGetLastError());
    vrwkv_count_launch(1);
    return VRWKV_OK;
}

extern "C" int vrwkv_sigmul_forward(size_t n, const uint16_t* x, const uint16_t* g, uint16_t* h, void* stream) {
    if (!x || !g || !h || (n % 8)) return vrwkv_fail(VRWKV_EINVAL, "sigmul_forward: null pointer or n %% 8 != 0");
    sigmul_fwd_kernel<<<148 * 4, 256, 0, (cudaStream_t)stream>>>(x, g, h, n / 8);
    VRWKV_CUDA(cudaGetLastError());
    vrwkv_count_launch(1);
    return VRWKV_OK;
}

extern "C" int vrwkv_sigmul_backward(size_t n, const uint16_t* x, const uint16_t* g, const uint16_t* dh, uint16_t* dx, uint16_t* dg, void* stream) {
    if (!x || !g || !dh || !dg || (n % 8)) return vrwkv_fail(VRWKV_EINVAL, "sigmul_backward: null pointer or n %% 8 != 0");
    sigmul_bwd_kernel<<<148 * 4, 256, 0, (cudaStream_t)stream>>>(x, g, dh, dx, dg, n / 8);
    VRWKV_CUDA(cudaGetLastError());
    vrwkv_count_launch(1);
    return VRWKV_OK;
}
