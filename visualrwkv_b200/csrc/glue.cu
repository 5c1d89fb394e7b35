// glue.cu — the index / gather work around the towers, one kernel each instead of strings of eager ops:
//   embed_scatter : VisualRWKV.preparing_embedding (VisualRWKV-v7/v7.00/src/model.py:473-494): emb(input_ids) with the rows where
//                   ids == image_token_index replaced by the image features in row-major order of appearance (bit-exact row
//                   copies; the k-th selected slot takes feature row k), and its backward (gather of the output gradient rows
//                   back into feature order; the embedding table is frozen, v7.00/train.py:196);
//   adaptive_pool : VisualRWKV.adaptive_pooling (model.py:442-447): [N, hw*hw, D] -> AdaptiveAvgPool2d(out) -> [N, out*out, D]
//                   (fp32 window sums, one rounding to bf16 — what torch's kernel does on a bf16 input).
// HBM-bound row copies: 16-byte accesses, one warp per token row.
#include "common.cuh"
#include "host_util.h"

namespace vrwkv {

constexpr int ES_TOK = 64;       // tokens per CTA
constexpr int ES_THREADS = 256;

// rank of every image-token slot of this CTA's 64 tokens (row-major order over the whole batch) -> srank[t] (or -1)
__device__ __forceinline__ void es_ranks(const long long* ids, int ntok, int start, long long img_id, int* srank, int* total_out) {
    __shared__ int red[ES_THREADS / 32];
    __shared__ int base_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int cnt = 0;
    for (int i = tid; i < start; i += ES_THREADS) cnt += (ids[i] == img_id);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0) red[warp] = cnt;
    __syncthreads();
    if (tid == 0) {
        int b = 0;
        for (int w = 0; w < ES_THREADS / 32; w++) b += red[w];
        base_s = b;
    }
    __syncthreads();
    if (warp < 2) {   // 64 tokens: two warps, ballot prefix
        const int t = start + tid;
        const bool f = t < ntok && ids[t] == img_id;
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (lane == 0) red[warp] = __popc(m);
        __syncwarp();
        const int local = __popc(m & ((1u << lane) - 1u));
        srank[tid] = f ? local : -1;
    }
    __syncthreads();
    if (tid < ES_TOK && srank[tid] >= 0) srank[tid] += base_s + (tid >= 32 ? red[0] : 0);
    if (tid == 0 && total_out && start + ES_TOK >= ntok) *total_out = base_s + red[0] + red[1];   // the last CTA knows the grand total
    __syncthreads();
}

__global__ void __launch_bounds__(ES_THREADS) embed_scatter_fwd_kernel(const long long* ids, int ntok, const uint16_t* emb, const uint16_t* feats,
                                                                        int nfeat, int D, long long img_id, uint16_t* out, int* total_out) {
    __shared__ int srank[ES_TOK];
    const int start = blockIdx.x * ES_TOK;
    es_ranks(ids, ntok, start, img_id, srank, total_out);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int tt = warp; tt < ES_TOK; tt += ES_THREADS / 32) {
        const int t = start + tt;
        if (t >= ntok) break;
        const int rk = srank[tt];
        // surplus slots (more image tokens than feature rows: the reference raises, model.py:487-493; here the host is told
        // through total_out) keep the embedding row of the image-token id
        const uint16_t* src = (rk >= 0 && rk < nfeat) ? feats + (size_t)rk * D : emb + (size_t)ids[t] * D;
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        uint4* d4 = reinterpret_cast<uint4*>(out + (size_t)t * D);
        for (int i = lane; i < D / 8; i += 32) d4[i] = __ldg(s4 + i);
    }
}

// d_feats[k] = dout[slot of the k-th image token]; rows of d_feats beyond the number of slots must have been zeroed
__global__ void __launch_bounds__(ES_THREADS) embed_scatter_bwd_kernel(const long long* ids, int ntok, const uint16_t* dout, int nfeat, int D,
                                                                        long long img_id, uint16_t* dfeats) {
    __shared__ int srank[ES_TOK];
    const int start = blockIdx.x * ES_TOK;
    es_ranks(ids, ntok, start, img_id, srank, nullptr);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int tt = warp; tt < ES_TOK; tt += ES_THREADS / 32) {
        const int t = start + tt;
        if (t >= ntok) break;
        const int rk = srank[tt];
        if (rk < 0 || rk >= nfeat) continue;
        const uint4* s4 = reinterpret_cast<const uint4*>(dout + (size_t)t * D);
        uint4* d4 = reinterpret_cast<uint4*>(dfeats + (size_t)rk * D);
        for (int i = lane; i < D / 8; i += 32) d4[i] = __ldg(s4 + i);
    }
}

// one thread per (image, output cell, 8 channels)
__global__ void __launch_bounds__(256) adaptive_pool_kernel(const uint16_t* x, uint16_t* y, int N, int hw, int out, int D) {
    const int d8 = D / 8;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)N * out * out * d8;
    if (idx >= total) return;
    const int c8 = (int)(idx % d8);
    const int ox = (int)((idx / d8) % out), oy = (int)((idx / ((size_t)d8 * out)) % out), n = (int)(idx / ((size_t)d8 * out * out));
    // AdaptiveAvgPool2d windows: [floor(i*hw/out), ceil((i+1)*hw/out))
    const int y0 = (oy * hw) / out, y1 = ((oy + 1) * hw + out - 1) / out;
    const int x0 = (ox * hw) / out, x1 = ((ox + 1) * hw + out - 1) / out;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int yy = y0; yy < y1; yy++)
        for (int xx = x0; xx < x1; xx++) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + (((size_t)n * hw + yy) * hw + xx) * D) + c8);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
                acc[2 * e] += bf16lo_to_f32(w[e]);
                acc[2 * e + 1] += bf16hi_to_f32(w[e]);
            }
        }
    const float inv = 1.f / (float)((y1 - y0) * (x1 - x0));
    uint4 o;
    o.x = pack_bf16x2(acc[0] * inv, acc[1] * inv);
    o.y = pack_bf16x2(acc[2] * inv, acc[3] * inv);
    o.z = pack_bf16x2(acc[4] * inv, acc[5] * inv);
    o.w = pack_bf16x2(acc[6] * inv, acc[7] * inv);
    *(reinterpret_cast<uint4*>(y + (((size_t)n * out + oy) * out + ox) * D) + c8) = o;
}

}  // namespace vrwkv

using namespace vrwkv;

extern "C" int vrwkv_embed_scatter_forward(int ntok, int D, int nfeat, long long image_token_index, const long long* ids, const uint16_t* emb,
                                           const uint16_t* feats, uint16_t* out, int* n_slots_out, void* stream) {
    if (ntok <= 0 || D <= 0 || D % 8 || nfeat < 0) return vrwkv_fail(VRWKV_EINVAL, "embed_scatter: bad shape (ntok %d, D %d, nfeat %d)", ntok, D, nfeat);
    if (!ids || !emb || !out || (nfeat > 0 && !feats)) return vrwkv_fail(VRWKV_EINVAL, "embed_scatter: null pointer");
    embed_scatter_fwd_kernel<<<(ntok + ES_TOK - 1) / ES_TOK, ES_THREADS, 0, (cudaStream_t)stream>>>(ids, ntok, emb, feats, nfeat, D, image_token_index, out,
                                                                                                  n_slots_out);
    VRWKV_CUDA(cudaGetLastError());
    vrwkv_count_launch(1);
    return VRWKV_OK;
}

extern "C" int vrwkv_embed_scatter_backward(int ntok, int D, int nfeat, long long image_token_index, const long long* ids, const uint16_t* dout,
                                            uint16_t* dfeats, void* stream) {
    if (ntok <= 0 || D <= 0 || D % 8 || nfeat <= 0) return vrwkv_fail(VRWKV_EINVAL, "embed_scatter_backward: bad shape");
    if (!ids || !dout || !dfeats) return vrwkv_fail(VRWKV_EINVAL, "embed_scatter_backward: null pointer");
    VRWKV_CUDA(cudaMemsetAsync(dfeats, 0, (size_t)nfeat * D * 2, (cudaStream_t)stream));
    embed_scatter_bwd_kernel<<<(ntok + ES_TOK - 1) / ES_TOK, ES_THREADS, 0, (cudaStream_t)stream>>>(ids, ntok, dout, nfeat, D, image_token_index, dfeats);
    VRWKV_CUDA(cudaGetLastError());
    vrwkv_count_launch(1);
    return VRWKV_OK;
}

extern "C" int vrwkv_adaptive_pool(int N, int hw, int out, int D, const uint16_t* x, uint16_t* y, void* stream) {
    if (N <= 0 || hw <= 0 || out <= 0 || D <= 0 || D % 8) return vrwkv_fail(VRWKV_EINVAL, "adaptive_pool: bad shape");
    if (!x || !y) return vrwkv_fail(VRWKV_EINVAL, "adaptive_pool: null pointer");
    const size_t total = (size_t)N * out * out * (D / 8);
    adaptive_pool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, y, N, hw, out, D);
    VRWKV_CUDA(cudaGetLastError());
    vrwkv_count_launch(1);
    return VRWKV_OK;
}
