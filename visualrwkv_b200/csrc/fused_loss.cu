// fused_loss.cu — shifted cross-entropy + L2Wrap over the [B,T,V] logits in two passes (one forward, one
// backward), replacing the reference's ~10 passes over the 2.1 GB logits tensor
// (VisualRWKV-v7/v7.00/src/model.py:418-434 training_step, :257-271 L2Wrap):
//
//   forward : per row (b,t): m = max_v x, am = argmax_v x, lse = m + log sum_v exp(x - m),
//             nll = lse - x[target]  with target = labels[b,t+1]  (no target for t = T-1 or label == ignore)
//   backward: dx[v] = wrow * (exp(x[v] - lse) - [v == target])          (rows with a target; wrow = g/(valid_b*B))
//                   + l2 * m * [v == am]                                 (every row; l2 = 1e-4/(B*T), L2Wrap)
//             written over the logits buffer in place (bf16).
//
// One CTA (256 threads) per row, 128-bit loads, fp32 math, online max/sum.  HBM bytes: forward 2 B/logit,
// backward 4 B/logit.
#include "host_util.h"
#include "rowops.cuh"

namespace vrwkv {

struct CeArgs {
    int rows, T, V, ignore_index;
    uint16_t* logits;        // [rows, V]; overwritten by the backward
    const long long* labels; // [rows] (row-major [B,T]); target of row (b,t) is labels[b,t+1]
    float *lse, *rowmax, *nll;
    int* argmax;
    const float* wrow;       // backward: [rows] weight of the CE term (0 for rows without a target)
    float l2;
};

// exp(m - nm) for the online-softmax merges; a partial that saw no element (m = -inf, s = 0) contributes exactly 0
// (exp(-inf - -inf) would be NaN: rows shorter than 2048 leave whole lanes empty)
__device__ __forceinline__ float ce_rescale(float m, float nm) { return m == -INFINITY ? 0.f : __expf(m - nm); }

__global__ void __launch_bounds__(256) ce_fwd_kernel(const CeArgs a) {
    __shared__ float smax[8], ssum[8];
    __shared__ int sidx[8];
    const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint16_t* x = a.logits + (size_t)row * a.V;
    float m = -INFINITY, s = 0.f;
    int am = 0;
    for (int i = tid * 8; i < a.V; i += 256 * 8) {
        const F8 v = ld_bf16x8(x + i);
        float vm = v.v[0];
        int vi = 0;
#pragma unroll
        for (int e = 1; e < 8; e++)
            if (v.v[e] > vm) { vm = v.v[e]; vi = e; }
        if (vm > m) {
            s *= ce_rescale(m, vm);
            m = vm;
            am = i + vi;
        }
#pragma unroll
        for (int e = 0; e < 8; e++) s += __expf(v.v[e] - m);
    }
    // warp then block reduction of (m, s, am); ties keep the smaller index
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, o), os = __shfl_xor_sync(0xffffffffu, s, o);
        const int oi = __shfl_xor_sync(0xffffffffu, am, o);
        const float nm = fmaxf(m, om);
        s = s * ce_rescale(m, nm) + os * ce_rescale(om, nm);
        if (om > m || (om == m && oi < am)) am = oi;
        m = nm;
    }
    if (lane == 0) { smax[warp] = m; ssum[warp] = s; sidx[warp] = am; }
    __syncthreads();
    if (tid == 0) {
        float M = smax[0], S = ssum[0];
        int A = sidx[0];
        for (int w = 1; w < 8; w++) {
            const float nm = fmaxf(M, smax[w]);
            S = S * ce_rescale(M, nm) + ssum[w] * ce_rescale(smax[w], nm);
            if (smax[w] > M || (smax[w] == M && sidx[w] < A)) A = sidx[w];
            M = nm;
        }
        const float lse = M + __logf(S);
        a.lse[row] = lse;
        a.rowmax[row] = M;
        a.argmax[row] = A;
        const int t = row % a.T;
        float nll = 0.f;
        if (t + 1 < a.T) {
            const long long tgt = a.labels[row + 1];
            if (tgt != a.ignore_index) nll = lse - bf16lo_to_f32((uint32_t)x[tgt]);
        }
        a.nll[row] = nll;
    }
}

__global__ void __launch_bounds__(256) ce_bwd_kernel(const CeArgs a) {
    const int row = blockIdx.x, tid = threadIdx.x;
    uint16_t* x = a.logits + (size_t)row * a.V;
    const float lse = a.lse[row], w = a.wrow[row];
    const int am = a.argmax[row];
    const float l2v = a.l2 * a.rowmax[row];
    const int t = row % a.T;
    long long tgt = -1;
    if (t + 1 < a.T) {
        tgt = a.labels[row + 1];
        if (tgt == a.ignore_index) tgt = -1;
    }
    for (int i = tid * 8; i < a.V; i += 256 * 8) {
        F8 v = ld_bf16x8(x + i);
#pragma unroll
        for (int e = 0; e < 8; e++) {
            float g = w != 0.f ? w * __expf(v.v[e] - lse) : 0.f;
            if (i + e == tgt) g -= w;
            // the eager graph adds two bf16 tensors (CE grad + L2Wrap scatter): round the CE term first
            g = rb(g);
            if (i + e == am) g += rb(l2v);
            v.v[e] = g;
        }
        st_bf16x8(x + i, v);
    }
}

}  // namespace vrwkv

using namespace vrwkv;

extern "C" int vrwkv_ce_forward(int rows, int T, int V, int ignore_index, const uint16_t* logits, const long long* labels,
                                float* lse, float* rowmax, int* argmax, float* nll, void* stream) {
    if (rows <= 0 || T <= 0 || rows % T || V <= 0 || V % 8) return vrwkv_fail(VRWKV_EINVAL, "ce_forward: bad shape (%d,%d,%d)", rows, T, V);
    if (!logits || !labels || !lse || !rowmax || !argmax || !nll) return vrwkv_fail(VRWKV_EINVAL, "ce_forward: null pointer");
    CeArgs a{};
    a.rows = rows; a.T = T; a.V = V; a.ignore_index = ignore_index; a.logits = const_cast<uint16_t*>(logits); a.labels = labels;
    a.lse = lse; a.rowmax = rowmax; a.argmax = argmax; a.nll = nll;
    ce_fwd_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(a);
    VRWKV_CUDA(cudaGetLastError());
    vrwkv_count_launch(1);
    return VRWKV_OK;
}

extern "C" int vrwkv_ce_backward(int rows, int T, int V, int ignore_index, uint16_t* logits_inout, const long long* labels,
                                 const float* lse, const float* rowmax, const int* argmax, const float* wrow, float l2,
                                 void* stream) {
    if (rows <= 0 || T <= 0 || rows % T || V <= 0 || V % 8) return vrwkv_fail(VRWKV_EINVAL, "ce_backward: bad shape (%d,%d,%d)", rows, T, V);
    if (!logits_inout || !labels || !lse || !rowmax || !argmax || !wrow) return vrwkv_fail(VRWKV_EINVAL, "ce_backward: null pointer");
    CeArgs a{};
    a.rows = rows; a.T = T; a.V = V; a.ignore_index = ignore_index; a.logits = logits_inout; a.labels = labels;
    a.lse = const_cast<float*>(lse); a.rowmax = const_cast<float*>(rowmax); a.argmax = const_cast<int*>(argmax);
    a.wrow = wrow; a.l2 = l2;
    ce_bwd_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(a);
    VRWKV_CUDA(cudaGetLastError());
    vrwkv_count_launch(1);
    return VRWKV_OK;
}
