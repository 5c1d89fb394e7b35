// optim.cu — the optimizer step of the training loop as one multi-tensor kernel.
//
// The reference trains with DeepSpeed's FusedAdam (adam_w_mode) over bf16 parameters with fp32 master weights and moments
// (VisualRWKV-v7/v7.00/src/model.py:376-410, train.py:134).  Stock PyTorch needs three passes for that (bf16 grads -> fp32,
// fused AdamW on the fp32 copies, fp32 -> bf16 parameters: 40 bytes per parameter); here one pass reads the bf16 gradient
// and the fp32 master weight / moments and writes them back together with the bf16 parameter: 28 bytes per parameter.
// HBM-bound streaming: 16-byte accesses, one CTA per 8192-element chunk of one tensor (chunk table built by the host).
#include "common.cuh"
#include "host_util.h"

namespace vrwkv {

constexpr int AD_CHUNK = 8192;
constexpr int AD_THREADS = 256;

struct AdamTensor {
    uint16_t* p16;          // bf16 parameter (updated)
    const uint16_t* g16;    // bf16 gradient
    long long state_off;    // offset of this tensor in the flat fp32 master / exp_avg / exp_avg_sq buffers
    long long numel;
};

__global__ void __launch_bounds__(AD_THREADS) adamw_kernel(const AdamTensor* tens, const int2* chunks, float* master, float* m1, float* m2,
                                                           const float* lr_ptr, float lr, float b1, float b2, float eps, float wd,
                                                           const int* step_ptr, float grad_scale) {
    const int2 ck = chunks[blockIdx.x];
    const AdamTensor t = tens[ck.x];
    const long long begin = (long long)ck.y * AD_CHUNK;
    const long long end = begin + AD_CHUNK < t.numel ? begin + AD_CHUNK : t.numel;
    const float step = (float)*step_ptr;
    if (lr_ptr) lr = *lr_ptr;
    // torch.optim.AdamW: p *= 1 - lr wd;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;
    //                    p -= (lr / (1 - b1^t)) m / (sqrt(v) / sqrt(1 - b2^t) + eps)
    const float bc1 = 1.f - powf(b1, step), bc2 = 1.f - powf(b2, step);
    const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2), decay = 1.f - lr * wd;
    float* const pm = master + t.state_off;
    float* const pa = m1 + t.state_off;
    float* const pv = m2 + t.state_off;
    const bool vec = ((t.state_off | (long long)(uintptr_t)t.p16 | (long long)(uintptr_t)t.g16) & 15) == 0 && (t.state_off & 3) == 0;
    auto upd = [&](float g, float& p, float& a, float& v) {
        g *= grad_scale;
        p *= decay;
        a = b1 * a + (1.f - b1) * g;
        v = b2 * v + (1.f - b2) * g * g;
        p -= step_size * a / (sqrtf(v) * inv_sqrt_bc2 + eps);
    };
    if (vec) {
        for (long long i = begin + (long long)threadIdx.x * 8; i + 8 <= end; i += AD_THREADS * 8) {
            const uint4 gu = *reinterpret_cast<const uint4*>(t.g16 + i);
            float4 p0 = *reinterpret_cast<float4*>(pm + i), p1 = *reinterpret_cast<float4*>(pm + i + 4);
            float4 a0 = *reinterpret_cast<float4*>(pa + i), a1 = *reinterpret_cast<float4*>(pa + i + 4);
            float4 v0 = *reinterpret_cast<float4*>(pv + i), v1 = *reinterpret_cast<float4*>(pv + i + 4);
            upd(bf16lo_to_f32(gu.x), p0.x, a0.x, v0.x); upd(bf16hi_to_f32(gu.x), p0.y, a0.y, v0.y);
            upd(bf16lo_to_f32(gu.y), p0.z, a0.z, v0.z); upd(bf16hi_to_f32(gu.y), p0.w, a0.w, v0.w);
            upd(bf16lo_to_f32(gu.z), p1.x, a1.x, v1.x); upd(bf16hi_to_f32(gu.z), p1.y, a1.y, v1.y);
            upd(bf16lo_to_f32(gu.w), p1.z, a1.z, v1.z); upd(bf16hi_to_f32(gu.w), p1.w, a1.w, v1.w);
            *reinterpret_cast<float4*>(pm + i) = p0; *reinterpret_cast<float4*>(pm + i + 4) = p1;
            *reinterpret_cast<float4*>(pa + i) = a0; *reinterpret_cast<float4*>(pa + i + 4) = a1;
            *reinterpret_cast<float4*>(pv + i) = v0; *reinterpret_cast<float4*>(pv + i + 4) = v1;
            uint4 o;
            o.x = pack_bf16x2(p0.x, p0.y); o.y = pack_bf16x2(p0.z, p0.w); o.z = pack_bf16x2(p1.x, p1.y); o.w = pack_bf16x2(p1.z, p1.w);
            *reinterpret_cast<uint4*>(t.p16 + i) = o;
        }
    }
    // tail of the chunk (numel % 8), or the whole chunk of an unaligned tensor
    const long long tail0 = vec ? begin + ((end - begin) / 8) * 8 : begin;
    for (long long i = tail0 + threadIdx.x; i < end; i += AD_THREADS) {
        float p = pm[i], a = pa[i], v = pv[i];
        upd(__uint_as_float((uint32_t)t.g16[i] << 16), p, a, v);
        pm[i] = p; pa[i] = a; pv[i] = v;
        t.p16[i] = f32_to_bf16_bits(p);
    }
}

__global__ void adam_step_inc_kernel(int* step) { *step += 1; }

}  // namespace vrwkv

using namespace vrwkv;

extern "C" int vrwkv_adamw_chunk() { return AD_CHUNK; }

/* tensors: device array of ntensors records {bf16* param, const bf16* grad, int64 state_off, int64 numel} (32 bytes each);
 * chunks: device array of nchunks int2 {tensor, chunk index}; master / exp_avg / exp_avg_sq: flat fp32 state; step: device
 * int, incremented here before the update (so the whole call is CUDA-graph capturable); lr_dev: optional device float. */
extern "C" int vrwkv_adamw_step(int nchunks, const void* tensors, const void* chunks, float* master, float* exp_avg, float* exp_avg_sq,
                                int* step, const float* lr_dev, float lr, float beta1, float beta2, float eps, float weight_decay,
                                float grad_scale, void* stream) {
    if (nchunks <= 0 || !tensors || !chunks || !master || !exp_avg || !exp_avg_sq || !step)
        return vrwkv_fail(VRWKV_EINVAL, "adamw_step: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    adam_step_inc_kernel<<<1, 1, 0, st>>>(step);
    adamw_kernel<<<nchunks, AD_THREADS, 0, st>>>(reinterpret_cast<const AdamTensor*>(tensors), reinterpret_cast<const int2*>(chunks), master,
                                                  exp_avg, exp_avg_sq, lr_dev, lr, beta1, beta2, eps, weight_decay, step, grad_scale);
    VRWKV_CUDA(cudaGetLastError());
    vrwkv_count_launch(2);
    return VRWKV_OK;
}
