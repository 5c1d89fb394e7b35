// torch_shim.cpp — registers the reference's operator surface on top of the C ABI.
//
// Library name and schema strings are VERBATIM those of the reference binding
// (VisualRWKV-v7/v7.00/cuda/wkv7_op.cpp:21-29) so `torch.ops.wind_backstepping.forward/backward`
// keep working for WindBackstepping (v7.00/src/model.py:45-65) and every v7.xx fork.
// Schema naming trap: schema (z, a) == kernel (a, b) == python (-kk, kk*a).
//
// Unlike the reference binding this one validates its arguments, switches to the tensors' device
// and launches on torch's CURRENT stream (the reference uses the legacy default stream,
// wkv7_cuda.cu:133,137).
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include "../../include/vrwkv_b200.h"

namespace {

inline const uint16_t* bfp(const at::Tensor& t) { return reinterpret_cast<const uint16_t*>(t.data_ptr()); }
inline uint16_t* bfp_mut(at::Tensor& t) { return reinterpret_cast<uint16_t*>(t.data_ptr()); }

void check_bf16_stream(const at::Tensor& t, const at::Tensor& like, const char* name) {
    TORCH_CHECK(t.is_cuda(), "wind_backstepping: ", name, " must be a CUDA tensor");
    TORCH_CHECK(t.scalar_type() == at::kBFloat16, "wind_backstepping: ", name, " must be bfloat16");
    TORCH_CHECK(t.is_contiguous(), "wind_backstepping: ", name, " must be contiguous");
    TORCH_CHECK(t.sizes() == like.sizes(), "wind_backstepping: ", name, " shape mismatch");
    TORCH_CHECK(t.device() == like.device(), "wind_backstepping: ", name, " on a different device");
}

void check_f32(const at::Tensor& t, const at::Tensor& like, int64_t numel, const char* name) {
    TORCH_CHECK(t.is_cuda() && t.device() == like.device(), "wind_backstepping: ", name, " device mismatch");
    TORCH_CHECK(t.scalar_type() == at::kFloat, "wind_backstepping: ", name, " must be float32");
    TORCH_CHECK(t.is_contiguous() && t.numel() == numel, "wind_backstepping: ", name, " has the wrong size/layout");
}

void forward(at::Tensor& w, at::Tensor& q, at::Tensor& k, at::Tensor& v, at::Tensor& z, at::Tensor& a, at::Tensor& y,
             at::Tensor& s, at::Tensor& sa) {
    TORCH_CHECK(w.dim() == 4, "wind_backstepping.forward: expected [B,T,H,64] tensors");
    const int64_t B = w.size(0), T = w.size(1), H = w.size(2), C = w.size(3);
    TORCH_CHECK(C == VRWKV_HEAD_SIZE, "wind_backstepping.forward: head size must be 64, got ", C);
    check_bf16_stream(w, w, "w"); check_bf16_stream(q, w, "q"); check_bf16_stream(k, w, "k");
    check_bf16_stream(v, w, "v"); check_bf16_stream(z, w, "z"); check_bf16_stream(a, w, "a");
    check_bf16_stream(y, w, "y");
    TORCH_CHECK(T % VRWKV_CHUNK_LEN == 0, "wind_backstepping.forward: T must be a multiple of 16, got ", T);
    check_f32(s, w, B * H * (T / VRWKV_CHUNK_LEN) * C * C, "s");
    check_f32(sa, w, B * T * H * C, "sa");
    c10::cuda::CUDAGuard guard(w.device());
    auto st = c10::cuda::getCurrentCUDAStream();
    int rc = vrwkv_wkv7_forward((int)B, (int)T, (int)H, bfp(w), bfp(q), bfp(k), bfp(v), bfp(z), bfp(a), bfp_mut(y),
                                s.data_ptr<float>(), sa.data_ptr<float>(), st.stream());
    TORCH_CHECK(rc == 0, "vrwkv_wkv7_forward failed (", rc, "): ", vrwkv_last_error());
}

void backward(at::Tensor& w, at::Tensor& q, at::Tensor& k, at::Tensor& v, at::Tensor& z, at::Tensor& a, at::Tensor& dy,
              at::Tensor& s, at::Tensor& sa, at::Tensor& dw, at::Tensor& dq, at::Tensor& dk, at::Tensor& dv,
              at::Tensor& dz, at::Tensor& da) {
    TORCH_CHECK(w.dim() == 4, "wind_backstepping.backward: expected [B,T,H,64] tensors");
    const int64_t B = w.size(0), T = w.size(1), H = w.size(2), C = w.size(3);
    TORCH_CHECK(C == VRWKV_HEAD_SIZE, "wind_backstepping.backward: head size must be 64, got ", C);
    check_bf16_stream(w, w, "w"); check_bf16_stream(q, w, "q"); check_bf16_stream(k, w, "k");
    check_bf16_stream(v, w, "v"); check_bf16_stream(z, w, "z"); check_bf16_stream(a, w, "a");
    check_bf16_stream(dy, w, "dy"); check_bf16_stream(dw, w, "dw"); check_bf16_stream(dq, w, "dq");
    check_bf16_stream(dk, w, "dk"); check_bf16_stream(dv, w, "dv"); check_bf16_stream(dz, w, "dz");
    check_bf16_stream(da, w, "da");
    TORCH_CHECK(T % VRWKV_CHUNK_LEN == 0, "wind_backstepping.backward: T must be a multiple of 16, got ", T);
    check_f32(s, w, B * H * (T / VRWKV_CHUNK_LEN) * C * C, "s");
    check_f32(sa, w, B * T * H * C, "sa");
    c10::cuda::CUDAGuard guard(w.device());
    auto st = c10::cuda::getCurrentCUDAStream();
    int rc = vrwkv_wkv7_backward((int)B, (int)T, (int)H, bfp(w), bfp(q), bfp(k), bfp(v), bfp(z), bfp(a), bfp(dy),
                                 s.data_ptr<float>(), sa.data_ptr<float>(), bfp_mut(dw), bfp_mut(dq), bfp_mut(dk),
                                 bfp_mut(dv), bfp_mut(dz), bfp_mut(da), st.stream());
    TORCH_CHECK(rc == 0, "vrwkv_wkv7_backward failed (", rc, "): ", vrwkv_last_error());
}

}  // namespace

TORCH_LIBRARY(wind_backstepping, m) {
    m.def("forward(Tensor w, Tensor q, Tensor k, Tensor v, Tensor z, Tensor a, Tensor(a!) y, Tensor(b!) s, Tensor(c!) sa) -> ()");
    m.def("backward(Tensor w, Tensor q, Tensor k, Tensor v, Tensor z, Tensor a, Tensor dy, Tensor s, Tensor sa, Tensor(a!) dw, Tensor(b!) dq, Tensor(c!) dk, Tensor(d!) dv, Tensor(e!) dz, Tensor(f!) da) -> ()");
}

TORCH_LIBRARY_IMPL(wind_backstepping, CUDA, m) {
    m.impl("forward", &forward);
    m.impl("backward", &backward);
}
