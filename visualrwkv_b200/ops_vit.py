"""ViT forward for the SigLIP tower: transformers' SiglipVisionModel arithmetic as called by
VisualRWKV-v7/v7.01/src/model.py:347-352,448-454, on our own kernels:
  patch embedding   im2col kernel + CTA-pair tcgen05 GEMM with the bias and the position table in the epilogue
  per layer         LayerNorm kernel -> q/k/v as ONE grouped GEMM launch (bias epilogue) -> tcgen05 attention kernel
                    -> out_proj GEMM (+bias +residual) -> LayerNorm -> fc1 GEMM (+bias, tanh-GELU) -> fc2 GEMM (+bias +residual)
  post_layernorm    LayerNorm kernel
Forward only (the tower is frozen, model.py:368-369).  Library fallbacks (cuBLAS / cuDNN SDPA) remain only for shapes the
kernels do not tile: more than 256 patches per image, or widths that are not multiples of 128."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib, fused


def _ln(x2, ln, eps):
    _, h, _ = fused.ln_mix_forward(x2, x2.shape[0], ln.weight, ln.bias, eps, [], want_h=True)
    return h


def _own(rows, D, mlp, S):
    return (rows % 8 == 0 and D % 128 == 0 and mlp % 128 == 0 and D % 64 == 0 and S <= 256 and fused.gemm2_supported(rows, D, D) and
            fused.gemm2_supported(rows, mlp, D) and fused.gemm2_supported(rows, D, mlp))


def siglip_forward(tower, pixels):
    c = tower.cfg
    vm = tower.vision_model
    D, nh, p = c["hidden"], c["heads"], c["patch"]
    N = pixels.shape[0]
    gh, gw = pixels.shape[2] // p, pixels.shape[3] // p
    S = gh * gw
    rows = N * S
    wpe = vm.embeddings.patch_embedding.weight.view(D, -1)
    if not (pixels.is_cuda and pixels.dtype == torch.bfloat16 and _own(rows, D, c["mlp"], S) and wpe.shape[1] % 64 == 0 and D // nh == 64 and p % 8 == 0):
        return _siglip_forward_library(tower, pixels)
    L = _lib.lib()
    st = _lib.cur_stream()
    px = pixels.contiguous()
    patches = torch.empty(rows, 3 * p * p, dtype=torch.bfloat16, device=px.device)
    fused._chk(L.vrwkv_im2col_patches(N, gh, gw, p, fused._p(px), fused._p(patches), st), "vrwkv_im2col_patches")
    # conv (k = s = patch) as a GEMM; bias and position embedding in the epilogue (the table row is the patch index)
    x = fused.gemm2(patches, wpe.contiguous(), fused.G2_TN, fused.EPI_BIAS_ADD, vm.embeddings.position_embedding.weight,
                    bias=vm.embeddings.patch_embedding.bias, r_rows=S)
    eps = c["eps"]
    for lyr in vm.encoder.layers:
        at = lyr.self_attn
        h = _ln(x, lyr.layer_norm1, eps)
        q, k, v = fused.gemm2_grouped([h, h, h], [at.q_proj.weight, at.k_proj.weight, at.v_proj.weight], fused.G2_TN, fused.EPI_BIAS,
                                      biases=[at.q_proj.bias, at.k_proj.bias, at.v_proj.bias])
        o = torch.empty_like(q)
        fused._chk(L.vrwkv_vit_attention(N, S, nh, fused._p(q), fused._p(k), fused._p(v), fused._p(o), st), "vrwkv_vit_attention")
        x = fused.gemm2(o, at.out_proj.weight, fused.G2_TN, fused.EPI_BIAS_ADD, x, bias=at.out_proj.bias)
        h = _ln(x, lyr.layer_norm2, eps)
        m = fused.gemm2(h, lyr.mlp.fc1.weight, fused.G2_TN, fused.EPI_BIAS_GELU, bias=lyr.mlp.fc1.bias)
        x = fused.gemm2(m, lyr.mlp.fc2.weight, fused.G2_TN, fused.EPI_BIAS_ADD, x, bias=lyr.mlp.fc2.bias)
    return _ln(x, vm.post_layernorm, eps).view(N, S, D)


def _siglip_forward_library(tower, pixels):
    c = tower.cfg
    vm = tower.vision_model
    D, nh, p = c["hidden"], c["heads"], c["patch"]
    hd = D // nh
    N = pixels.shape[0]
    gh, gw = pixels.shape[2] // p, pixels.shape[3] // p
    patches = pixels.view(N, 3, gh, p, gw, p).permute(0, 2, 4, 1, 3, 5).reshape(N * gh * gw, 3 * p * p)
    wpe = vm.embeddings.patch_embedding.weight.view(D, -1)
    x = torch.addmm(vm.embeddings.patch_embedding.bias, patches.to(wpe.dtype), wpe.t()).view(N, gh * gw, D)
    x = x + vm.embeddings.position_embedding.weight
    S = x.shape[1]
    for lyr in vm.encoder.layers:
        h = F.layer_norm(x, (D,), lyr.layer_norm1.weight, lyr.layer_norm1.bias, c["eps"])
        at = lyr.self_attn
        q = at.q_proj(h).view(N, S, nh, hd).transpose(1, 2)
        k = at.k_proj(h).view(N, S, nh, hd).transpose(1, 2)
        v = at.v_proj(h).view(N, S, nh, hd).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(N, S, D)
        x = x + at.out_proj(o)
        h = F.layer_norm(x, (D,), lyr.layer_norm2.weight, lyr.layer_norm2.bias, c["eps"])
        x = x + lyr.mlp.fc2(F.gelu(lyr.mlp.fc1(h), approximate="tanh"))
    return F.layer_norm(x, (D,), vm.post_layernorm.weight, vm.post_layernorm.bias, c["eps"])
