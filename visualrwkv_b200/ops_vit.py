"""ViT forward for the SigLIP tower."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def siglip_forward(tower, pixels):
    c = tower.cfg
    vm = tower.vision_model
    D, nh, p = c["hidden"], c["heads"], c["patch"]
    hd = D // nh
    N = pixels.shape[0]
    # patch embedding as a GEMM over im2col'd 16x16x3 patches (Conv2d k=p, s=p)
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
