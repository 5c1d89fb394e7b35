"""Operator layer between the reference-shaped modules (model.py) and the native kernels.

Each function cites the reference lines it implements.  GEMMs that are plain library GEMMs go to
cuBLAS through torch.matmul; everything else on the path is a hand-written sm_100a kernel reached
through the C ABI (fused.py) — there is no CPU path here.
"""
from __future__ import annotations

import warnings

import torch
import torch.nn.functional as F

from .wkv7 import RUN_CUDA_RWKV7g


def _shift(x):
    """nn.ZeroPad2d((0,0,1,-1)) (model.py:149)."""
    return torch.cat([torch.zeros_like(x[:, :1]), x[:, :-1]], dim=1)


def layer_norm(x, ln):
    return F.layer_norm(x, (x.shape[-1],), ln.weight, ln.bias, ln.eps)


def tmix_forward(m, x, v_first):
    """RWKV_Tmix_x070.forward (model.py:163-195)."""
    B, T, C = x.size()
    H = m.n_head
    xx = _shift(x) - x
    xr = x + xx * m.x_r
    xw = x + xx * m.x_w
    xk = x + xx * m.x_k
    xv = x + xx * m.x_v
    xa = x + xx * m.x_a
    xg = x + xx * m.x_g
    r = m.receptance(xr)
    w = -F.softplus(-(m.w0 + torch.tanh(xw @ m.w1) @ m.w2)) - 0.5
    k = m.key(xk)
    v = m.value(xv)
    if m.layer_id == 0:
        v_first = v
    else:
        v = v + (v_first - v) * torch.sigmoid(m.v0 + (xv @ m.v1) @ m.v2)
    a = torch.sigmoid(m.a0 + (xa @ m.a1) @ m.a2)
    g = torch.sigmoid(xg @ m.g1) @ m.g2
    kk = F.normalize((k * m.k_k).view(B, T, H, -1), dim=-1, p=2.0).view(B, T, C)
    k = k * (1 + (a - 1) * m.k_a)
    y = RUN_CUDA_RWKV7g(r.contiguous(), w.contiguous(), k.contiguous(), v.contiguous(), (-kk).contiguous(),
                        (kk * a).contiguous())
    y = F.group_norm(y.view(B * T, C), H, m.ln_x.weight, m.ln_x.bias, m.ln_x.eps).view(B, T, C)
    y = y + ((r.view(B, T, H, -1) * k.view(B, T, H, -1) * m.r_k).sum(dim=-1, keepdim=True) * v.view(B, T, H, -1)).view(B, T, C)
    return m.output(y * g), v_first


def cmix_forward(m, x):
    """RWKV_CMix_x070.forward (model.py:221-227)."""
    xx = _shift(x) - x
    k = x + xx * m.x_k
    k = torch.relu(m.key(k)) ** 2
    return m.value(k)


def block_forward(blk, x, v_first):
    """Block.forward (model.py:247-254)."""
    if blk.layer_id == 0:
        x = layer_norm(x, blk.ln0)
    xx, v_first = blk.att(layer_norm(x, blk.ln1), v_first)
    x = x + xx
    x = x + blk.ffn(layer_norm(x, blk.ln2))
    return x, v_first


def projector_forward(m, x):
    """MLPWithContextGating.forward (model.py:335-338)."""
    gating = torch.sigmoid(m.gate(x))
    return layer_norm(m.o_proj(x * gating), m.ln_v)


def adaptive_pooling(feats, out_hw):
    """VisualRWKV.adaptive_pooling (model.py:442-447): [N,L,D] -> [N,out_hw^2,D]."""
    N, L, D = feats.shape
    hw = int(L ** 0.5)
    if hw == out_hw:
        return feats
    f = feats.view(N, hw, hw, D).permute(0, 3, 1, 2)
    f = F.adaptive_avg_pool2d(f, out_hw)
    return f.reshape(N, D, -1).permute(0, 2, 1).contiguous()


def embed_and_scatter(emb_weight, input_ids, image_features, image_token_index, sample_ids=None):
    """preparing_embedding (model.py:481-493): emb(input_ids) with the rows where ids == 65535 replaced by
    the image features in row-major order of appearance (bit-exact copy)."""
    B, L = input_ids.shape
    D = emb_weight.shape[1]
    x = F.embedding(input_ids, emb_weight).view(B * L, D)
    sel = input_ids.view(-1) == image_token_index
    feats = image_features.reshape(-1, D)
    n_sel = int(sel.sum())
    if n_sel != feats.shape[0]:
        warnings.warn(f"sample_id: {':::'.join(sample_ids or [])}, image tokens: {n_sel}, but image features: {feats.shape[0]}")
        feats = feats[:n_sel]
    x = x.clone() if not x.requires_grad and not feats.requires_grad else x
    x = x.masked_scatter(sel.unsqueeze(-1), feats.to(x.dtype)) if feats.requires_grad or x.requires_grad else _scatter_inplace(x, sel, feats)
    return x.view(B, L, D)


def _scatter_inplace(x, sel, feats):
    x[sel] = feats.to(x.dtype)
    return x


def training_loss(logits, targets, ignore_index, l2wrap):
    """VisualRWKV.training_step (model.py:418-434)."""
    shift_logits = logits[..., :-1, :].contiguous()
    shift_labels = targets[..., 1:].contiguous()
    valid = (shift_labels != ignore_index).sum(1).clamp(min=1)
    loss = F.cross_entropy(shift_logits.view(-1, shift_logits.size(-1)), shift_labels.view(-1),
                           ignore_index=ignore_index, reduction="none")
    loss = (loss.view(shift_labels.size()).sum(1) / valid).mean()
    return l2wrap.apply(loss, logits)
