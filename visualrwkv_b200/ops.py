"""Operator layer between the reference-shaped modules (model.py) and the native kernels.

Each function cites the reference lines it implements.  Plain GEMMs go to cuBLAS through torch.matmul;
everything else on the path is a hand-written sm_100a kernel reached through the C ABI (fused.py, wkv7.py).
There is no CPU / eager fallback here: CUDA bf16 tensors are required and the native library must load.
"""
from __future__ import annotations

import warnings

import torch
import torch.nn.functional as F

from . import fused


def _need_cuda_bf16(x, what):
    if not (x.is_cuda and x.dtype == torch.bfloat16):
        raise RuntimeError(f"{what}: the sm_100a path needs CUDA bf16 tensors (got {x.device}, {x.dtype}); "
                           "there is no CPU / fp32 fallback")


def layer_norm(x, ln):
    _need_cuda_bf16(x, "layer_norm")
    return fused.LayerNormFn.apply(x, ln.weight, ln.bias, ln.eps)


def _tmix_args(m):
    has = m.layer_id != 0
    return (m.x_r, m.x_w, m.x_k, m.x_v, m.x_a, m.x_g, m.w0, m.w1, m.w2, m.a0, m.a1, m.a2,
            m.v0 if has else None, m.v1 if has else None, m.v2 if has else None, m.g1, m.g2, m.k_k, m.k_a, m.r_k,
            m.receptance.weight, m.key.weight, m.value.weight, m.output.weight, m.ln_x.weight, m.ln_x.bias)


def tmix_forward(m, x, v_first):
    """RWKV_Tmix_x070.forward (model.py:163-195) on an already layer-normed x: returns (att_out, v_first)."""
    _need_cuda_bf16(x, "RWKV_Tmix_x070")
    out, v = fused.TmixBlockFn.apply(x, v_first if m.layer_id != 0 else None, None, None, *_tmix_args(m),
                                     m.layer_id, m.n_head, 1e-5, m.ln_x.eps, False)
    return out, (v if m.layer_id == 0 else v_first)


def cmix_forward(m, x):
    """RWKV_CMix_x070.forward (model.py:221-227) on an already layer-normed x."""
    _need_cuda_bf16(x, "RWKV_CMix_x070")
    return fused.CmixBlockFn.apply(x, None, None, m.x_k, m.key.weight, m.value.weight, 1e-5, False)


def block_forward(blk, x, v_first):
    """Block.forward (model.py:247-254): [ln0] ; x += att(ln1 x) ; x += ffn(ln2 x), LayerNorms and residual adds
    fused into the neighbouring kernels."""
    _need_cuda_bf16(x, "Block")
    if blk.layer_id == 0:
        x = layer_norm(x, blk.ln0)
    m = blk.att
    x, v = fused.TmixBlockFn.apply(x, v_first if blk.layer_id != 0 else None, blk.ln1.weight, blk.ln1.bias, *_tmix_args(m),
                                   blk.layer_id, m.n_head, blk.ln1.eps, m.ln_x.eps, True)
    if blk.layer_id == 0:
        v_first = v
    f = blk.ffn
    x = fused.CmixBlockFn.apply(x, blk.ln2.weight, blk.ln2.bias, f.x_k, f.key.weight, f.value.weight, blk.ln2.eps, True)
    return x, v_first


def projector_forward(m, x):
    """MLPWithContextGating.forward (model.py:335-338): LN(o_proj(x * sigmoid(gate(x))))."""
    _need_cuda_bf16(x, "MLPWithContextGating")
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


_pending_count = None  # (event, pinned count tensor, n_features, sample_ids) of the previous call


def _check_pending_count():
    global _pending_count
    if _pending_count is None:
        return
    ev, host, n_feat, sample_ids = _pending_count
    _pending_count = None
    ev.synchronize()
    n_sel = int(host)
    if n_sel != n_feat:
        warnings.warn(f"sample_id: {':::'.join(sample_ids or [])}, image tokens: {n_sel}, but image features: {n_feat}")
    if n_sel > n_feat:
        raise RuntimeError(f"{n_sel} image-token slots but only {n_feat} image feature rows (model.py:487-493)")


def embed_and_scatter(emb_weight, input_ids, image_features, image_token_index, sample_ids=None):
    """preparing_embedding (model.py:481-493): emb(input_ids) with the rows where ids == 65535 replaced by the
    image features in row-major order of appearance (bit-exact copy; gradient flows to the features).  Surplus
    feature rows are dropped like the reference's `image_features[:selected.sum()]` (:487-491).

    The reference reads `selected.sum()` on the host every step (a device synchronisation in the middle of the
    forward).  Here the k-th selected row gathers feature row k on the device, and the count travels
    to pinned host memory asynchronously: the reference's mismatch warning (and an error when there are more slots
    than features) is raised at the next call / `flush_checks()` instead of stalling this one."""
    global _pending_count
    _check_pending_count()
    B, L = input_ids.shape
    D = emb_weight.shape[1]
    x = F.embedding(input_ids, emb_weight).view(B * L, D)
    sel = input_ids.view(-1) == image_token_index
    feats = image_features.reshape(-1, D).to(x.dtype)
    rank = torch.cumsum(sel, dim=0, dtype=torch.int32) - 1          # k for the k-th selected row
    if feats.shape[0] == 0:
        return x.view(B, L, D)
    gathered = feats.index_select(0, rank.clamp(0, feats.shape[0] - 1).to(torch.int64))
    x = torch.where(sel.unsqueeze(-1), gathered, x)
    if input_ids.is_cuda:
        host = torch.empty((), dtype=torch.int32, pin_memory=True)
        host.copy_(rank[-1] + 1, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        _pending_count = (ev, host, feats.shape[0], sample_ids)
    else:
        n_sel = int(rank[-1]) + 1
        if n_sel != feats.shape[0]:
            warnings.warn(f"sample_id: {':::'.join(sample_ids or [])}, image tokens: {n_sel}, but image features: {feats.shape[0]}")
    return x.view(B, L, D)


def flush_checks():
    """Raise / warn now for the deferred image-token count check of the last embed_and_scatter call."""
    _check_pending_count()


def training_loss(logits, targets, ignore_index, l2wrap):
    """VisualRWKV.training_step (model.py:418-434)."""
    shift_logits = logits[..., :-1, :].contiguous()
    shift_labels = targets[..., 1:].contiguous()
    valid = (shift_labels != ignore_index).sum(1).clamp(min=1)
    loss = F.cross_entropy(shift_logits.view(-1, shift_logits.size(-1)), shift_labels.view(-1),
                           ignore_index=ignore_index, reduction="none")
    loss = (loss.view(shift_labels.size()).sum(1) / valid).mean()
    return l2wrap.apply(loss, logits)


def head_loss(features, head_weight, targets, ignore_index):
    """head GEMM + training_step loss + L2Wrap fused (model.py:323-325, 418-434, 257-271); `features` = ln_out output."""
    _need_cuda_bf16(features, "head_loss")
    return fused.HeadLossFn.apply(features, head_weight, targets, ignore_index)
