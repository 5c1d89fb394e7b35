"""Operator layer between the reference-shaped modules (model.py) and the native kernels.

Each function cites the reference lines it implements.  Plain GEMMs go to cuBLAS through torch.matmul;
everything else on the path is a hand-written sm_100a kernel reached through the C ABI (fused.py, wkv7.py).
There is no CPU / eager fallback here: CUDA bf16 tensors are required and the native library must load.
"""
from __future__ import annotations

import ctypes
import warnings

import torch
import torch.nn.functional as F

from . import _lib, fused


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
    shp = x.shape
    x2 = x.reshape(-1, shp[-1]).contiguous()
    rows, Dv = x2.shape
    C = m.o_proj.weight.shape[0]
    if fused.gemm2_supported(rows, Dv, Dv) and fused.gemm2_supported(rows, C, Dv) and fused.gemm2_supported(Dv, Dv, rows) and \
            fused.gemm2_supported(C, Dv, rows) and fused.gemm2_supported(rows, Dv, C):
        o = fused.ProjectorFn.apply(x2, m.gate.weight, m.o_proj.weight)    # CTA-pair GEMMs + the gate kernel, hand-written backward
        return layer_norm(o.view(*shp[:-1], C), m.ln_v)
    gating = torch.sigmoid(m.gate(x))
    return layer_norm(m.o_proj(x * gating), m.ln_v)


def adaptive_pooling(feats, out_hw):
    """VisualRWKV.adaptive_pooling (model.py:442-447): [N,L,D] -> [N,out_hw^2,D] (csrc/glue.cu; the tower is frozen and
    detached, so there is no backward)."""
    N, L, D = feats.shape
    hw = int(L ** 0.5)
    if hw == out_hw:
        return feats
    if not (feats.is_cuda and feats.dtype == torch.bfloat16 and D % 8 == 0) or feats.requires_grad:
        f = feats.view(N, hw, hw, D).permute(0, 3, 1, 2)
        f = F.adaptive_avg_pool2d(f, out_hw)
        return f.reshape(N, D, -1).permute(0, 2, 1).contiguous()
    x = feats.contiguous()
    y = torch.empty(N, out_hw * out_hw, D, dtype=x.dtype, device=x.device)
    fused._chk(_lib.lib().vrwkv_adaptive_pool(N, hw, out_hw, D, fused._p(x), fused._p(y), _lib.cur_stream()), "vrwkv_adaptive_pool")
    return y


class _PendingCount:
    """Deferred image-token count check of one embed_and_scatter call (per call, not module state: re-entrant across
    models / streams)."""

    def __init__(self, event, host, n_feat, sample_ids):
        self.event, self.host, self.n_feat, self.sample_ids = event, host, n_feat, sample_ids

    def check(self):
        self.event.synchronize()
        n_sel = int(self.host)
        if n_sel != self.n_feat:
            warnings.warn(f"sample_id: {':::'.join(self.sample_ids or [])}, image tokens: {n_sel}, but image features: {self.n_feat}")
        if n_sel > self.n_feat:
            raise RuntimeError(f"{n_sel} image-token slots but only {self.n_feat} image feature rows (model.py:487-493)")


_pending = []  # checks not yet looked at (drained by the next call on any model and by flush_checks())


def _drain_pending():
    while _pending:
        _pending.pop(0).check()


class EmbedScatterFn(torch.autograd.Function):
    """(emb table [V,D], ids [B,L] int64, features [F,D]) -> [B,L,D]; gradient to the features only (frozen table)."""

    @staticmethod
    def forward(ctx, emb_weight, input_ids, feats, image_token_index, count_dev):
        B, L = input_ids.shape
        D = emb_weight.shape[1]
        ids = input_ids.contiguous()
        out = torch.empty(B * L, D, dtype=emb_weight.dtype, device=emb_weight.device)
        fused._chk(_lib.lib().vrwkv_embed_scatter_forward(B * L, D, feats.shape[0], ctypes.c_longlong(image_token_index), fused._p(ids),
                                                          fused._p(emb_weight), fused._p(feats), fused._p(out), fused._p(count_dev),
                                                          _lib.cur_stream()), "vrwkv_embed_scatter_forward")
        ctx.save_for_backward(ids)
        ctx.meta = (B * L, D, feats.shape[0], image_token_index)
        return out.view(B, L, D)

    @staticmethod
    def backward(ctx, dout):
        (ids,) = ctx.saved_tensors
        ntok, D, nfeat, idx = ctx.meta
        do = dout.reshape(ntok, D).contiguous()
        dfeats = torch.empty(nfeat, D, dtype=do.dtype, device=do.device)
        fused._chk(_lib.lib().vrwkv_embed_scatter_backward(ntok, D, nfeat, ctypes.c_longlong(idx), fused._p(ids), fused._p(do), fused._p(dfeats),
                                                           _lib.cur_stream()), "vrwkv_embed_scatter_backward")
        return None, None, dfeats, None, None


def embed_and_scatter(emb_weight, input_ids, image_features, image_token_index, sample_ids=None):
    """preparing_embedding (model.py:481-493): emb(input_ids) with the rows where ids == 65535 replaced by the
    image features in row-major order of appearance (bit-exact copy; gradient flows to the features).  Surplus
    feature rows are dropped like the reference's `image_features[:selected.sum()]` (:487-491).

    One kernel (csrc/glue.cu) instead of embedding + cumsum + index_select + where.  The reference reads
    `selected.sum()` on the host every step (a device synchronisation in the middle of the forward);
    here the count travels to pinned host memory asynchronously and the reference's mismatch warning (an error when
    there are more slots than features) is raised at the next call / `flush_checks()` instead of stalling this one."""
    if not (input_ids.is_cuda and torch.cuda.is_current_stream_capturing()):
        _drain_pending()
    B, L = input_ids.shape
    D = emb_weight.shape[1]
    feats = image_features.reshape(-1, D).to(emb_weight.dtype)
    on_gpu = input_ids.is_cuda and emb_weight.dtype == torch.bfloat16 and D % 8 == 0 and not emb_weight.requires_grad
    if not on_gpu:   # CPU / trainable-table path (tests of the host logic): eager
        x = F.embedding(input_ids, emb_weight).view(B * L, D)
        sel = input_ids.view(-1) == image_token_index
        rank = torch.cumsum(sel, dim=0, dtype=torch.int32) - 1
        if feats.shape[0] > 0:
            gathered = feats.index_select(0, rank.clamp(0, feats.shape[0] - 1).to(torch.int64))
            x = torch.where(sel.unsqueeze(-1), gathered, x)
        n_sel = int(rank[-1]) + 1
        if n_sel != feats.shape[0]:
            warnings.warn(f"sample_id: {':::'.join(sample_ids or [])}, image tokens: {n_sel}, but image features: {feats.shape[0]}")
        if n_sel > feats.shape[0]:
            raise RuntimeError(f"{n_sel} image-token slots but only {feats.shape[0]} image feature rows (model.py:487-493)")
        return x.view(B, L, D)
    count_dev = torch.empty((), dtype=torch.int32, device=input_ids.device)
    out = EmbedScatterFn.apply(emb_weight, input_ids, feats.contiguous(), int(image_token_index), count_dev)
    if not torch.cuda.is_current_stream_capturing():
        host = torch.empty((), dtype=torch.int32, pin_memory=True)
        host.copy_(count_dev, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        _pending.append(_PendingCount(ev, host, feats.shape[0], sample_ids))
    return out


def flush_checks():
    """Raise / warn now for the deferred image-token count checks of earlier embed_and_scatter calls."""
    _drain_pending()


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
