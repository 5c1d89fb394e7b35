"""Recurrent (stateful) inference of the RWKV-x070 tower: prefill + O(1)-per-token decode.

The reference's `VisualRWKV.generate` (VisualRWKV-v7/v7.00/src/model.py:496-535) re-runs `self.rwkv` on the whole, growing
sequence for every new token — O(T^2).  RWKV-7 is a recurrence: what a later token needs from the past is, per layer,
  * the WKV state S [B,H,64,64] fp32 (wkv7_cuda.cu:17-51 keeps exactly this in registers),
  * the previous token's input to the time-mix and to the channel-mix (the nn.ZeroPad2d token shift, model.py:149,166,205,222).
`rwkv_forward_recurrent` carries those, so a prompt is processed once (T a multiple of 64: the chunked tensor-core WKV7
kernel; otherwise the step-by-step kernel) and every further token costs one pass over the weights.

Built from the same kernels as the training path (csrc/fused_ln.cu, fused_tmix.cu, gemm2_sm100.cu, the stateful WKV7
entry point); the token shift across the call boundary is obtained by prepending the carried row to the sequence the
LayerNorm + shift + mix kernel sees and dropping its outputs — the kernel's own arithmetic, no second implementation.
There is no reference oracle for this row: tests check split-sequence equivalence against the stateless path."""
from __future__ import annotations

from dataclasses import dataclass, field

import torch

from . import fused
from . import wkv7 as _wkv7


@dataclass
class LayerState:
    att_prev: torch.Tensor | None = None   # [B, C] bf16: the block input x of the last token (before ln1)
    ffn_prev: torch.Tensor | None = None   # [B, C] bf16: the channel-mix input of the last token (before ln2)
    wkv: torch.Tensor | None = None        # [B, H, 64, 64] fp32
    static: bool = False                   # True: the tensors above keep their addresses (updated in place)


@dataclass
class RecurrentState:
    layers: list = field(default_factory=list)
    tokens_seen: int = 0


def _mm(x, w):
    """x [M,K] @ w[N,K]^T on the CTA-pair GEMM (any M; K, N multiples of 8), else the library."""
    if x.shape[1] % 8 == 0 and w.shape[0] % 8 == 0:
        return fused.gemm2(x, w)
    return x @ w.t()


def _mix(x2, B, T, prev, ln, coefs):
    """LayerNorm + token shift + lerps of `x2` [B*T, C] with the previous call's last row as the row before t = 0."""
    C = x2.shape[1]
    if prev is None:
        outs, _, _ = fused.ln_mix_forward(x2, T, ln.weight, ln.bias, ln.eps, coefs)
    else:
        xe = torch.cat([prev.view(B, 1, C), x2.view(B, T, C)], dim=1).reshape(B * (T + 1), C)
        outs_e, _, _ = fused.ln_mix_forward(xe, T + 1, ln.weight, ln.bias, ln.eps, coefs)
        outs = [o.view(B, T + 1, C)[:, 1:].reshape(B * T, C).contiguous() for o in outs_e]
    return outs, x2.view(B, T, C)[:, -1].contiguous()


def _carry(st: LayerState, name: str, new):
    """Store the carried row: in place when the state is static (its address is baked into a CUDA graph)."""
    old = getattr(st, name)
    if st.static and old is not None:
        old.copy_(new)
    else:
        setattr(st, name, new)


def tmix_step(blk, x2, B, T, v_first, st: LayerState):
    """x + att(ln1(x)) for T more tokens (model.py:163-195, :250-251); returns (out [B*T,C], v_first)."""
    m = blk.att
    C, H = x2.shape[1], m.n_head
    coefs = [c.reshape(C) for c in (m.x_r, m.x_w, m.x_k, m.x_v, m.x_a, m.x_g)]
    (xr, xw, xk, xv, xa, xg), new_prev = _mix(x2, B, T, st.att_prev, blk.ln1, coefs)
    _carry(st, "att_prev", new_prev)
    r, k, v = fused.gemm2_grouped([xr, xk, xv], [m.receptance.weight, m.key.weight, m.value.weight])
    has_vres = m.layer_id != 0
    downs = [(xw, m.w1, fused.ACT_TANH), (xa, m.a1, fused.ACT_NONE), (xg, m.g1, fused.ACT_SIGMOID)] + \
        ([(xv, m.v1, fused.ACT_NONE)] if has_vres else [])
    ups = [m.w2, m.a2, m.g2] + ([m.v2] if has_vres else [])
    hs = fused.gemm2_grouped([t for t, _, _ in downs], [w for _, w, _ in downs], fused.G2_NN, fused.EPI_ACT, acts=[a for _, _, a in downs])
    outs = fused.gemm2_grouped(hs, ups, fused.G2_NN)
    ww, aa, g = outs[:3]
    vv = outs[3] if has_vres else None
    w, k2, v2, nkk, kka = fused.tmix_mid_forward(k, v, v_first if has_vres else None, ww, aa, vv, m.w0.reshape(C), m.a0.reshape(C),
                                                 m.v0.reshape(C) if has_vres else None, m.k_k.reshape(C), m.k_a.reshape(C))
    v4 = lambda t: t.view(B, T, H, 64)
    inplace = st.static and st.wkv is not None and T % 64 != 0   # static state (CUDA-graphed decode): updated where it lives
    y, st.wkv = _wkv7.wkv7_forward_state(v4(w), v4(r), v4(k2), v4(v2), v4(nkk), v4(kka), state_in=st.wkv, bounded_decay=True,
                                         state_out=st.wkv if inplace else None)
    z = fused.tmix_post_forward(y.view(B * T, C), r, k2, v2, g, m.ln_x.weight, m.ln_x.bias, m.r_k.reshape(C), m.ln_x.eps)
    out = fused.gemm2(z, m.output.weight, fused.G2_TN, fused.EPI_ADD, x2)
    return out, (v if not has_vres else v_first)


def cmix_step(blk, x2, B, T, st: LayerState):
    """x + ffn(ln2(x)) for T more tokens (model.py:221-227, :252-253)."""
    f = blk.ffn
    C = x2.shape[1]
    (xk,), new_prev = _mix(x2, B, T, st.ffn_prev, blk.ln2, [f.x_k.reshape(C)])
    _carry(st, "ffn_prev", new_prev)
    act = fused.gemm2(xk, f.key.weight, fused.G2_TN, fused.EPI_RELU_SQ)
    return fused.gemm2(act, f.value.weight, fused.G2_TN, fused.EPI_ADD, x2)


@torch.no_grad()
def rwkv_forward_recurrent(rwkv, x, state: RecurrentState | None = None, last_only: bool = True):
    """RWKV.forward (model.py:306-325) on `x` [B,T,C] continuing from `state`; no left padding (the state carries on
    where the previous call stopped).  Returns (logits [B,V] of the last token, or [B,T,V]), new state."""
    from . import ops
    B, T, C = x.shape
    if x.dtype != torch.bfloat16 or not x.is_cuda:
        raise RuntimeError("rwkv_forward_recurrent: CUDA bf16 input expected")
    if C % 64 or C > 1024 * 2:
        raise RuntimeError("rwkv_forward_recurrent: n_embd must be a multiple of 64")
    if state is None:
        state = RecurrentState(layers=[LayerState() for _ in rwkv.blocks])
    x2 = x.reshape(B * T, C).contiguous()
    v_first = None
    for blk, st in zip(rwkv.blocks, state.layers):
        if blk.layer_id == 0:
            x2 = ops.layer_norm(x2, blk.ln0)
        x2, v_first = tmix_step(blk, x2, B, T, v_first, st)
        x2 = cmix_step(blk, x2, B, T, st)
    feats = x2.view(B, T, C)[:, -1].contiguous() if last_only else x2
    h = ops.layer_norm(feats, rwkv.ln_out)
    logits = _mm(h, rwkv.head.weight)
    state.tokens_seen += T
    return (logits if last_only else logits.view(B, T, -1)), state


@torch.no_grad()
def generate_recurrent(model, input_ids, images, max_new_tokens, stop_token_idx):
    """Greedy decoding as VisualRWKV.generate does it (model.py:496-535: argmax, its logit, its softmax probability),
    with the prompt processed once and one recurrent step per new token."""
    samples = {"input_ids": input_ids, "images": images, "labels": torch.full_like(input_ids, -100)}
    x, _ = model.preparing_embedding(samples)
    toks, tok_logits, tok_probs = [], [], []
    logits, state = rwkv_forward_recurrent(model.rwkv, x.to(torch.bfloat16))
    for _ in range(max_new_tokens):
        nxt = torch.argmax(logits, dim=-1, keepdim=True)                      # [B,1]
        probs = torch.softmax(logits.float(), dim=-1)
        toks.append(int(nxt[0, 0]))
        tok_logits.append(float(logits.gather(-1, nxt)[0, 0]))
        tok_probs.append(float(probs.gather(-1, nxt)[0, 0]))
        if toks[-1] == stop_token_idx:
            break
        logits, state = rwkv_forward_recurrent(model.rwkv, model.rwkv.emb(nxt).to(torch.bfloat16), state)
    return toks, tok_logits, tok_probs


class GraphedDecoder:
    """One decode step (T = 1, fixed batch) replayed as a CUDA graph: ~25 small launches per layer are launch-bound when
    issued one by one.  The state of `state` becomes static (updated in place by the graph); `step(tokens)` returns the
    logits [B, V] of the next position (a static buffer, overwritten by the next call)."""

    def __init__(self, model, state: RecurrentState, batch: int):
        self.model, self.state = model, state
        rwkv = model.rwkv
        dev = rwkv.head.weight.device
        for st in state.layers:
            st.static = True
        self.tokens = torch.zeros(batch, 1, dtype=torch.long, device=dev)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        saved = [(s.att_prev.clone(), s.ffn_prev.clone(), s.wkv.clone()) for s in state.layers]
        with torch.cuda.stream(side):   # warm-up outside the capture (lazy initialisations, workspaces)
            for _ in range(2):
                rwkv_forward_recurrent(rwkv, rwkv.emb(self.tokens).to(torch.bfloat16), state)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.logits, _ = rwkv_forward_recurrent(rwkv, rwkv.emb(self.tokens).to(torch.bfloat16), state)
        for s, (a, f, w) in zip(state.layers, saved):   # the warm-up and the capture advanced nothing the caller should see
            s.att_prev.copy_(a); s.ffn_prev.copy_(f); s.wkv.copy_(w)
        state.tokens_seen -= 3

    @torch.no_grad()
    def step(self, tokens):
        self.tokens.copy_(tokens.view_as(self.tokens))
        self.graph.replay()
        self.state.tokens_seen += 1
        return self.logits
