"""Functional restatement of the reference model path (TEST INFRASTRUCTURE ONLY).

Pure functions over a state-dict (reference key names, SURVEY.md §5 "Checkpoint / resume"), stock
PyTorch ops only, device/dtype agnostic:

  * CPU fp64 / fp32  -> ground truth for the parity tests and the CPU baseline of bench.py
    (WKV7 through the C oracle, oracle/wkv7_oracle.c);
  * CUDA bf16 with `wkv=ref_kernel_wkv` -> the module-level GPU reference of BASELINE.md §2.1
    (reference kernel from oracle/_ref, stock PyTorch for everything else).

Follows VisualRWKV-v7/v7.00/src/model.py:
  time-mix :163-195   channel-mix :221-227   block :247-254   RWKV.forward :306-325 (+pad :286-304)
  projector :328-338  pooling :442-447       scatter :473-494  loss :418-434  L2Wrap :257-271
and, for the v7.01 SigLIP tower, transformers' SiglipVisionModel.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from . import wkv7 as O

IGNORE_INDEX = -100        # v7.00/src/dataset.py:17
IMAGE_TOKEN_INDEX = 65535  # dataset.py:18
STOP_TOKEN_INDEX = 261     # dataset.py:19
CHUNK_LEN = 16


# --------------------------------------------------------------------------------------------
# WKV7 back-ends
# --------------------------------------------------------------------------------------------
class _OracleWKV(torch.autograd.Function):
    """CPU autograd node around the C oracle; computes in float64 or float32 (no bf16 rounding of
    the outputs unless the caller's tensors are bf16)."""

    @staticmethod
    def forward(ctx, w, q, k, v, a, b, precision):
        y, s, sa = O.forward(w, q, k, v, a, b, precision=precision)
        ctx.save_for_backward(w, q, k, v, a, b)
        ctx.aux = (s, sa, precision)
        return torch.from_numpy(y).to(w.dtype)

    @staticmethod
    def backward(ctx, dy):
        w, q, k, v, a, b = ctx.saved_tensors
        s, sa, precision = ctx.aux
        g = O.backward(w, q, k, v, a, b, dy.contiguous(), s, sa, precision=precision)
        return tuple(torch.from_numpy(x).to(w.dtype) for x in g) + (None,)


def oracle_wkv(precision: str = "f64"):
    def run(r, w, k, v, a, b):  # python argument order of RUN_CUDA_RWKV7g (model.py:67)
        B, T, HC = r.shape
        r4, w4, k4, v4, a4, b4 = [x.reshape(B, T, HC // 64, 64).contiguous() for x in (r, w, k, v, a, b)]
        return _OracleWKV.apply(w4, r4, k4, v4, a4, b4, precision).reshape(B, T, HC)
    return run


class _RefKernelWKV(torch.autograd.Function):
    @staticmethod
    def forward(ctx, w, q, k, v, a, b):
        from . import ref_kernel as RK
        y, s, sa = RK.forward(w, q, k, v, a, b)
        ctx.save_for_backward(w, q, k, v, a, b, s, sa)
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import ref_kernel as RK
        w, q, k, v, a, b, s, sa = ctx.saved_tensors
        return RK.backward(w, q, k, v, a, b, dy.contiguous(), s, sa)


def ref_kernel_available() -> bool:
    from oracle import ref_kernel as RK
    return torch.cuda.is_available() and RK.available()


def ref_kernel_wkv(r, w, k, v, a, b):
    """The unmodified reference CUDA kernel (oracle/_ref) behind the RUN_CUDA_RWKV7g signature."""
    B, T, HC = r.shape
    r4, w4, k4, v4, a4, b4 = [x.view(B, T, HC // 64, 64).contiguous() for x in (r, w, k, v, a, b)]
    return _RefKernelWKV.apply(w4, r4, k4, v4, a4, b4).view(B, T, HC)


# --------------------------------------------------------------------------------------------
# Blocks
# --------------------------------------------------------------------------------------------
def token_shift(x):
    """nn.ZeroPad2d((0,0,1,-1)): x[:, t-1], zero at t = 0 (model.py:149)."""
    return torch.cat([torch.zeros_like(x[:, :1]), x[:, :-1]], dim=1)


def tmix(P, pre, x, v_first, layer_id, n_head, wkv, head_size_divisor=8):
    B, T, C = x.shape
    H = n_head
    g = lambda n: P[pre + n]
    xx = token_shift(x) - x
    xr, xw, xk, xv, xa, xg = [x + xx * g(n) for n in ("x_r", "x_w", "x_k", "x_v", "x_a", "x_g")]
    r = xr @ g("receptance.weight").T
    w = -F.softplus(-(g("w0") + torch.tanh(xw @ g("w1")) @ g("w2"))) - 0.5
    k = xk @ g("key.weight").T
    v = xv @ g("value.weight").T
    if layer_id == 0:
        v_first = v
    else:
        v = v + (v_first - v) * torch.sigmoid(g("v0") + (xv @ g("v1")) @ g("v2"))
    a = torch.sigmoid(g("a0") + (xa @ g("a1")) @ g("a2"))
    gate = torch.sigmoid(xg @ g("g1")) @ g("g2")
    kk = F.normalize((k * g("k_k")).view(B, T, H, -1), dim=-1, p=2.0).view(B, T, C)
    k = k * (1 + (a - 1) * g("k_a"))
    y = wkv(r, w, k, v, -kk, kk * a)
    y = F.group_norm(y.reshape(B * T, C), H, g("ln_x.weight"), g("ln_x.bias"),
                     eps=1e-5 * head_size_divisor ** 2).view(B, T, C)
    y = y + ((r.view(B, T, H, -1) * k.view(B, T, H, -1) * g("r_k")).sum(-1, keepdim=True)
             * v.view(B, T, H, -1)).view(B, T, C)
    return (y * gate) @ g("output.weight").T, v_first


def cmix(P, pre, x):
    xx = token_shift(x) - x
    k = x + xx * P[pre + "x_k"]
    k = torch.relu(k @ P[pre + "key.weight"].T) ** 2
    return k @ P[pre + "value.weight"].T


def layer_norm(P, pre, x):
    return F.layer_norm(x, (x.shape[-1],), P[pre + "weight"], P[pre + "bias"], 1e-5)


def block(P, i, x, v_first, n_head, wkv, pre="rwkv."):
    bp = f"{pre}blocks.{i}."
    if i == 0:
        x = layer_norm(P, bp + "ln0.", x)
    xx, v_first = tmix(P, bp + "att.", layer_norm(P, bp + "ln1.", x), v_first, i, n_head, wkv)
    x = x + xx
    x = x + cmix(P, bp + "ffn.", layer_norm(P, bp + "ln2.", x))
    return x, v_first


def rwkv_forward(P, x, n_layer, n_head, wkv, pre="rwkv.", with_head=True):
    """RWKV.forward: left-pad to a multiple of 16 with emb(261), L blocks, ln_out, head, un-pad."""
    T = x.shape[1]
    pad = (CHUNK_LEN - T % CHUNK_LEN) % CHUNK_LEN
    if pad:
        eos = P[pre + "emb.weight"][torch.full((x.shape[0], pad), STOP_TOKEN_INDEX, dtype=torch.long, device=x.device)]
        x = torch.cat([eos.to(x.dtype), x], dim=1)
    v_first = torch.empty_like(x)
    for i in range(n_layer):
        x, v_first = block(P, i, x, v_first, n_head, wkv, pre)
    x = layer_norm(P, pre + "ln_out.", x)
    if with_head:
        x = x @ P[pre + "head.weight"].T
    return x[:, pad:] if pad else x


def projector(P, x, pre="proj."):
    """MLPWithContextGating (model.py:328-338)."""
    gating = torch.sigmoid(x @ P[pre + "gate.weight"].T)
    return layer_norm(P, pre + "ln_v.", (x * gating) @ P[pre + "o_proj.weight"].T)


def adaptive_pooling(feats, num_token_per_image):
    B, L, D = feats.shape
    hw = int(L ** 0.5)
    out = int(num_token_per_image ** 0.5)
    f = feats.view(B, hw, hw, D).permute(0, 3, 1, 2)
    f = F.adaptive_avg_pool2d(f.float(), out).to(feats.dtype)
    return f.reshape(B, D, -1).permute(0, 2, 1)


def scatter_image_features(P, input_ids, image_features, pre="rwkv."):
    """preparing_embedding (model.py:481-493): rows where ids == 65535 take the image features in
    row-major order of appearance; features are truncated if there are fewer slots."""
    emb = P[pre + "emb.weight"][input_ids]
    B, L, D = emb.shape
    flat = emb.reshape(B * L, D).clone()
    sel = input_ids.reshape(-1) == IMAGE_TOKEN_INDEX
    feats = image_features.reshape(-1, D)[: int(sel.sum())]
    flat[sel] = feats.to(flat.dtype)
    return flat.view(B, L, D)


# --------------------------------------------------------------------------------------------
# SigLIP vision tower (third-party arithmetic: transformers.SiglipVisionModel, Appendix A.3b)
# --------------------------------------------------------------------------------------------
def siglip_forward(P, pixels, cfg, pre="vit."):
    """last_hidden_state of SiglipVisionModel, restated.  cfg: dict(hidden, layers, heads, patch,
    eps).  Keys follow HF: vision_model.embeddings.patch_embedding.{weight,bias}, ...position_embedding.weight,
    vision_model.encoder.layers.N.{layer_norm1,self_attn.{q,k,v,out}_proj,layer_norm2,mlp.fc1,mlp.fc2},
    vision_model.post_layernorm."""
    vp = pre + "vision_model."
    D, nh, p = cfg["hidden"], cfg["heads"], cfg["patch"]
    x = F.conv2d(pixels, P[vp + "embeddings.patch_embedding.weight"], P[vp + "embeddings.patch_embedding.bias"], stride=p)
    x = x.flatten(2).transpose(1, 2) + P[vp + "embeddings.position_embedding.weight"]
    N, S, _ = x.shape
    hd = D // nh
    for i in range(cfg["layers"]):
        lp = f"{vp}encoder.layers.{i}."
        h = F.layer_norm(x, (D,), P[lp + "layer_norm1.weight"], P[lp + "layer_norm1.bias"], cfg["eps"])
        q = F.linear(h, P[lp + "self_attn.q_proj.weight"], P[lp + "self_attn.q_proj.bias"]).view(N, S, nh, hd).transpose(1, 2)
        k = F.linear(h, P[lp + "self_attn.k_proj.weight"], P[lp + "self_attn.k_proj.bias"]).view(N, S, nh, hd).transpose(1, 2)
        v = F.linear(h, P[lp + "self_attn.v_proj.weight"], P[lp + "self_attn.v_proj.bias"]).view(N, S, nh, hd).transpose(1, 2)
        sc = q @ k.transpose(-1, -2)
        sc = sc.float() if sc.dtype in (torch.bfloat16, torch.float16) else sc  # softmax in >= fp32 (HF eager)
        att = torch.softmax(sc * hd ** -0.5, dim=-1).to(q.dtype)
        o = (att @ v).transpose(1, 2).reshape(N, S, D)
        x = x + F.linear(o, P[lp + "self_attn.out_proj.weight"], P[lp + "self_attn.out_proj.bias"])
        h = F.layer_norm(x, (D,), P[lp + "layer_norm2.weight"], P[lp + "layer_norm2.bias"], cfg["eps"])
        h = F.gelu(F.linear(h, P[lp + "mlp.fc1.weight"], P[lp + "mlp.fc1.bias"]), approximate="tanh")
        x = x + F.linear(h, P[lp + "mlp.fc2.weight"], P[lp + "mlp.fc2.bias"])
    return F.layer_norm(x, (D,), P[vp + "post_layernorm.weight"], P[vp + "post_layernorm.bias"], cfg["eps"])


# --------------------------------------------------------------------------------------------
# Whole model + loss
# --------------------------------------------------------------------------------------------
def visual_forward(P, samples, cfg, wkv):
    """VisualRWKV.forward(samples) -> (logits, targets)  (model.py:412-416,473-494)."""
    ids = samples["input_ids"]
    if "images" in samples:
        with torch.no_grad():
            feats = siglip_forward(P, samples["images"], cfg["vit"])
        feats = adaptive_pooling(feats.detach(), cfg["num_token_per_image"])
        x = scatter_image_features(P, ids, projector(P, feats))
    else:
        x = P["rwkv.emb.weight"][ids]
    return rwkv_forward(P, x, cfg["n_layer"], cfg["n_head"], wkv), samples["labels"]


class _L2Wrap(torch.autograd.Function):
    @staticmethod
    def forward(ctx, loss, y):
        ctx.save_for_backward(y)
        return loss

    @staticmethod
    def backward(ctx, grad_output):
        y = ctx.saved_tensors[0]
        factor = 1e-4 / (y.shape[0] * y.shape[1])
        maxx, ids = torch.max(y, -1, keepdim=True)
        gy = torch.zeros_like(y)
        gy.scatter_(-1, ids, maxx * factor)
        return grad_output, gy


def training_loss(logits, targets):
    """training_step (model.py:418-434): shifted CE, per-sample mean over valid labels, batch mean,
    wrapped in L2Wrap."""
    sl = logits[..., :-1, :].contiguous()
    tl = targets[..., 1:].contiguous()
    valid = (tl != IGNORE_INDEX).sum(1).clamp(min=1)
    loss = F.cross_entropy(sl.view(-1, sl.size(-1)).float() if sl.dtype != torch.float64 else sl.view(-1, sl.size(-1)),
                           tl.view(-1), ignore_index=IGNORE_INDEX, reduction="none")
    loss = (loss.view(tl.size()).sum(1) / valid).mean()
    return _L2Wrap.apply(loss.to(logits.dtype) if logits.dtype == torch.float64 else loss, logits)


# --------------------------------------------------------------------------------------------
# Synthetic batch (model level)
# --------------------------------------------------------------------------------------------
def make_batch(B, T, n_img_tok, image_size, seed=0, vocab=65536, device="cpu", img_dtype=torch.float32,
               human_tokens=40):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, vocab - 1, (B, T), generator=g)
    labels = ids.clone()
    for b in range(B):
        start = int(torch.randint(0, 9, (1,), generator=g))
        ids[b, start:start + n_img_tok] = IMAGE_TOKEN_INDEX
        labels[b, : start + n_img_tok + human_tokens] = IGNORE_INDEX
    images = torch.randn(B, 3, image_size, image_size, generator=g)
    return {"input_ids": ids.to(device), "labels": labels.to(device), "images": images.to(device=device, dtype=img_dtype),
            "sample_id": [str(i) for i in range(B)]}
