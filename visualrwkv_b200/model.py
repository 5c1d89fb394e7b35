"""Host-side mirror of the reference model interface (VisualRWKV-v7/v7.00/src/model.py).

Same class names, constructor arguments (`args` namespace), forward signatures and state-dict key
names as the reference, so the reference `train.py` / checkpoints can use these classes unchanged
(SURVEY.md §5 "Checkpoint / resume", §8b).  The arithmetic is delegated to the sm_100a kernels in
csrc/ through `visualrwkv_b200.ops`; nothing here falls back to a CPU implementation.

  WindBackstepping / RUN_CUDA_RWKV7g   model.py:45-70     -> wkv7.py
  RWKV_Tmix_x070                        model.py:76-195
  RWKV_CMix_x070                        model.py:200-227
  Block                                 model.py:233-254
  L2Wrap                                model.py:257-271
  RWKV                                  model.py:274-325
  MLPWithContextGating                  model.py:328-338
  VisualRWKV                            model.py:341-530   (vision tower: the v7.01 HF-SigLIP variant,
                                        VisualRWKV-v7/v7.01/src/model.py:347-352,448-454)
"""
from __future__ import annotations

import math
import os
from types import SimpleNamespace

import torch
import torch.nn as nn
from torch.nn import functional as F

from . import ops
from .vision import SiglipVisionTower
from .wkv7 import CHUNK_LEN, RUN_CUDA_RWKV7g, WindBackstepping  # noqa: F401  (re-exported, reference names)

try:  # the reference derives from LightningModule (model.py:274,341); optional here
    import pytorch_lightning as pl  # type: ignore
    _Base = pl.LightningModule
except Exception:  # pragma: no cover - lightning is not installed in the build image
    pl = None
    _Base = nn.Module

IGNORE_INDEX = -100        # dataset.py:17
IMAGE_TOKEN_INDEX = 65535  # dataset.py:18
STOP_TOKEN_INDEX = 261     # dataset.py:19


def default_args(**kw) -> SimpleNamespace:
    """The `args` fields the model consumes, defaults = RWKV-x070 0.1B."""
    a = dict(n_embd=768, n_layer=12, dim_att=768, head_size_a=64, head_size_divisor=8, vocab_size=65536,
             dropout=0.0, grad_cp=0, load_model="", vision_tower_path="siglip-base-patch16-224",
             proj_type="mlp", num_token_per_image=576, ctx_len=2048, weight_decay=0.0, lr_init=1e-4,
             betas=(0.9, 0.99), adam_eps=1e-8)
    a.update(kw)
    return SimpleNamespace(**a)


def _lora_rank(C: int, coef: float, power: float = 0.5) -> int:
    return max(32, int(round((coef * (C ** power)) / 32) * 32))


def _ortho(shape, scale):
    x = torch.zeros(shape)
    gain = math.sqrt(shape[0] / shape[1]) if shape[0] > shape[1] else 1.0
    nn.init.orthogonal_(x, gain=gain * scale)
    return x


########################################################################################################
# RWKV TimeMix
########################################################################################################
class RWKV_Tmix_x070(nn.Module):
    """Parameters, shapes and initial values as model.py:96-160."""

    def __init__(self, args, layer_id):
        super().__init__()
        self.args = args
        self.layer_id = layer_id
        self.head_size = args.head_size_a
        self.n_head = args.dim_att // self.head_size
        assert args.dim_att % self.n_head == 0
        H, N, C = self.n_head, self.head_size, args.n_embd

        with torch.no_grad():
            r01 = layer_id / max(args.n_layer - 1, 1)
            r10 = 1.0 - (layer_id / args.n_layer)
            ddd = (torch.arange(C, dtype=torch.float32) / C).view(1, 1, C)
            self.x_r = nn.Parameter(1.0 - torch.pow(ddd, 0.2 * r10))
            self.x_w = nn.Parameter(1.0 - torch.pow(ddd, 0.9 * r10))
            self.x_k = nn.Parameter(1.0 - (torch.pow(ddd, 0.9 * r10) + 0.4 * r01))
            self.x_v = nn.Parameter(1.0 - (torch.pow(ddd, 0.4 * r10) + 0.6 * r01))
            self.x_a = nn.Parameter(1.0 - torch.pow(ddd, 0.9 * r10))
            self.x_g = nn.Parameter(1.0 - torch.pow(ddd, 0.2 * r10))

            D_w = _lora_rank(C, 1.8)
            self.w1 = nn.Parameter(torch.zeros(C, D_w))
            self.w2 = nn.Parameter(_ortho((D_w, C), 0.1))
            n = torch.arange(C, dtype=torch.float32)
            decay_speed = -7 + 5 * (n / max(C - 1, 1)) ** (0.85 + 1.0 * r01 ** 0.5)
            self.w0 = nn.Parameter(decay_speed.view(1, 1, C) + 0.5)

            D_a = _lora_rank(C, 1.8)
            self.a1 = nn.Parameter(torch.zeros(C, D_a))
            self.a2 = nn.Parameter(_ortho((D_a, C), 0.1))
            self.a0 = nn.Parameter(torch.zeros(1, 1, C))

            if layer_id != 0:
                D_v = _lora_rank(C, 1.3)
                self.v1 = nn.Parameter(torch.zeros(C, D_v))
                self.v2 = nn.Parameter(_ortho((D_v, C), 0.1))
                self.v0 = nn.Parameter(torch.zeros(1, 1, C) + 1.0)

            D_g = _lora_rank(C, 0.6, 0.8)
            self.g1 = nn.Parameter(torch.zeros(C, D_g))
            self.g2 = nn.Parameter(_ortho((D_g, C), 0.1))

            self.k_k = nn.Parameter(torch.ones(1, 1, C) * 0.85)
            self.k_a = nn.Parameter(torch.ones(1, 1, C))
            self.r_k = nn.Parameter(torch.zeros(H, N))

            self.receptance = nn.Linear(C, C, bias=False)
            self.key = nn.Linear(C, C, bias=False)
            self.value = nn.Linear(C, C, bias=False)
            self.output = nn.Linear(C, C, bias=False)
            self.ln_x = nn.GroupNorm(H, C, eps=(1e-5) * (args.head_size_divisor ** 2))

            self.receptance.weight.data.uniform_(-0.5 / (C ** 0.5), 0.5 / (C ** 0.5))
            self.key.weight.data.uniform_(-0.05 / (C ** 0.5), 0.05 / (C ** 0.5))
            self.value.weight.data.uniform_(-0.5 / (C ** 0.5), 0.5 / (C ** 0.5))
            self.output.weight.data.zero_()

    def forward(self, x, v_first):
        """model.py:163-195.  x: ln1 output [B,T,C] bf16.  Returns (att_out, v_first)."""
        return ops.tmix_forward(self, x, v_first)


########################################################################################################
# RWKV ChannelMix
########################################################################################################
class RWKV_CMix_x070(nn.Module):
    def __init__(self, args, layer_id):
        super().__init__()
        self.args = args
        self.layer_id = layer_id
        C = args.n_embd
        with torch.no_grad():
            r10 = 1.0 - (layer_id / args.n_layer)
            ddd = (torch.arange(C, dtype=torch.float32) / C).view(1, 1, C)
            self.x_k = nn.Parameter(1.0 - torch.pow(ddd, r10 ** 4))
        self.key = nn.Linear(C, C * 4, bias=False)   # hidden is 4*n_embd, not dim_ffn (model.py:214)
        self.value = nn.Linear(C * 4, C, bias=False)
        self.key.weight.data.uniform_(-0.5 / (C ** 0.5), 0.5 / (C ** 0.5))
        self.value.weight.data.zero_()

    def forward(self, x):
        """model.py:221-227."""
        return ops.cmix_forward(self, x)


########################################################################################################
# RWKV Block
########################################################################################################
class Block(nn.Module):
    def __init__(self, args, layer_id):
        super().__init__()
        self.args = args
        self.layer_id = layer_id
        if layer_id == 0:
            self.ln0 = nn.LayerNorm(args.n_embd)
        self.ln1 = nn.LayerNorm(args.n_embd)
        self.ln2 = nn.LayerNorm(args.n_embd)
        self.att = RWKV_Tmix_x070(args, layer_id)
        self.ffn = RWKV_CMix_x070(args, layer_id)

    def forward(self, x, v_first):
        """model.py:247-254."""
        return ops.block_forward(self, x, v_first)


class L2Wrap(torch.autograd.Function):
    """model.py:257-271: identity on the loss; adds 1e-4/(B*T) * max_logit at the argmax to dlogits."""

    @staticmethod
    def forward(ctx, loss, y):
        ctx.save_for_backward(y)
        return loss

    @staticmethod
    def backward(ctx, grad_output):
        (y,) = ctx.saved_tensors
        scale = 1e-4 / (y.shape[0] * y.shape[1])
        top, idx = y.max(dim=-1, keepdim=True)
        return grad_output, torch.zeros_like(y).scatter_(-1, idx, top * scale)


class RWKV(_Base):
    def __init__(self, args):
        super().__init__()
        self.args = args
        if args.n_embd % 64 or args.n_embd > 2048:
            # the fused row kernels keep a row's channels in one CTA / warp (csrc/fused_ln.cu, fused_tmix.cu): 0.1B .. 1.5B
            # (n_embd 768 .. 2048) are covered; RWKV-7 2.9B (n_embd 2560) is not
            raise ValueError(f"visualrwkv_b200: n_embd={args.n_embd} must be a multiple of 64 and <= 2048")
        self.emb = nn.Embedding(args.vocab_size, args.n_embd)
        self.blocks = nn.ModuleList([Block(args, i) for i in range(args.n_layer)])
        self.ln_out = nn.LayerNorm(args.n_embd)
        self.head = nn.Linear(args.n_embd, args.vocab_size, bias=False)
        if args.dropout > 0:
            self.drop0 = nn.Dropout(p=args.dropout)

    def pad_left(self, x, num_tokens_to_pad):
        """model.py:286-298: left-pad with the embedding of the stop token (261)."""
        if num_tokens_to_pad == 0:
            return x
        eos = torch.full((x.size(0), num_tokens_to_pad), STOP_TOKEN_INDEX, dtype=torch.long, device=x.device)
        return torch.cat((self.emb(eos).to(x.dtype), x), dim=1)

    def unpad(self, x, num_tokens_to_pad):
        return x[:, num_tokens_to_pad:] if num_tokens_to_pad > 0 else x

    def forward_features(self, x):
        """Everything of RWKV.forward up to and including ln_out (padded length)."""
        args = self.args
        pad = (CHUNK_LEN - x.size(1) % CHUNK_LEN) % CHUNK_LEN
        x = self.pad_left(x, pad)
        if args.dropout > 0:
            x = self.drop0(x)
        v_first = torch.empty_like(x)
        for blk in self.blocks:
            if args.grad_cp == 1 and torch.is_grad_enabled():
                x, v_first = torch.utils.checkpoint.checkpoint(blk, x, v_first, use_reentrant=False)
            else:
                x, v_first = blk(x, v_first)
        return ops.layer_norm(x, self.ln_out), pad

    def forward(self, x):
        """model.py:306-325."""
        x, pad = self.forward_features(x)
        return self.unpad(self.head(x), pad)


class MLPWithContextGating(nn.Module):
    def __init__(self, in_dim, n_embd):
        super().__init__()
        self.gate = nn.Linear(in_dim, in_dim, bias=False)
        self.o_proj = nn.Linear(in_dim, n_embd, bias=False)
        self.ln_v = nn.LayerNorm(n_embd)

    def forward(self, x):
        """model.py:335-338: LN(o_proj(x * sigmoid(gate(x))))."""
        return ops.projector_forward(self, x)


class VisualRWKV(_Base):
    def __init__(self, args):
        super().__init__()
        self.args = args
        self.rwkv = RWKV(args)
        if len(args.load_model) > 0:
            self.load_rwkv_from_pretrained(args.load_model)
        self.vit = SiglipVisionTower(args.vision_tower_path)
        self.freeze_vit()
        if args.proj_type == "linear":
            self.proj = nn.Linear(self.vit.embed_dim, args.n_embd, bias=False)
        else:
            self.proj = MLPWithContextGating(self.vit.embed_dim, args.n_embd)
        self.pool_out = int(args.num_token_per_image ** 0.5)

    def load_rwkv_from_pretrained(self, path):
        self.rwkv.load_state_dict(torch.load(path, map_location="cpu", weights_only=True))

    # ---- freeze policy (model.py:368-389) ----
    def freeze_vit(self):
        self.vit.requires_grad_(False)

    def freeze_rwkv(self, num_layers_to_freeze):
        if num_layers_to_freeze == self.args.n_layer:
            self.rwkv.requires_grad_(False)
        for i, blk in enumerate(self.rwkv.blocks):
            blk.requires_grad_(i >= num_layers_to_freeze)

    def freeze_emb(self):
        self.rwkv.emb.requires_grad_(False)

    def freeze_proj(self):
        self.proj.requires_grad_(False)

    def configure_optimizers(self):
        """model.py:391-410 without DeepSpeed: same two weight-decay groups, torch fused AdamW."""
        nd = [p for p in self.parameters() if p.requires_grad and len(p.squeeze().shape) < 2]
        wd = [p for p in self.parameters() if p.requires_grad and len(p.squeeze().shape) >= 2]
        groups = []
        if nd:
            groups.append({"params": nd, "weight_decay": 0.0})
        if wd:
            groups.append({"params": wd, "weight_decay": self.args.weight_decay if self.args.weight_decay > 0 else 0.0})
        return torch.optim.AdamW(groups, lr=self.args.lr_init, betas=self.args.betas, eps=self.args.adam_eps,
                                 fused=torch.cuda.is_available())

    # ---- forward path ----
    def adaptive_pooling(self, image_features):
        """model.py:442-447."""
        return ops.adaptive_pooling(image_features, self.pool_out)

    @torch.no_grad()
    def _vit_features(self, images):
        return self.vit(images)

    def encode_images(self, images, minibatch_size=4):
        """model.py:449-471.  `images` is a [N,3,H,W] tensor or the reference's dict with a 'siglip' entry.
        The reference's mini-batching + empty_cache() exist to cap a 40 GB activation peak on small
        GPUs (model.py:452); on a 180 GB part the tower runs the whole batch at once."""
        if isinstance(images, dict):
            images = images["siglip"]
        feats = self._vit_features(images).detach()
        feats = self.adaptive_pooling(feats)
        return self.proj(feats)

    def preparing_embedding(self, samples):
        """model.py:473-494."""
        if "images" not in samples:
            return self.rwkv.emb(samples["input_ids"]), samples["labels"]
        image_features = self.encode_images(samples["images"])
        x = ops.embed_and_scatter(self.rwkv.emb.weight, samples["input_ids"], image_features, IMAGE_TOKEN_INDEX,
                                  samples.get("sample_id"))
        return x, samples["labels"]

    def forward(self, samples):
        x, targets = self.preparing_embedding(samples)
        logits = self.rwkv(x)
        return logits, targets

    def training_step(self, batch, batch_idx=0):
        """model.py:418-434: shifted CE / valid length per sample, batch mean, L2Wrap."""
        x, targets = self.preparing_embedding(batch)
        feats, pad = self.rwkv.forward_features(x)
        if pad:  # left-padded rows carry no label (model.py:309-312,325)
            feats = feats[:, pad:]
        # head + shifted CE + L2Wrap in two passes over the logits (ops.head_loss); `self(batch)` followed by
        # ops.training_loss(logits, targets, ...) is the unfused, reference-shaped equivalent
        return ops.head_loss(feats.contiguous(), self.rwkv.head.weight, targets, IGNORE_INDEX)

    @torch.no_grad()
    def generate(self, input_ids, images, do_sample, temperature, top_p, max_new_tokens, stop_token_idx, recurrent=False):
        """model.py:496-530 (greedy only, as the reference).  Default: re-runs the full sequence per token exactly like the
        reference (left-padded to a multiple of 16 each time).  recurrent=True: the prompt is processed once and every new
        token is one stateful step (visualrwkv_b200/recurrent.py) — same arithmetic without the per-call
        left padding."""
        if recurrent:
            if do_sample:
                raise NotImplementedError
            from .recurrent import generate_recurrent
            return generate_recurrent(self, input_ids, images, max_new_tokens, stop_token_idx)
        samples = {"input_ids": input_ids, "images": images, "labels": torch.full_like(input_ids, IGNORE_INDEX)}
        x, _ = self.preparing_embedding(samples)
        toks, logits_l, probs_l = [], [], []
        for _ in range(max_new_tokens):
            logits = self.rwkv(x)[:, -1, :]
            if do_sample:
                raise NotImplementedError
            nxt = torch.argmax(logits, dim=-1, keepdim=True)
            toks.append(nxt.item())
            logits_l.append(logits.gather(-1, nxt).item())
            probs_l.append(torch.softmax(logits.float(), dim=-1).gather(-1, nxt).item())
            if toks[-1] == stop_token_idx:
                break
            x = torch.cat((x, self.rwkv.emb(nxt).to(x.dtype)), dim=-2)[:, -self.args.ctx_len:, :]
        return toks, logits_l, probs_l


def randomize_zero_init(model: nn.Module, seed: int = 1234) -> None:
    """`output`/`value`/LoRA-down weights are zero-initialised in the reference, which would
    make parity trivially pass; re-randomise them (and head/emb ~ N(0, 0.02)) with a fixed seed."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith(("att.output.weight", "ffn.value.weight")):
                p.copy_((torch.rand(p.shape, generator=g) - 0.5) * (1.0 / p.shape[1] ** 0.5))
            elif name.endswith((".w1", ".a1", ".v1", ".g1")):
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
            elif name.endswith(".r_k"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
            elif name.endswith(("head.weight", "emb.weight")):
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
