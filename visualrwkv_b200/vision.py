"""SigLIP vision tower (the v7.01 variant of the reference: HF `SiglipVisionModel(...).last_hidden_state`,
VisualRWKV-v7/v7.01/src/model.py:347-352,448-454; arithmetic defined by transformers'
modeling_siglip.py).  Parameter names follow the HF checkpoint layout
(`vision_model.embeddings.*`, `vision_model.encoder.layers.N.*`, `vision_model.post_layernorm.*`) so a
real SigLIP checkpoint loads with load_state_dict.  The attention-pooling head, whose output the
reference discards, is not evaluated.
"""
from __future__ import annotations

import json
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

SIGLIP_CONFIGS = {
    # name: hidden, layers, heads, mlp, image, patch, eps
    "siglip-base-patch16-224": dict(hidden=768, layers=12, heads=12, mlp=3072, image=224, patch=16, eps=1e-6),
    "siglip-base-patch16-256": dict(hidden=768, layers=12, heads=12, mlp=3072, image=256, patch=16, eps=1e-6),
    "siglip-base-patch16-384": dict(hidden=768, layers=12, heads=12, mlp=3072, image=384, patch=16, eps=1e-6),
    "siglip-tiny-test": dict(hidden=128, layers=2, heads=2, mlp=256, image=64, patch=16, eps=1e-6),
}


def _config_from_hf(d: dict) -> dict:
    """HF config.json (SiglipVisionConfig, or the `vision_config` of a full SiglipConfig) -> our keys."""
    d = d.get("vision_config", d)
    return dict(hidden=int(d.get("hidden_size", 768)), layers=int(d.get("num_hidden_layers", 12)),
                heads=int(d.get("num_attention_heads", 12)), mlp=int(d.get("intermediate_size", 3072)),
                image=int(d.get("image_size", 224)), patch=int(d.get("patch_size", 16)),
                eps=float(d.get("layer_norm_eps", 1e-6)))


def resolve_config(name_or_path: str) -> dict:
    """A checkpoint directory (config.json wins) or a known tower name.  Names match on their tail with `siglip2-`
    folded onto `siglip-` (same architecture; the reference's scripts pass e.g. google/siglip2-base-patch16-256,
    VisualRWKV-v7/v7.01/scripts)."""
    cfg_json = os.path.join(name_or_path, "config.json")
    if os.path.isfile(cfg_json):
        with open(cfg_json) as f:
            return _config_from_hf(json.load(f))
    key = name_or_path.rstrip("/").split("/")[-1].replace("siglip2-", "siglip-")
    for k, v in SIGLIP_CONFIGS.items():
        if key.endswith(k):
            return dict(v)
    raise ValueError(f"unknown SigLIP tower {name_or_path!r}; known: {sorted(SIGLIP_CONFIGS)} or a directory with config.json")


def _checkpoint_tensors(path: str) -> dict | None:
    """Weights of a HF checkpoint directory (model.safetensors / sharded *.safetensors / pytorch_model.bin), or None."""
    if not os.path.isdir(path):
        return None
    st = sorted(f for f in os.listdir(path) if f.endswith(".safetensors"))
    if st:
        from safetensors.torch import load_file
        out = {}
        for f in st:
            out.update(load_file(os.path.join(path, f)))
        return out
    pt = os.path.join(path, "pytorch_model.bin")
    if os.path.isfile(pt):
        return torch.load(pt, map_location="cpu", weights_only=True)
    return None


class _Embeddings(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.patch_embedding = nn.Conv2d(3, c["hidden"], kernel_size=c["patch"], stride=c["patch"], padding="valid")
        n = (c["image"] // c["patch"]) ** 2
        self.position_embedding = nn.Embedding(n, c["hidden"])


class _Attn(nn.Module):
    def __init__(self, c):
        super().__init__()
        D = c["hidden"]
        self.q_proj, self.k_proj, self.v_proj, self.out_proj = (nn.Linear(D, D) for _ in range(4))


class _MLP(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.fc1 = nn.Linear(c["hidden"], c["mlp"])
        self.fc2 = nn.Linear(c["mlp"], c["hidden"])


class _Layer(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.layer_norm1 = nn.LayerNorm(c["hidden"], eps=c["eps"])
        self.self_attn = _Attn(c)
        self.layer_norm2 = nn.LayerNorm(c["hidden"], eps=c["eps"])
        self.mlp = _MLP(c)


class _Encoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.layers = nn.ModuleList([_Layer(c) for _ in range(c["layers"])])


class _VisionModel(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.embeddings = _Embeddings(c)
        self.encoder = _Encoder(c)
        self.post_layernorm = nn.LayerNorm(c["hidden"], eps=c["eps"])


class SiglipVisionTower(nn.Module):
    def __init__(self, name_or_path: str):
        super().__init__()
        self.cfg = resolve_config(name_or_path)
        self.embed_dim = self.cfg["hidden"]
        self.vision_model = _VisionModel(self.cfg)
        # the reference calls SiglipVisionModel.from_pretrained(vision_tower_path) (v7.01/src/model.py:347-350): when the
        # path is a checkpoint directory its weights are loaded; a bare name (no files offline) keeps the random init
        tensors = _checkpoint_tensors(name_or_path)
        self.pretrained = tensors is not None
        if tensors is not None:
            self.load_state_dict(tensors)

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        """Accepts HF SiglipVisionModel / SiglipModel checkpoints: the attention-pooling head (`vision_model.head.*`,
        whose output the reference discards) and the text tower are dropped before the strict check."""
        sd = {}
        for k, v in state_dict.items():
            if k.startswith("text_model.") or k in ("logit_scale", "logit_bias"):
                continue
            if ".head." in k and k.split(".head.")[0].endswith("vision_model"):
                continue
            sd[k] = v
        return super().load_state_dict(sd, strict=strict, assign=assign)

    @torch.no_grad()
    def forward(self, pixels):
        """[N,3,H,W] -> last_hidden_state [N,(H/p)*(W/p),D]  (frozen tower: no autograd graph)."""
        from . import ops_vit
        return ops_vit.siglip_forward(self, pixels)
