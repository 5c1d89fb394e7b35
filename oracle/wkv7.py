"""ctypes/numpy front-end of oracle/wkv7_oracle.c (TEST INFRASTRUCTURE ONLY).

Reference anchors: VisualRWKV-v7/v7.00/cuda/wkv7_cuda.cu:10-52 (forward), :54-130 (backward),
VisualRWKV-v6/v6.xx/RWKV-v7_simple.py:20-32 (fp64 spec), v7.00/src/model.py:45-70 (op contract).
Input generators follow SURVEY.md §8(d) / BASELINE.md §2.3 and, for the stress set,
VisualRWKV-v6/v6.xx/test_kernel.py:47-50.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
CHUNK_LEN = 16  # v7.00/src/model.py:41


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libwkv7_oracle.so")
    src = os.path.join(_HERE, "wkv7_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _f32(x) -> np.ndarray:
    """Any tensor/array -> contiguous float32 numpy (values unchanged; bf16 upcasts exactly)."""
    if isinstance(x, torch.Tensor):
        x = x.detach().float().cpu().numpy()
    return np.ascontiguousarray(x, dtype=np.float32)


def _ptr(x: np.ndarray | None):
    return None if x is None else x.ctypes.data_as(ctypes.c_void_p)


def forward(w, q, k, v, a, b, precision: str = "f64", chunk: int = CHUNK_LEN, s0=None,
            want_final_state: bool = False):
    """Returns (y, s, sa[, sT]); y/sa [B,T,H,N], s [B,H,T//chunk,N,N] (transposed checkpoints)."""
    w, q, k, v, a, b = map(_f32, (w, q, k, v, a, b))
    B, T, H, N = w.shape
    assert T % chunk == 0
    dt = np.float64 if precision == "f64" else np.float32
    y = np.empty((B, T, H, N), dt)
    sa = np.empty((B, T, H, N), dt)
    s = np.empty((B, H, T // chunk, N, N), dt)
    s0a = None if s0 is None else np.ascontiguousarray(s0, dtype=dt)
    sT = np.empty((B, H, N, N), dt) if want_final_state else None
    fn = getattr(lib(), f"wkv7_fwd_{precision}")
    fn(B, T, H, N, chunk, _ptr(w), _ptr(q), _ptr(k), _ptr(v), _ptr(a), _ptr(b), _ptr(y), _ptr(s),
       _ptr(sa), _ptr(s0a), _ptr(sT))
    return (y, s, sa, sT) if want_final_state else (y, s, sa)


def backward(w, q, k, v, a, b, dy, s, sa, precision: str = "f64", chunk: int = CHUNK_LEN):
    """Reference-algorithm backward (reverse-time state reconstruction). Returns dw,dq,dk,dv,da,db."""
    w, q, k, v, a, b, dy = map(_f32, (w, q, k, v, a, b, dy))
    B, T, H, N = w.shape
    dt = np.float64 if precision == "f64" else np.float32
    s = np.ascontiguousarray(s, dtype=dt)
    sa = np.ascontiguousarray(sa, dtype=dt)
    outs = [np.empty((B, T, H, N), dt) for _ in range(6)]
    fn = getattr(lib(), f"wkv7_bwd_{precision}")
    fn(B, T, H, N, chunk, _ptr(w), _ptr(q), _ptr(k), _ptr(v), _ptr(a), _ptr(b), _ptr(dy), _ptr(s),
       _ptr(sa), *[_ptr(o) for o in outs])
    return tuple(outs)


def backward_exact(w, q, k, v, a, b, dy):
    """Division-free fp64 adjoint (stores all states; small T only)."""
    w, q, k, v, a, b, dy = map(_f32, (w, q, k, v, a, b, dy))
    B, T, H, N = w.shape
    outs = [np.empty((B, T, H, N), np.float64) for _ in range(6)]
    lib().wkv7_bwd_exact_f64(B, T, H, N, _ptr(w), _ptr(q), _ptr(k), _ptr(v), _ptr(a), _ptr(b),
                             _ptr(dy), *[_ptr(o) for o in outs])
    return tuple(outs)


# ----------------------------------------------------------------------------------------------
# Synthetic inputs
# ----------------------------------------------------------------------------------------------
def make_inputs(B: int, T: int, H: int, N: int = 64, seed: int = 42, kind: str = "realistic",
                layer_ratio: float = 0.5):
    """bf16 CPU tensors (w, q, k, v, a, b, dy) in KERNEL order, each [B,T,H,N].

    realistic: built the way RWKV_Tmix_x070 builds them (model.py:121-124,176,183-190):
       r,k,v ~ 0.5 N(0,1); w = -softplus(-(w0+d)) - 0.5, w0 the decay ramp, d ~ N(0,0.3);
       gate ~ U(0,1); kk = normalize(0.85 k) per head; args (r, w, k(1+(gate-1)), v, -kk, kk*gate).
    stress: U(-1,1) for q,k,v,a,b and w ~ U(-8,1) -> clamped to the op's domain (<= -0.5)
       (VisualRWKV-v6/v6.xx/test_kernel.py:47-50 adapted to wkv7's w convention).
    """
    g = torch.Generator().manual_seed(seed)
    C = H * N
    if kind == "realistic":
        r = 0.5 * torch.randn(B, T, C, generator=g)
        k = 0.5 * torch.randn(B, T, C, generator=g)
        v = 0.5 * torch.randn(B, T, C, generator=g)
        n = torch.arange(C, dtype=torch.float32)
        w0 = -7 + 5 * (n / max(C - 1, 1)) ** (0.85 + 1.0 * layer_ratio ** 0.5) + 0.5
        d = 0.3 * torch.randn(B, T, C, generator=g)
        w = -torch.nn.functional.softplus(-(w0 + d)) - 0.5
        gate = torch.rand(B, T, C, generator=g)
        kk = torch.nn.functional.normalize((k * 0.85).view(B, T, H, N), dim=-1, p=2.0).view(B, T, C)
        k2 = k * (1 + (gate - 1))
        a, b = -kk, kk * gate
        q = r
        k = k2
    elif kind == "stress":
        q, k, v, a, b = [torch.rand(B, T, C, generator=g) * 2 - 1 for _ in range(5)]
        # keep the rank-1 term contractive like the model does (|a||b| <= 1 per head)
        a = -torch.nn.functional.normalize(a.view(B, T, H, N), dim=-1).view(B, T, C)
        b = torch.nn.functional.normalize(b.view(B, T, H, N), dim=-1).view(B, T, C) * \
            torch.rand(B, T, H, 1, generator=g).expand(B, T, H, N).reshape(B, T, C)
        b = -a * b.abs().clamp(max=1.0) / (a.abs() + 1e-6) * a.abs()  # b = -a * gate, gate in [0,1]
        w = (torch.rand(B, T, C, generator=g) * 9 - 8).clamp(max=-0.5)
    else:
        raise ValueError(kind)
    dy = torch.randn(B, T, C, generator=g)
    outs = [x.view(B, T, H, N).to(torch.bfloat16).contiguous() for x in (w, q, k, v, a, b, dy)]
    return tuple(outs)


# ----------------------------------------------------------------------------------------------
# Comparison utilities
# ----------------------------------------------------------------------------------------------
def to_bf16_f32(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even to bf16, returned as float32 (what the reference's to_bf store does)."""
    return torch.from_numpy(np.asarray(x, dtype=np.float32)).to(torch.bfloat16).float().numpy()


def err_ratio(x, ref) -> float:
    """RMS error ratio ||x-ref|| / ||ref|| (VisualRWKV-v6/v6.xx/test_kernel.py:27-30)."""
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.sqrt(np.mean((x - ref) ** 2)) / max(np.sqrt(np.mean(ref ** 2)), 1e-300))


def bf16_ulp_diff(x, ref) -> np.ndarray:
    """Distance in bf16 ulps between two bf16-representable float arrays."""
    def key(z):
        u = np.ascontiguousarray(z, dtype=np.float32).view(np.uint32) >> 16
        u = u.astype(np.int64)
        return np.where(u & 0x8000, 0x8000 - u, u)
    return np.abs(key(x) - key(ref))
