"""ctypes driver for the reference WKV7 CUDA kernel built by oracle/build_ref.py
(TEST INFRASTRUCTURE ONLY; GPU only).

Calls the reference launchers `cuda_forward` / `cuda_backward`
(VisualRWKV-v7/v7.00/cuda/wkv7_cuda.cu:132-138) with caller-allocated torch tensors, exactly the
way WindBackstepping does (v7.00/src/model.py:47-65).  The launchers use the legacy default
stream, so callers must be on torch's default stream.
"""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libwkv7_ref.so")
_lib = None
_fwd = _bwd = None


def available() -> bool:
    return os.path.exists(SO)


def _load():
    global _lib, _fwd, _bwd
    if _lib is None:
        _lib = ctypes.CDLL(SO)
        syms = subprocess.check_output(["nm", "-D", "--defined-only", SO], text=True).split()
        f = [s for s in syms if s.startswith("_Z12cuda_forward")]
        b = [s for s in syms if s.startswith("_Z13cuda_backward")]
        assert len(f) == 1 and len(b) == 1, (f, b)
        _fwd, _bwd = getattr(_lib, f[0]), getattr(_lib, b[0])
        _fwd.restype = _bwd.restype = None
    return _fwd, _bwd


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def forward(w, q, k, v, a, b):
    """Reference forward: returns (y bf16, s f32 [B,H,T/16,64,64], sa f32 [B,T,H,64])."""
    fwd, _ = _load()
    B, T, H, C = w.shape
    assert C == 64 and T % 16 == 0
    assert all(i.dtype == torch.bfloat16 and i.is_contiguous() for i in (w, q, k, v, a, b))
    y = torch.empty_like(v)
    s = torch.empty(B, H, T // 16, C, C, dtype=torch.float32, device=w.device)
    sa = torch.empty(B, T, H, C, dtype=torch.float32, device=w.device)
    fwd(B, T, H, _p(w), _p(q), _p(k), _p(v), _p(a), _p(b), _p(y), _p(s), _p(sa))
    return y, s, sa


def backward(w, q, k, v, a, b, dy, s, sa):
    _, bwd = _load()
    B, T, H, C = w.shape
    outs = [torch.empty_like(x) for x in (w, q, k, v, a, b)]
    bwd(B, T, H, _p(w), _p(q), _p(k), _p(v), _p(a), _p(b), _p(dy), _p(s), _p(sa),
        *[_p(o) for o in outs])
    return tuple(outs)
