"""Loads the native libraries.  There is NO fallback: if the CUDA extension is missing the import
of anything that needs it raises, loudly (tier rule: a product path must never route to a CPU or
library fallback)."""
from __future__ import annotations

import ctypes
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libvrwkv_b200.so")
SHIM_PATH = os.path.join(PKG, "libvrwkv_torch_shim.so")

_lib = None
_shim_loaded = False


class NativeLibraryMissing(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """The C-ABI library (include/vrwkv_b200.h)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryMissing(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU / PyTorch fallback for the hot path)")
        import torch  # noqa: F401  (loads the CUDA runtime libvrwkv_b200.so links against)
        _lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        _lib.vrwkv_last_error.restype = ctypes.c_char_p
    return _lib


def load_torch_ops() -> None:
    """Registers torch.ops.wind_backstepping.{forward,backward} (reference schema, wkv7_op.cpp:21-29)."""
    global _shim_loaded
    if _shim_loaded:
        return
    lib()
    if not os.path.exists(SHIM_PATH):
        raise NativeLibraryMissing(f"{SHIM_PATH} not found: run __graft_entry__.build()")
    import torch
    torch.ops.load_library(SHIM_PATH)
    _shim_loaded = True


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {lib().vrwkv_last_error().decode()}")


def ptr(t) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def cur_stream() -> ctypes.c_void_p:
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
