"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/vrwkv_b200.h declares, argument validation fails loudly before touching a GPU, and the torch shim
registers the reference's schemas verbatim (VisualRWKV-v7/v7.00/cuda/wkv7_op.cpp:21-29)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from visualrwkv_b200 import _build, _lib
    _build.build_all()
    return _lib.lib()


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "vrwkv_b200.h")).read()
    names = set(re.findall(r"\b(vrwkv_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 6
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/vrwkv_b200.h but not exported"


def test_validation_fails_loudly_without_gpu(lib):
    lib.vrwkv_last_error.restype = ctypes.c_char_p
    fake = ctypes.c_void_p(0x1000)
    args = [fake] * 9 + [ctypes.c_void_p(0)]
    assert lib.vrwkv_wkv7_forward(1, 24, 1, *args) == -1  # T % 16 != 0
    assert b"multiple of 16" in lib.vrwkv_last_error()
    assert lib.vrwkv_wkv7_forward(0, 16, 1, *args) == -1
    bad = [ctypes.c_void_p(0x1004)] + [fake] * 8 + [ctypes.c_void_p(0)]
    assert lib.vrwkv_wkv7_forward(1, 16, 1, *bad) == -1  # misaligned
    assert b"aligned" in lib.vrwkv_last_error()
    nul = [ctypes.c_void_p(0)] + [fake] * 8 + [ctypes.c_void_p(0)]
    assert lib.vrwkv_wkv7_forward(1, 16, 1, *nul) == -1
    bargs = [fake] * 15 + [ctypes.c_void_p(0)]
    assert lib.vrwkv_wkv7_backward(1, 20, 1, *bargs) == -1


def test_torch_schema_is_verbatim(lib):
    from visualrwkv_b200 import _lib
    _lib.load_torch_ops()
    f = str(torch.ops.wind_backstepping.forward.default._schema)
    b = str(torch.ops.wind_backstepping.backward.default._schema)
    assert f == ("wind_backstepping::forward(Tensor w, Tensor q, Tensor k, Tensor v, Tensor z, Tensor a, "
                 "Tensor(a!) y, Tensor(b!) s, Tensor(c!) sa) -> ()")
    assert b == ("wind_backstepping::backward(Tensor w, Tensor q, Tensor k, Tensor v, Tensor z, Tensor a, Tensor dy, "
                 "Tensor s, Tensor sa, Tensor(a!) dw, Tensor(b!) dq, Tensor(c!) dk, Tensor(d!) dv, Tensor(e!) dz, "
                 "Tensor(f!) da) -> ()")


def test_cpu_tensors_are_rejected_not_silently_computed(lib):
    from visualrwkv_b200 import _lib
    _lib.load_torch_ops()
    w = torch.zeros(1, 16, 1, 64, dtype=torch.bfloat16)
    with pytest.raises((RuntimeError, NotImplementedError)):
        torch.ops.wind_backstepping.forward(w, w, w, w, w, w, w, torch.zeros(1), torch.zeros(1))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "visualrwkv_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"
                assert "oracle/" not in src or f.endswith(".md"), f"{f} references oracle/"
