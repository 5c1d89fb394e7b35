"""In-tree build of the native code (no JIT cache: the .so files travel to the GPU box with gpurun).

  visualrwkv_b200/libvrwkv_b200.so        nvcc, sm_100a only: every csrc/*.cu behind the C ABI
                                          declared in include/vrwkv_b200.h (links cudart only)
  visualrwkv_b200/libvrwkv_torch_shim.so  g++: TORCH_LIBRARY(wind_backstepping) on top of the C ABI
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libvrwkv_b200.so")
SHIM = os.path.join(PKG, "libvrwkv_torch_shim.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
if os.environ.get("VRWKV_PHASE_STAMPS"):   # development builds: per-phase clock stamps in the x6 / x3 WKV7 kernels
    NVCC_FLAGS.append("-DVRWKV_PHASE_STAMPS")
CXX = "/usr/bin/g++"


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force: bool = False, verbose: bool = False) -> str:
    cus = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = cus + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(ROOT, "include", "*.h"))
    if force or _newer(LIB, deps):
        objs = []
        procs = []
        os.makedirs(os.path.join(PKG, "build"), exist_ok=True)
        for cu in cus:
            obj = os.path.join(PKG, "build", os.path.basename(cu)[:-3] + ".o")
            objs.append(obj)
            if force or _newer(obj, deps):
                cmd = ["nvcc", *NVCC_FLAGS, "-c", cu, "-o", obj] + (["-Xptxas", "-v"] if verbose else [])
                procs.append((cu, subprocess.Popen(cmd)))
        for cu, p in procs:
            if p.wait() != 0:
                raise RuntimeError(f"nvcc failed on {cu}")
        # shared cudart: the library must share torch's CUDA runtime instance (same per-thread device/context
        # state, e.g. on autograd worker threads); torch loads libcudart.so.12 before we are dlopen'ed
        subprocess.check_call(["nvcc", "-shared", "-cudart", "shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"])
    return LIB


def build_shim(force: bool = False) -> str:
    src = os.path.join(CSRC, "torch_shim.cpp")
    if force or _newer(SHIM, [src, LIB, os.path.join(ROOT, "include", "vrwkv_b200.h")]):
        import torch
        from torch.utils import cpp_extension as ce
        inc = [f"-I{p}" for p in ce.include_paths("cuda")]
        tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
        cmd = [CXX, "-O2", "-std=c++17", "-fPIC", "-shared", src, "-o", SHIM, *inc,
               f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
               f"-L{tlib}", "-ltorch", "-ltorch_cpu", "-lc10", "-lc10_cuda", "-ltorch_cuda",
               f"-L{PKG}", "-lvrwkv_b200", "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{tlib}"]
        subprocess.check_call(cmd)
    return SHIM


def build_all(force: bool = False, verbose: bool = False):
    return build_lib(force, verbose), build_shim(force)


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
