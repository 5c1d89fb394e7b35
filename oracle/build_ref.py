"""Compile the UNMODIFIED reference WKV7 kernel for sm_100a (TEST INFRASTRUCTURE ONLY).

Sources are compiled where they lie under /root/reference (never copied into this repo):
    VisualRWKV-v7/v7.00/cuda/wkv7_cuda.cu        (forward_kernel / backward_kernel + launchers)
Flags exactly as VisualRWKV-v7/v7.00/src/model.py:42 plus the sm_100a gencode.  The torch binding
(wkv7_op.cpp) is deliberately NOT built: it would register TORCH_LIBRARY(wind_backstepping), the
very namespace the product registers; the two launchers `cuda_forward/cuda_backward`
(wkv7_cuda.cu:132-138) are called through ctypes by oracle/ref_kernel.py instead.

Output: oracle/_ref/libwkv7_ref.so  (git-ignored, travels to the GPU box with gpurun).
/root/reference does not exist on the GPU box, so this only runs in the build container.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CU = "/root/reference/VisualRWKV-v7/v7.00/cuda/wkv7_cuda.cu"
OUT = os.path.join(HERE, "_ref", "libwkv7_ref.so")


def build(force: bool = False) -> str | None:
    if not os.path.exists(REF_CU):
        return OUT if os.path.exists(OUT) else None
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(REF_CU):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["nvcc", "-shared", "-Xcompiler", "-fPIC", "-D_C_=64", "-D_CHUNK_LEN_=16",
           "--use_fast_math", "-O3", "-Xptxas", "-O3", "--extra-device-vectorization",
           "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-o", OUT, REF_CU]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
