"""gemm2 with the TMA traffic or the MMAs switched off (VRWKV_GEMM2_DBG=1/2): which side bounds the kernel."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visualrwkv_b200 import fused
from visualrwkv_b200.benchutil import time_cuda
for (M, N, K, lay) in [(16384, 3072, 768, "tn"), (16384, 768, 3072, "tn"), (16384, 65536, 768, "tn"), (16384, 768, 65536, "nn")]:
    a = (0.5 * torch.randn(M, K, device="cuda")).to(torch.bfloat16)
    b = (torch.randn((N, K) if lay == "tn" else (K, N), device="cuda") / K ** 0.5).to(torch.bfloat16)
    t, _, _ = time_cuda(lambda: fused.gemm2(a, b, fused.G2_TN if lay == "tn" else fused.G2_NN), iters=8, warmup=3)
    print(f"mode {os.environ.get('VRWKV_GEMM2_DBG','0')} {M}x{N}x{K} {lay}: {t:.4f} ms  {2.0*M*N*K/t/1e9:.0f} TF/s-equivalent", flush=True)
