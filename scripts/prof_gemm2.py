import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visualrwkv_b200 import fused
M, N, K = 16384, 3072, 768
a = (0.5 * torch.randn(M, K, device="cuda")).to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
for _ in range(3):
    c = fused.gemm2(a, w, fused.G2_TN, 1)
torch.cuda.synchronize()
