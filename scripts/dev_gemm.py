"""Parity + timing of the tcgen05 GEMM (csrc/gemm_sm100.cu) against fp32 torch / cuBLAS bf16."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from visualrwkv_b200 import fused
from visualrwkv_b200.benchutil import time_cuda

torch.manual_seed(0)
res = []
for (M, N, K, epi) in [(128, 128, 64, 0), (256, 256, 128, 0), (384, 768, 768, 0), (384, 768, 768, 1), (384, 768, 768, 2),
                       (16384, 768, 768, 0), (16384, 3072, 768, 1), (16384, 768, 3072, 2), (16384, 65536, 768, 0)]:
    a = (0.5 * torch.randn(M, K, device="cuda")).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    r = torch.randn(M, N, device="cuda").to(torch.bfloat16)
    c = fused.gemm_tn(a, w, epi, r if epi == 2 else None)
    torch.cuda.synchronize()
    idx = torch.arange(0, M, max(1, M // 512), device="cuda")
    ref = a[idx].float() @ w.float().t()
    if epi == 1:
        ref = torch.relu(ref.to(torch.bfloat16).float()) ** 2
    if epi == 2:
        ref = ref + r[idx].float()
    err = ((c[idx].float() - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    out = {"M": M, "N": N, "K": K, "epi": epi, "rel_err": err}
    if M >= 16384:
        t_ours, _, _ = time_cuda(lambda: fused.gemm_tn(a, w, epi, r if epi == 2 else None), iters=8, warmup=3)
        if epi == 0:
            t_ref, _, _ = time_cuda(lambda: a @ w.t(), iters=8, warmup=3)
        elif epi == 1:
            t_ref, _, _ = time_cuda(lambda: torch.relu(a @ w.t()) ** 2, iters=8, warmup=3)
        else:
            t_ref, _, _ = time_cuda(lambda: torch.addmm(r, a, w.t()), iters=8, warmup=3)
        fl = 2.0 * M * N * K
        out.update(ours_ms=t_ours, ours_tflops=fl / t_ours / 1e9, torch_ms=t_ref, torch_tflops=fl / t_ref / 1e9)
    print(json.dumps(out)); res.append(out)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "dev_gemm.json"), "w"), indent=1)
