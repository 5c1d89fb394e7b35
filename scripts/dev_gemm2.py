"""Parity + timing of the CTA-pair tcgen05 GEMM (csrc/gemm2_sm100.cu), every layout, against torch / cuBLAS bf16."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from visualrwkv_b200 import fused
from visualrwkv_b200.benchutil import time_cuda

torch.manual_seed(0)
res = []
def rel(c, ref):
    return ((c.float() - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
cases = [  # (M, N, K, layout, epi, ksplit)
    (256, 128, 64, "tn", 0, 1), (512, 256, 128, "tn", 0, 1), (384, 768, 768, "tn", 0, 1), (384, 768, 768, "tn", 1, 1), (384, 768, 768, "tn", 2, 1),
    (512, 256, 128, "nn", 0, 1), (384, 768, 768, "nn", 2, 1), (256, 256, 512, "tt", 0, 1), (768, 768, 1024, "tt", 3, 4),
    (16384, 768, 768, "tn", 0, 1), (16384, 3072, 768, "tn", 1, 1), (16384, 768, 3072, "tn", 2, 1), (16384, 65536, 768, "tn", 0, 1),
    (16384, 768, 768, "nn", 0, 1), (16384, 768, 3072, "nn", 0, 1), (16384, 3072, 768, "nn", 0, 1), (16384, 768, 65536, "nn", 0, 1),
    (768, 768, 16384, "tt", 3, 8), (3072, 768, 16384, "tt", 3, 2), (768, 3072, 16384, "tt", 3, 2), (65536, 768, 16384, "tt", 0, 1),
]
for (M, N, K, lay, epi, ks) in cases:
    layout = {"tn": fused.G2_TN, "nn": fused.G2_NN, "tt": fused.G2_TT}[lay]
    a = (0.5 * torch.randn((K, M) if lay == "tt" else (M, K), device="cuda")).to(torch.bfloat16)
    b = (torch.randn((N, K) if lay == "tn" else (K, N), device="cuda") / K ** 0.5).to(torch.bfloat16)
    r = torch.randn(M, N, device="cuda").to(torch.bfloat16) if epi == 2 else None
    def run():
        return fused.gemm2(a, b, layout, 0 if epi == 3 else epi, r, ks)
    c = run()
    torch.cuda.synchronize()
    A_ = a.t() if lay == "tt" else a
    B_ = b.t() if lay == "tn" else b
    idx = torch.arange(0, M, max(1, M // 256), device="cuda")
    ref = A_[idx].float() @ B_.float()
    if epi == 1: ref = torch.relu(ref.to(torch.bfloat16).float()) ** 2
    if epi == 2: ref = ref + r[idx].float()
    out = {"M": M, "N": N, "K": K, "layout": lay, "epi": epi, "ksplit": ks, "rel_err": rel(c[idx], ref)}
    if M * N * K >= 16384 * 768 * 768:
        t_ours, _, _ = time_cuda(run, iters=8, warmup=3)
        t_ref, _, _ = time_cuda(lambda: A_ @ B_, iters=8, warmup=3)
        fl = 2.0 * M * N * K
        out.update(ours_ms=t_ours, ours_tflops=fl / t_ours / 1e9, cublas_ms=t_ref, cublas_tflops=fl / t_ref / 1e9)
    print(json.dumps(out), flush=True); res.append(out)
# grouped launches: r/k/v projections (3 x TN) and the four C x C weight gradients (4 x TT, split-K)
x3 = [(0.5 * torch.randn(16384, 768, device="cuda")).to(torch.bfloat16) for _ in range(4)]
w3 = [(torch.randn(768, 768, device="cuda") / 768 ** 0.5).to(torch.bfloat16) for _ in range(4)]
cs = fused.gemm2_grouped(x3[:3], w3[:3])
for x_, w_, c_ in zip(x3, w3, cs):
    assert rel(c_[:256], x_[:256].float() @ w_.float().t()) < 2e-3
t_g, _, _ = time_cuda(lambda: fused.gemm2_grouped(x3[:3], w3[:3]), iters=8, warmup=3)
t_c, _, _ = time_cuda(lambda: [x_ @ w_.t() for x_, w_ in zip(x3[:3], w3[:3])], iters=8, warmup=3)
print(json.dumps({"grouped_rkv_ms": t_g, "cublas_3x_ms": t_c, "tflops": 3 * 2.0 * 16384 * 768 * 768 / t_g / 1e9}))
for ks in (1, 2, 4):
    gs = fused.gemm2_grouped(x3, x3[::-1], fused.G2_TT, ksplit=ks)
    for i_, g_ in enumerate(gs):
        assert rel(g_, x3[i_].float().t() @ x3[3 - i_].float()) < 2e-3, (ks, i_)
    gs2 = fused.gemm2_grouped(x3, x3[::-1], fused.G2_TT, ksplit=ks)   # the workspace was left clean
    assert all(torch.equal(u, v) or rel(u, v.float()) < 1e-3 for u, v in zip(gs, gs2))
    t_g, _, _ = time_cuda(lambda: fused.gemm2_grouped(x3, x3[::-1], fused.G2_TT, ksplit=ks), iters=8, warmup=3)
    t_c, _, _ = time_cuda(lambda: [x_.t() @ y_ for x_, y_ in zip(x3, x3[::-1])], iters=8, warmup=3)
    print(json.dumps({"grouped_wgrad4_ksplit": ks, "ms": t_g, "cublas_4x_ms": t_c, "tflops": 4 * 2.0 * 16384 * 768 * 768 / t_g / 1e9}))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "dev_gemm2.json"), "w"), indent=1)
