"""The four C x C weight gradients of a time-mix block: split contraction (fp32 atomics) vs 128-wide tiles."""
import sys, torch
sys.path.insert(0, ".")
from visualrwkv_b200 import fused
rows, C = 16384, 768
mk = lambda *s: (0.5 * torch.randn(*s, device="cuda")).to(torch.bfloat16)
dys = [mk(rows, C) for _ in range(4)]
xs = [mk(rows, C) for _ in range(4)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def t(fn, n=10):
    for _ in range(2):
        fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n


ref = [d.float().t() @ x.float() for d, x in zip(dys, xs)]
rel = lambda a, b: ((a.float() - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item()
for ks in (1, 2, 4):
    outs = fused.gemm2_grouped(dys, xs, fused.G2_TT, ksplit=ks)
    ms = t(lambda: fused.gemm2_grouped(dys, xs, fused.G2_TT, ksplit=ks))
    print("4 x (C x C) wgrad  ksplit=%d  %.4f ms  %.0f TF/s  rel %.2e" % (ks, ms, 4 * 2 * rows * C * C / ms / 1e9, max(rel(o, r) for o, r in zip(outs, ref))))
print("_ksplit ->", fused._ksplit(4, C, C, rows))
