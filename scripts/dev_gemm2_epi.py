"""Isolated timings of the epilogue variants of the CTA-pair GEMM at the channel-mix shapes."""
import sys, torch
sys.path.insert(0, ".")
from visualrwkv_b200 import fused

M, C, Hd = 16384, 768, 3072
mk = lambda *s, sc=1.0: (sc * torch.randn(*s, device="cuda")).to(torch.bfloat16)
x, Wk, Wv = mk(M, C, sc=0.5), mk(Hd, C, sc=0.03), mk(C, Hd, sc=0.02)
act = mk(M, Hd).abs()
dy = mk(M, C, sc=0.5)
res = mk(M, C)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def t(fn, n=10):
    for _ in range(2):
        fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n


fl = 2 * M * C * Hd
for name, fn in [("key fwd TN relu^2        ", lambda: fused.gemm2(x, Wk, fused.G2_TN, fused.EPI_RELU_SQ)),
                 ("key fwd TN plain         ", lambda: fused.gemm2(x, Wk, fused.G2_TN)),
                 ("dhk NN plain (K=768)     ", lambda: fused.gemm2(dy, Wv, fused.G2_NN)),
                 ("dhk NN relusq_bwd        ", lambda: fused.gemm2(dy, Wv, fused.G2_NN, fused.EPI_RELUSQ_BWD, act)),
                 ("dhk NN + residual (ADD)  ", lambda: fused.gemm2(dy, Wv, fused.G2_NN, fused.EPI_ADD, act)),
                 ("value fwd TN + residual  ", lambda: fused.gemm2(act, Wv, fused.G2_TN, fused.EPI_ADD, res)),
                 ("value fwd TN plain       ", lambda: fused.gemm2(act, Wv, fused.G2_TN)),
                 ("dxk NN (K=3072) plain    ", lambda: fused.gemm2(act, Wk, fused.G2_NN))]:
    ms = t(fn)
    print("%s %.4f ms  %.0f TF/s" % (name, ms, fl / ms / 1e9))
