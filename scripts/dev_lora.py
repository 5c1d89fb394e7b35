"""Time the six grouped LoRA launches of one time-mix block against the library sequence they replace."""
import torch, torch.nn.functional as F
from visualrwkv_b200 import fused

rows, C = 16384, 768
ranks = (64, 64, 128, 32)
Rp = 128
mk = lambda *s, sc=1.0: (sc * torch.randn(*s, device="cuda")).to(torch.bfloat16)
xs = [mk(rows, C, sc=0.5) for _ in ranks]
d = [mk(C, r, sc=C ** -0.5) for r in ranks]
u = [mk(r, C, sc=r ** -0.5) for r in ranks]
dp_, up_ = d, u
douts = [mk(rows, C, sc=0.5) for _ in ranks]
acts = [1, 0, 2, 0]


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


hs = fused.gemm2_grouped(xs, dp_, fused.G2_NN, fused.EPI_ACT, acts=acts)
dps = fused.gemm2_grouped(douts, up_, fused.G2_TN, fused.EPI_ACT_BWD, residuals=hs, acts=acts)
print("down+act   %.4f" % timeit(lambda: fused.gemm2_grouped(xs, dp_, fused.G2_NN, fused.EPI_ACT, acts=acts)))
print("up         %.4f" % timeit(lambda: fused.gemm2_grouped(hs, up_, fused.G2_NN)))
print("dh+actbwd  %.4f" % timeit(lambda: fused.gemm2_grouped(douts, up_, fused.G2_TN, fused.EPI_ACT_BWD, residuals=hs, acts=acts)))
print("dx         %.4f" % timeit(lambda: fused.gemm2_grouped(dps, dp_, fused.G2_TN)))
for ks in (1, 2, 4, 8):
    print("dU ks=%d    %.4f" % (ks, timeit(lambda: fused.gemm2_grouped(hs, douts, fused.G2_TT, ksplit=ks))))
    print("dD ks=%d    %.4f" % (ks, timeit(lambda: fused.gemm2_grouped(xs, dps, fused.G2_TT, ksplit=ks))))
h0 = [None] * 4


def lib_fwd():
    for i in range(4):
        t = xs[i] @ d[i]
        h0[i] = torch.tanh(t) if acts[i] == 1 else torch.sigmoid(t) if acts[i] == 2 else t
    return [h0[i] @ u[i] for i in range(4)]


def lib_bwd():
    for i in range(4):
        dh = douts[i] @ u[i].t()
        du = h0[i].t() @ douts[i]
        if acts[i] == 1:
            dh = torch.ops.aten.tanh_backward(dh, h0[i])
        elif acts[i] == 2:
            dh = torch.ops.aten.sigmoid_backward(dh, h0[i])
        dx = dh @ d[i].t()
        dd = xs[i].t() @ dh


print("library fwd %.4f  bwd %.4f" % (timeit(lib_fwd), timeit(lib_bwd)))
for ks in (2, 4):
    print("dU^T (dout^T h) ks=%d  %.4f" % (ks, timeit(lambda: fused.gemm2_grouped(douts, hs, fused.G2_TT, ksplit=ks))))
print("dU^T ks=4 + 4 transposes  %.4f" % timeit(lambda: [t.t().contiguous() for t in fused.gemm2_grouped(douts, hs, fused.G2_TT, ksplit=4)]))
