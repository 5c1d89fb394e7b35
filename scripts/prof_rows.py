"""One time-mix + channel-mix block forward/backward at cfg2 shapes, for ncu captures of the row kernels."""
import sys, torch
sys.path.insert(0, ".")
from visualrwkv_b200 import fused

torch.manual_seed(0)
rows, T, C = 16384, 2048, 768
mk = lambda *s, sc=1.0: (sc * torch.randn(*s, device="cuda")).to(torch.bfloat16)
x = mk(rows, C)
g, b = mk(C, sc=0.1) + 1, mk(C, sc=0.1)
coefs6 = [torch.rand(C, device="cuda").to(torch.bfloat16) for _ in range(6)]
for it in range(3):
    outs, _, stats = fused.ln_mix_forward(x, T, g, b, 1e-5, coefs6)
    o1, _, st1 = fused.ln_mix_forward(x, T, g, b, 1e-5, coefs6[:1])
    douts = [mk(rows, C) for _ in range(6)]
    dres = mk(rows, C)
    fused.ln_mix_backward(x, T, stats, g, b, coefs6, douts, dresid=dres)
    fused.ln_mix_backward(x, T, st1, g, b, coefs6[:1], douts[:1], dresid=dres)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
ev[0].record()
for _ in range(20):
    fused.ln_mix_forward(x, T, g, b, 1e-5, coefs6)
ev[1].record()
for _ in range(20):
    fused.ln_mix_forward(x, T, g, b, 1e-5, coefs6[:1])
ev[2].record()
for _ in range(20):
    fused.ln_mix_backward(x, T, stats, g, b, coefs6, douts, dresid=dres)
ev[3].record()
for _ in range(20):
    fused.ln_mix_backward(x, T, st1, g, b, coefs6[:1], douts[:1], dresid=dres)
ev[4].record()
torch.cuda.synchronize()
names = ["ln_mix_fwd<6>", "ln_mix_fwd<1>", "ln_mix_bwd<6> (2 passes + reduce)", "ln_mix_bwd<1> (2 passes + reduce)"]
byts = [rows * C * 2 * 7, rows * C * 2 * 2, rows * C * 2 * 9, rows * C * 2 * 4]
for i in range(4):
    ms = ev[i].elapsed_time(ev[i + 1]) / 20
    print("%-36s %.4f ms   %.0f GB/s of algorithmic bytes" % (names[i], ms, byts[i] / ms / 1e6))
