"""tmix_mid / tmix_post forward + backward at cfg2 shapes: timing, and the target of ncu captures."""
import sys, torch
sys.path.insert(0, ".")
from visualrwkv_b200 import fused

torch.manual_seed(0)
rows, C = 16384, 768
mk = lambda *s, sc=1.0: (sc * torch.randn(*s, device="cuda")).to(torch.bfloat16)
k, v, vf, ww, aa, vv = [mk(rows, C, sc=0.5) for _ in range(6)]
w0, a0, v0, k_k, k_a = [mk(C, sc=0.3) for _ in range(5)]
gr = [mk(rows, C, sc=0.5) for _ in range(7)]
y, r, g, dz = [mk(rows, C, sc=0.5) for _ in range(4)]
gam, bet, r_k = mk(C, sc=0.1) + 1, mk(C, sc=0.1), mk(C, sc=0.3)


def run():
    w, k2, v2, nkk, kka = fused.tmix_mid_forward(k, v, vf, ww, aa, vv, w0, a0, v0, k_k, k_a)
    fused.tmix_mid_backward(k, v, vf, ww, aa, vv, w0, a0, v0, k_k, k_a, *gr)
    fused.tmix_post_forward(y, r, k2, v2, g, gam, bet, r_k, 64e-5)
    fused.tmix_post_backward(y, r, k2, v2, g, gam, bet, r_k, 64e-5, dz)
    return k2, v2


for _ in range(3):
    k2, v2 = run()
torch.cuda.synchronize()
cases = [("tmix_mid_fwd", lambda: fused.tmix_mid_forward(k, v, vf, ww, aa, vv, w0, a0, v0, k_k, k_a), 11),
         ("tmix_mid_bwd (+reduce)", lambda: fused.tmix_mid_backward(k, v, vf, ww, aa, vv, w0, a0, v0, k_k, k_a, *gr), 19),
         ("tmix_post_fwd", lambda: fused.tmix_post_forward(y, r, k2, v2, g, gam, bet, r_k, 64e-5), 6),
         ("tmix_post_bwd (+reduce)", lambda: fused.tmix_post_backward(y, r, k2, v2, g, gam, bet, r_k, 64e-5, dz), 11)]
for name, fn, ntens in cases:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("%-26s %.4f ms   %.0f GB/s of algorithmic bytes (%d tensors)" % (name, ms, ntens * rows * C * 2 / ms / 1e6, ntens))
