"""Times every fused row-wise kernel at the cfg2 shape (B8 T2048 C768) and prints achieved GB/s against the
algorithmic bytes of each (dev harness; run on the GPU box)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from visualrwkv_b200 import fused  # noqa: E402
from visualrwkv_b200.benchutil import time_cuda  # noqa: E402

B, T, C, H = 8, 2048, 768, 12
rows = B * T
dev = "cuda"
torch.manual_seed(0)
bf = lambda *s: (0.5 * torch.randn(*s, device=dev)).to(torch.bfloat16)
x = bf(rows, C)
g, b = bf(C) + 1, bf(C)
coefs = [torch.rand(C, device=dev).to(torch.bfloat16) for _ in range(6)]
MB = rows * C * 2 / 1e6
res = {}


def t(name, fn, nbytes_mb):
    med, mn, _ = time_cuda(fn, iters=10, warmup=3)
    res[name] = {"ms": round(med, 4), "GBs": round(nbytes_mb / med, 1)}
    print(name, res[name])


outs, _, stats = fused.ln_mix_forward(x, T, g, b, 1e-5, coefs)
t("ln_mix_fwd6", lambda: fused.ln_mix_forward(x, T, g, b, 1e-5, coefs), 7 * MB)
t("ln_mix_fwd1", lambda: fused.ln_mix_forward(x, T, g, b, 1e-5, coefs[:1]), 2 * MB)
t("ln_fwd0", lambda: fused.ln_mix_forward(x, T, g, b, 1e-5, [], want_h=True), 2 * MB)
douts = [bf(rows, C) for _ in range(6)]
t("ln_mix_bwd6", lambda: fused.ln_mix_backward(x, T, stats, g, b, coefs, douts, dresid=x), 9 * MB)
t("ln_mix_bwd1", lambda: fused.ln_mix_backward(x, T, stats, g, b, coefs[:1], douts[:1], dresid=x), 4 * MB)
t("ln_bwd0", lambda: fused.ln_mix_backward(x, T, stats, g, b, [], [], dh=douts[0]), 3 * MB)
k, v, vf, ww, aa, vv = [bf(rows, C) for _ in range(6)]
pr = [bf(C) for _ in range(5)]
t("tmix_mid_fwd", lambda: fused.tmix_mid_forward(k, v, vf, ww, aa, vv, *pr), 11 * MB)
gr = [bf(rows, C) for _ in range(5)]
t("tmix_mid_bwd", lambda: fused.tmix_mid_backward(k, v, vf, ww, aa, vv, *pr, *gr), 17 * MB)
y, r, k2, v2, gg = [bf(rows, C) for _ in range(5)]
t("tmix_post_fwd", lambda: fused.tmix_post_forward(y, r, k2, v2, gg, g, b, pr[0], 64e-5), 6 * MB)
t("tmix_post_bwd", lambda: fused.tmix_post_backward(y, r, k2, v2, gg, g, b, pr[0], 64e-5, gr[0]), 11 * MB)
hk = bf(rows, 4 * C)
t("relu_sq_fwd", lambda: fused.relu_sq_forward(hk), 8 * MB)
t("relu_sq_bwd", lambda: fused.relu_sq_backward(hk, hk), 12 * MB)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "dev_fused.json"), "w"), indent=1)
