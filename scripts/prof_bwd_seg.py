"""Kernel-level times of the segmented backward (variant 3) at the cfg2 shape, via torch.profiler."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import wkv7 as O  # noqa: E402
import visualrwkv_b200.wkv7 as W  # noqa: E402
from visualrwkv_b200 import _lib  # noqa: E402

_lib.load_torch_ops()
B, T, H = 8, 2048, 12
w, q, k, v, a, b, dy = [x.cuda() for x in O.make_inputs(B, T, H, 64, seed=42)]
y = torch.empty_like(v)
s = torch.empty(B, H, T // 16, 64, 64, dtype=torch.float32, device="cuda")
sa = torch.empty(B, T, H, 64, dtype=torch.float32, device="cuda")
g = [torch.empty_like(w) for _ in range(6)]
for fv, bv in ((3, 3), (3, 5)):
    W.set_variant(fv, bv)
    for _ in range(3):
        torch.ops.wind_backstepping.forward(w, q, k, v, a, b, y, s, sa)
        torch.ops.wind_backstepping.backward(w, q, k, v, a, b, dy, s, sa, *g)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            torch.ops.wind_backstepping.forward(w, q, k, v, a, b, y, s, sa)
            torch.ops.wind_backstepping.backward(w, q, k, v, a, b, dy, s, sa, *g)
        torch.cuda.synchronize()
    print(f"--- fv={fv} bv={bv}")
    for e in prof.key_averages():
        if e.device_time_total > 0:
            print(f"{e.key[:60]:60s} n={e.count:3d} avg_us={e.device_time_total / e.count:9.1f}")
