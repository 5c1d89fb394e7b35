"""One forward + one backward of the WKV7 op at the cfg2 shape (for ncu captures)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import wkv7 as O
from visualrwkv_b200 import wkv7 as W, _lib
_lib.load_torch_ops()
fv = int(sys.argv[1]) if len(sys.argv) > 1 else 0
bv = int(sys.argv[2]) if len(sys.argv) > 2 else 0
B, T, H = 8, 2048, 12
w, q, k, v, a, b, dy = [x.cuda() for x in O.make_inputs(B, T, H, 64, seed=42)]
y = torch.empty_like(v)
s = torch.empty(B, H, T // 16, 64, 64, dtype=torch.float32, device="cuda")
sa = torch.empty(B, T, H, 64, dtype=torch.float32, device="cuda")
g = [torch.empty_like(w) for _ in range(6)]
if fv == 0 and bv == 0:  # the model path: bounded decay, chunk-granularity checkpoints
    for _ in range(3):
        y, s, sa = W.forward_raw(w, q, k, v, a, b, bounded_decay=True, chunk_checkpoints=True)
        g = W.backward_raw(w, q, k, v, a, b, dy, s, sa, bounded_decay=True)
else:
    W.set_variant(fv, bv)
    for _ in range(3):
        torch.ops.wind_backstepping.forward(w, q, k, v, a, b, y, s, sa)
        torch.ops.wind_backstepping.backward(w, q, k, v, a, b, dy, s, sa, *g)
torch.cuda.synchronize()
