"""Per-phase clock stamps of chunk 1 (CTA 0,0) of the chunked WKV7 forward."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import wkv7 as O  # noqa: E402
import visualrwkv_b200.wkv7 as W  # noqa: E402
from visualrwkv_b200 import _lib  # noqa: E402

_lib.load_torch_ops()
dbg = torch.zeros(4096, dtype=torch.float32, device="cuda")
_lib.check(_lib.lib().vrwkv_wkv7_chunk_debug(ctypes.c_void_p(dbg.data_ptr())), "dbg")
W.set_variant(3, 1)
T = 128
cpu = list(O.make_inputs(1, T, 1, 64, seed=3))[:6]
w, q, k, v, a, b = [x.cuda() for x in cpu]
y = torch.zeros_like(v)
s = torch.zeros(1, 1, T // 16, 64, 64, dtype=torch.float32, device="cuda")
sa = torch.zeros(1, T, 1, 64, dtype=torch.float32, device="cuda")
torch.ops.wind_backstepping.forward(w, q, k, v, a, b, y, s, sa)
torch.cuda.synchronize()
ts = [int(x) for x in dbg.cpu().numpy()[2048:2048 + 40]]
print("stamps:", ts)
print("deltas:", [ts[i + 1] - ts[i] for i in range(len(ts) - 1) if ts[i + 1] > 0])
dbg.zero_()
W.wkv7_forward_state(w, q, k, v, a, b)
torch.cuda.synchronize()
ts = [int(x) for x in dbg.cpu().numpy()[2048:2048 + 40]]
print("no-ckpt deltas:", [ts[i + 1] - ts[i] for i in range(len(ts) - 1) if ts[i + 1] > 0])
_lib.check(_lib.lib().vrwkv_wkv7_chunk_debug(ctypes.c_void_p(0)), "dbg")

# ---- chunked backward: clock stamps of CTA (0,0,chunk 1) ----
_lib.check(_lib.lib().vrwkv_wkv7_chunk_debug(ctypes.c_void_p(dbg.data_ptr())), "dbg")
dbg.zero_()
cpu = list(O.make_inputs(1, T, 1, 64, seed=3))
w, q, k, v, a, b, dy = [x.cuda() for x in cpu]
W.set_variant(3, 5)
torch.ops.wind_backstepping.forward(w, q, k, v, a, b, y, s, sa)
g = [torch.empty_like(w) for _ in range(6)]
torch.ops.wind_backstepping.backward(w, q, k, v, a, b, dy, s, sa, *g)
torch.cuda.synchronize()
ts = [int(x) for x in dbg.cpu().numpy()[3072:3072 + 48]]
print("bwd stamps:", ts)
print("bwd deltas:", [ts[i + 1] - ts[i] for i in range(len(ts) - 1) if ts[i + 1] > 0])
_lib.check(_lib.lib().vrwkv_wkv7_chunk_debug(ctypes.c_void_p(0)), "dbg")
