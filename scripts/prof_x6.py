import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import wkv7 as O
import visualrwkv_b200.wkv7 as W
B, T, H = 8, 2048, 12
w, q, k, v, a, b, dy = [x.cuda() for x in O.make_inputs(B, T, H, 64, seed=42)]
fv = int(os.environ.get("FV", "6")); bv = int(os.environ.get("BV", "0"))
W.set_variant(fv, bv)
for _ in range(3):
    y, s, sa = W.forward_raw(w, q, k, v, a, b, bounded_decay=True, chunk_checkpoints=True)
    if "--bwd" in sys.argv:
        g = W.backward_raw(w, q, k, v, a, b, dy, s, sa, bounded_decay=True)
torch.cuda.synchronize()
