import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import wkv7 as O
from visualrwkv_b200 import wkv7 as W
B, T, H = [int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (1, 64, 1))]
ck = len(sys.argv) > 4 and sys.argv[4] == "ck64"
cpu = O.make_inputs(B, T, H, 64, seed=3)
w, q, k, v, a, b, dy = [x.cuda() for x in cpu]
W.set_variant(6, 0)
y, s, sa = W.forward_raw(w, q, k, v, a, b, bounded_decay=True, chunk_checkpoints=ck)
torch.cuda.synchronize()
y64, s64, sa64 = O.forward(*cpu[:6])
if ck: s64 = s64[:, :, 3::4]
print("sa err", O.err_ratio(sa.cpu().numpy(), sa64), "s err", O.err_ratio(s.cpu().numpy(), s64), "y err", O.err_ratio(y.float().cpu().numpy(), y64))
