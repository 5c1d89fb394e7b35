"""GPU dev harness for the tensor-core chunked WKV7 forward (variant 3): staged parity checks + timing."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import wkv7 as O  # noqa: E402
import visualrwkv_b200.wkv7 as W  # noqa: E402
from visualrwkv_b200.benchutil import time_cuda  # noqa: E402
from visualrwkv_b200 import _lib  # noqa: E402

_lib.load_torch_ops()


def fwd(inp, fv):
    w, q, k, v, a, b = inp
    B, T, H, C = w.shape
    W.set_variant(fv, 1)
    y = torch.zeros_like(v)
    s = torch.zeros(B, H, T // 16, C, C, dtype=torch.float32, device=w.device)
    sa = torch.zeros(B, T, H, C, dtype=torch.float32, device=w.device)
    torch.ops.wind_backstepping.forward(w, q, k, v, a, b, y, s, sa)
    torch.cuda.synchronize()
    return y.float().cpu().numpy(), s.cpu().numpy(), sa.cpu().numpy()


def er(x, r):
    d = float(np.sqrt(((x.astype(np.float64) - r) ** 2).mean()))
    n = float(np.sqrt((r.astype(np.float64) ** 2).mean()))
    return d / n if n > 0 else d


def case(shape, mode, seed=3, kind="realistic"):
    B, T, H = shape
    cpu = list(O.make_inputs(B, T, H, 64, seed=seed, kind=kind))[:6]
    if mode == "a0":   # no rank-1 term: y = tril(Qt Kt^T) V + Qt S0^T, state = V^T Kt
        cpu[4] = torch.zeros_like(cpu[4])
    if mode == "b0":   # sa path alive (a != 0) but it never feeds back
        cpu[5] = torch.zeros_like(cpu[5])
    y64, s64, sa64 = O.forward(*cpu)
    res = {"shape": shape, "mode": mode, "kind": kind}
    for fv in (1, 3):
        y, s, sa = fwd([x.cuda() for x in cpu], fv)
        res[f"v{fv}"] = {"y": er(y, y64), "sa": er(sa, sa64), "s": er(s, s64),
                         "nan": int(np.isnan(y).sum() + np.isnan(s).sum() + np.isnan(sa).sum())}
        if fv == 3:
            L = 64
            res["v3_chunks"] = [{"y": er(y[:, c * L:(c + 1) * L], y64[:, c * L:(c + 1) * L]),
                                 "sa": er(sa[:, c * L:(c + 1) * L], sa64[:, c * L:(c + 1) * L]),
                                 "s": [er(s[:, :, 4 * c + g], s64[:, :, 4 * c + g]) for g in range(4)]}
                                for c in range(min(T // L, 3))]
    return res


def timing(shape):
    B, T, H = shape
    w, q, k, v, a, b = [x.cuda() for x in list(O.make_inputs(B, T, H, 64, seed=42))[:6]]
    y = torch.empty_like(v)
    s = torch.empty(B, H, T // 16, 64, 64, dtype=torch.float32, device="cuda")
    sa = torch.empty(B, T, H, 64, dtype=torch.float32, device="cuda")
    out = {"shape": shape}
    for fv in (1, 3):
        W.set_variant(fv, 1)
        med, mn, _ = time_cuda(lambda: torch.ops.wind_backstepping.forward(w, q, k, v, a, b, y, s, sa), iters=20, warmup=5)
        out[f"fwd_v{fv}_ms"] = float(med)
        out[f"fwd_v{fv}_GBs"] = 14 * B * T * H * 64 / float(med) / 1e6
        med, mn, _ = time_cuda(lambda: W.wkv7_forward_state(w, q, k, v, a, b), iters=20, warmup=5)
        out[f"fwd_v{fv}_nockpt_ms"] = float(med)
    return out


def split_state(fv):
    """y(T) in one call == two calls chained through state_out/state_in."""
    w, q, k, v, a, b = [x.cuda() for x in list(O.make_inputs(2, 256, 2, 64, seed=5))[:6]]
    W.set_variant(fv, 1)
    y, st = W.wkv7_forward_state(w, q, k, v, a, b)
    h = 128
    y1, s1 = W.wkv7_forward_state(*[x[:, :h].contiguous() for x in (w, q, k, v, a, b)])
    y2, s2 = W.wkv7_forward_state(*[x[:, h:].contiguous() for x in (w, q, k, v, a, b)], state_in=s1)
    torch.cuda.synchronize()
    yy = torch.cat([y1, y2], dim=1)
    return {"fv": fv, "y_mismatch": float((yy != y).float().mean()), "y_maxdiff": float((yy.float() - y.float()).abs().max()),
            "state_rel": float(((s2 - st).norm() / st.norm()))}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    cases = [((1, 64, 1), "a0"), ((1, 128, 1), "a0"), ((1, 64, 1), "b0"), ((1, 64, 1), "full"), ((1, 128, 2), "full")]
    if not args.quick:
        cases += [((2, 512, 3), "full"), ((1, 2048, 2), "full")]
    for shape, mode in cases:
        print(json.dumps(case(shape, mode)), flush=True)
    if not args.quick:
        print(json.dumps(case((2, 256, 2), "full", kind="stress")), flush=True)
    print(json.dumps(split_state(3)), flush=True)
    if args.time:
        print(json.dumps(timing((8, 2048, 12))), flush=True)
