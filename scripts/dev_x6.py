"""Development harness for the round-2 x6 WKV7 kernels: parity vs the fp64 oracle (strict element-wise bounds) and timing
at cfg2.  Usage: python scripts/dev_x6.py [--time] [--bwd]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import wkv7 as O  # noqa: E402
from visualrwkv_b200 import wkv7 as W  # noqa: E402

NAMES = ["dw", "dq", "dk", "dv", "da", "db"]
out = {}


def check(shape, kind, seed, fv, bv, ck, do_bwd):
    B, T, H = shape
    cpu = O.make_inputs(B, T, H, 64, seed=seed, kind=kind)
    w, q, k, v, a, b, dy = [x.cuda() for x in cpu]
    W.set_variant(fv, bv)
    try:
        y, s, sa = W.forward_raw(w, q, k, v, a, b, bounded_decay=True, chunk_checkpoints=ck)
        g = W.backward_raw(w, q, k, v, a, b, dy, s, sa, bounded_decay=True) if do_bwd else None
        torch.cuda.synchronize()
    finally:
        W.set_variant(0, 0)
    W.domain_check()
    y64, s64, sa64 = O.forward(*cpu[:6])
    if ck:
        s64 = s64[:, :, 3::4]
    res = {}
    san, sn = sa.cpu().numpy(), s.cpu().numpy()
    res["sa_err"] = O.err_ratio(san, sa64)
    res["sa_fail"] = int((~np.isclose(san, sa64, rtol=1e-3, atol=1e-5)).sum())
    res["s_err"] = O.err_ratio(sn, s64)
    res["s_fail"] = int((~np.isclose(sn, s64, rtol=1e-3, atol=1e-5)).sum())
    yn = y.float().cpu().numpy()
    res["y_err"] = O.err_ratio(yn, y64)
    res["y_ident"] = float((O.bf16_ulp_diff(yn, O.to_bf16_f32(y64)) == 0).mean())
    res["y_1ulp"] = float((O.bf16_ulp_diff(yn, O.to_bf16_f32(y64)) <= 1).mean())
    if do_bwd:
        g64 = O.backward(*cpu, s64 if not ck else O.forward(*cpu[:6])[1], sa64)
        for n, x, r_ in zip(NAMES, g, g64):
            xn = x.float().cpu().numpy()
            res[n] = O.err_ratio(xn, r_)
            res[n + "_ident"] = float((O.bf16_ulp_diff(xn, O.to_bf16_f32(r_)) == 0).mean())
    return res


def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(n):
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    do_bwd = "--bwd" in sys.argv
    fv = int(os.environ.get("FV", "6"))
    bv = int(os.environ.get("BV", "0"))
    cases = [((1, 64, 1), "realistic", 3), ((2, 128, 2), "realistic", 4), ((2, 512, 3), "realistic", 5), ((1, 2048, 2), "realistic", 6),
             ((2, 256, 2), "stress", 8), ((1, 1024, 2), "stress", 4)]
    for shape, kind, seed in cases:
        for ck in (True, False):
            key = f"{shape}-{kind}-{'ck64' if ck else 'ck16'}"
            try:
                out[key] = check(shape, kind, seed, fv, bv, ck, do_bwd)
            except Exception as e:  # noqa: BLE001
                out[key] = {"error": repr(e)}
            print(key, json.dumps(out[key]), flush=True)
    if "--time" in sys.argv:
        B, T, H = 8, 2048, 12
        w, q, k, v, a, b, dy = [x.cuda() for x in O.make_inputs(B, T, H, 64, seed=42)]
        for name, f_, ck in (("x6-ck64", 6, True), ("x6-ck16", 6, False), ("tf32-ck64", 3, True), ("step", 1, False)):
            W.set_variant(f_, 0)
            try:
                ms = timeit(lambda: W.forward_raw(w, q, k, v, a, b, bounded_decay=True, chunk_checkpoints=ck))
            finally:
                W.set_variant(0, 0)
            out["time_fwd_" + name] = ms
            print("fwd", name, f"{ms:.4f} ms  ({14 * B * T * H * 64 / ms / 1e6:.0f} GB/s alg.)", flush=True)
        if do_bwd:
            for name, f_, b_, ck in (("x6", 6, bv, True), ("tf32", 3, 5, True), ("step", 1, 1, False)):
                W.set_variant(f_, b_)
                try:
                    y, s, sa = W.forward_raw(w, q, k, v, a, b, bounded_decay=True, chunk_checkpoints=ck)
                    ms = timeit(lambda: W.backward_raw(w, q, k, v, a, b, dy, s, sa, bounded_decay=True))
                finally:
                    W.set_variant(0, 0)
                out["time_bwd_" + name] = ms
                print("bwd", name, f"{ms:.4f} ms  ({26 * B * T * H * 64 / ms / 1e6:.0f} GB/s alg.)", flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/dev_x6.json", "w"), indent=1)


if __name__ == "__main__":
    main()
