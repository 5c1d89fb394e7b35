"""Recurrent inference timing (BASELINE config 5: batch 32, streaming state): prefill tokens/s and decode
tokens/s of the RWKV-x070 tower, eager launches vs one CUDA graph per decode step.  Random-init weights, synthetic ids."""
import argparse, json, sys, time
import torch
sys.path.insert(0, ".")
from visualrwkv_b200.model import RWKV, default_args, randomize_zero_init
from visualrwkv_b200.recurrent import GraphedDecoder, rwkv_forward_recurrent

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="0.1B", choices=["0.1B", "1.5B"])
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--prefill", type=int, default=512)
ap.add_argument("--steps", type=int, default=64)
a = ap.parse_args()
L, C = {"0.1B": (12, 768), "1.5B": (24, 2048)}[a.model]
torch.manual_seed(0)
args = default_args(n_layer=L, n_embd=C, dim_att=C, ctx_len=4096)
rwkv = RWKV(args)
randomize_zero_init(rwkv)
rwkv = rwkv.to(device="cuda", dtype=torch.bfloat16).eval()


class _M:   # GraphedDecoder only needs .rwkv
    pass


m = _M(); m.rwkv = rwkv
B = a.batch
ids = torch.randint(0, 65000, (B, a.prefill), device="cuda")
x = rwkv.emb(ids).to(torch.bfloat16)
ev = lambda: torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    for _ in range(2):
        logits, state = rwkv_forward_recurrent(rwkv, x)
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    logits, state = rwkv_forward_recurrent(rwkv, x)
    e1.record()
    torch.cuda.synchronize()
    prefill_ms = e0.elapsed_time(e1)
    # eager decode
    tok = torch.argmax(logits, -1, keepdim=True)
    for _ in range(3):
        logits, state = rwkv_forward_recurrent(rwkv, rwkv.emb(tok).to(torch.bfloat16), state)
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(a.steps):
        logits, state = rwkv_forward_recurrent(rwkv, rwkv.emb(tok).to(torch.bfloat16), state)
        tok = torch.argmax(logits, -1, keepdim=True)
    e1.record()
    torch.cuda.synchronize()
    eager_ms = e0.elapsed_time(e1) / a.steps
    dec = GraphedDecoder(m, state, B)
    for _ in range(3):
        logits = dec.step(tok)
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(a.steps):
        logits = dec.step(tok)
        tok = torch.argmax(logits, -1, keepdim=True)
    e1.record()
    torch.cuda.synchronize()
    graph_ms = e0.elapsed_time(e1) / a.steps
wbytes = sum(p.numel() * 2 for n, p in rwkv.named_parameters() if "emb" not in n)
sbytes = L * B * (C // 64) * 64 * 64 * 4 * 2
print(json.dumps({"workload": f"RWKV-x070 {a.model} recurrent inference, batch {B}, prefill {a.prefill} then {a.steps} decode steps",
                  "prefill_ms": prefill_ms, "prefill_tokens_per_s": B * a.prefill / prefill_ms * 1e3,
                  "decode_ms_per_step_eager": eager_ms, "decode_ms_per_step_graph": graph_ms,
                  "decode_tokens_per_s_graph": B / graph_ms * 1e3,
                  "weight_bytes_per_step": wbytes, "state_bytes_per_step": sbytes,
                  "decode_hbm_gbps_graph": (wbytes + sbytes) / graph_ms / 1e6}))
