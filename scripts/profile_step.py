"""Per-kernel device-time breakdown of ONE steady-state training step at the bench config (torch.profiler / CUPTI)."""
import json, os, sys, collections, re
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from visualrwkv_b200 import _lib
from visualrwkv_b200.model import VisualRWKV, default_args, randomize_zero_init
from visualrwkv_b200.synthetic import make_batch
_lib.load_torch_ops()
torch.manual_seed(0)
args = default_args()
model = VisualRWKV(args); randomize_zero_init(model)
model = model.to(device="cuda", dtype=torch.bfloat16); model.freeze_emb()
batch = make_batch(8, 2048, 576, 224, seed=1, device="cuda", img_dtype=torch.bfloat16)
def step():
    for p in model.parameters(): p.grad = None
    loss = model.training_step(batch); loss.backward(); return loss
for _ in range(3): step()
torch.cuda.synchronize()
import time
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f'CPU enqueue {1e3*(t1-t0):.1f} ms, step wall {1e3*(t2-t0):.1f} ms')
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        n = re.sub(r"<.*", "", ev.name); n = re.sub(r"\(.*", "", n)[:70]
        agg[n][0] += 1; agg[n][1] += ev.device_time
tot = sum(v for _, v in agg.values())
rows = sorted(agg.items(), key=lambda x: -x[1][1])
print(f"total device time {tot/1000:.2f} ms over {sum(n for n,_ in agg.values())} kernels")
out = []
for k, (n, v) in rows[:45]:
    print(f"{100*v/tot:5.1f}%  {v/1000:7.2f} ms  n={n:4d}  {k}")
    out.append({"kernel": k, "n": n, "ms": v / 1000})
cpu = sorted([(e.self_cpu_time_total, e.key, e.count) for e in prof.key_averages()], reverse=True)[:25]
print('--- top CPU self time (us)')
for t, k, n in cpu:
    print(f'{t:9.0f}  n={n:4d}  {k[:80]}')
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "profile_step.json"), "w"), indent=1)
