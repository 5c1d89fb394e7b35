import sys, torch
sys.path.insert(0, ".")
from oracle import model_ref as MR
from visualrwkv_b200.graph import GraphedTrainStep
from visualrwkv_b200.model import VisualRWKV, default_args, randomize_zero_init

torch.manual_seed(0)
args = default_args(n_embd=128, n_layer=2, dim_att=128, vision_tower_path="siglip-tiny-test", num_token_per_image=16, ctx_len=128)
model = VisualRWKV(args)
randomize_zero_init(model)
model = model.to(device="cuda", dtype=torch.bfloat16)
model.freeze_emb()
b1 = MR.make_batch(2, 128, 16, 64, seed=1, device="cuda", img_dtype=torch.bfloat16)
b2 = MR.make_batch(2, 128, 16, 64, seed=2, device="cuda", img_dtype=torch.bfloat16)
rel = lambda a, b: ((a.float() - b.float()).pow(2).mean().sqrt() / (b.float().pow(2).mean().sqrt() + 1e-30)).item()


def eager(batch):
    for p in model.parameters():
        p.grad = None
    loss = model.training_step(batch)
    loss.backward()
    return float(loss), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}


l1, g1 = eager(b1)
l1b, g1b = eager(b1)
l2, g2 = eager(b2)
print("eager repeat:", l1, l1b, max(rel(g1b[n], g1[n]) for n in g1))
print("valid targets", [(b["targets"] != -100).sum().item() if "targets" in b else None for b in (b1, b2)], list(b1.keys()))
step = GraphedTrainStep(model, b1)
for it, (batch, lref, gref) in enumerate(((b1, l1, g1), (b2, l2, g2), (b1, l1, g1))):
    loss = step(batch)
    torch.cuda.synchronize()
    bad = [(n, round(rel(p.grad, gref[n]), 5)) for n, p in model.named_parameters() if n in gref and rel(p.grad, gref[n]) > 1e-3]
    print(it, float(loss), lref, bad[:8], len(bad))
    if it == 1:
        hw = dict(model.named_parameters())["rwkv.head.weight"].grad
        d = (hw.float() - gref["rwkv.head.weight"].float()).abs()
        rows = d.amax(1)
        print("rows differing:", (rows > 0).sum().item(), "of", rows.numel(), "first", torch.nonzero(rows > 0)[:10].flatten().tolist())
        print("vs g1:", rel(hw, g1["rwkv.head.weight"]))
