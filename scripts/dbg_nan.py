import sys, torch
sys.path.insert(0, ".")
from oracle import model_ref as MR
from visualrwkv_b200 import fused
from visualrwkv_b200.model import VisualRWKV, default_args, randomize_zero_init

torch.manual_seed(0)
args = default_args(n_embd=128, n_layer=2, dim_att=128, vision_tower_path="siglip-tiny-test", num_token_per_image=16, ctx_len=128)
model = VisualRWKV(args)
randomize_zero_init(model)
model = model.to(device="cuda", dtype=torch.bfloat16)
model.freeze_emb()
b1 = MR.make_batch(2, 128, 16, 64, seed=1, device="cuda", img_dtype=torch.bfloat16)
# poison the caching allocator: freed blocks keep NaN bit patterns
junk = [torch.full((1 << 22,), float("nan"), device="cuda") for _ in range(16)]
del junk
orig_lmb = fused.ln_mix_backward


def spy(x2d, T, stats, gamma, beta, coefs, douts, dh=None, dresid=None):
    out = orig_lmb(x2d, T, stats, gamma, beta, coefs, douts, dh, dresid)
    torch.cuda.synchronize()
    nin = [int(torch.isnan(d.float()).sum()) for d in douts] + [int(torch.isnan(t.float()).sum()) if t is not None else -1 for t in (dh, dresid)]
    nout = [int(torch.isnan(o.float()).sum()) for o in (out[0], out[1], out[2])] + [int(torch.isnan(o.float()).sum()) for o in out[3]]
    if sum(nout) or sum(n for n in nin if n > 0):
        print("  ln_mix_backward ncoef", len(coefs), "rows", x2d.shape, "NaN in", nin, "NaN out", nout,
              "cols", torch.nonzero(torch.isnan(out[1].float())).flatten()[:12].tolist())
    return out


fused.ln_mix_backward = spy
for it in range(3):
    for p in model.parameters():
        p.grad = None
    loss = model.training_step(b1)
    loss.backward()
    torch.cuda.synchronize()
    bad = [(n, int(torch.isnan(p.grad.float()).sum())) for n, p in model.named_parameters() if p.grad is not None and torch.isnan(p.grad.float()).any()]
    print(it, float(loss), bad)
