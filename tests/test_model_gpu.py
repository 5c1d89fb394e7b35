"""GPU parity of the whole path (ViT -> pool -> projector -> scatter -> RWKV -> head -> loss, forward + backward)
against the fp64 CPU oracle, and of the fused head+CE+L2Wrap against the reference-shaped unfused graph."""
import numpy as np
import pytest
import torch

from oracle import model_ref as MR

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach().double().cpu().numpy(), b.detach().double().cpu().numpy()
    return float(np.sqrt(np.mean((a - b) ** 2)) / max(np.sqrt(np.mean(b ** 2)), 1e-30))


def _tiny(seed=0, T=64, n_img_tok=16):
    from visualrwkv_b200.model import VisualRWKV, default_args, randomize_zero_init
    torch.manual_seed(seed)
    args = default_args(n_embd=128, n_layer=2, dim_att=128, vision_tower_path="siglip-tiny-test",
                        num_token_per_image=n_img_tok, ctx_len=T)
    m = VisualRWKV(args)
    randomize_zero_init(m)
    return m.to(device="cuda", dtype=torch.bfloat16), args


def _oracle(m, batch, args, n_img_tok):
    P = {k: v.detach().double().cpu() for k, v in m.state_dict().items()}
    for k in P:
        if not k.startswith("vit.") and "emb.weight" not in k:
            P[k].requires_grad_(True)
    cfg = {"vit": m.vit.cfg, "num_token_per_image": n_img_tok, "n_layer": args.n_layer, "n_head": args.dim_att // 64}
    cb = {"input_ids": batch["input_ids"].cpu(), "labels": batch["labels"].cpu(), "images": batch["images"].double().cpu()}
    logits, targets = MR.visual_forward(P, cb, cfg, MR.oracle_wkv("f64"))
    return P, logits, targets


@pytest.mark.parametrize("T", [64, 56])  # 56: exercises the left-pad to a multiple of 16 (model.py:309-312)
def test_visual_rwkv_forward_backward_vs_oracle(T):
    n_img = 16
    m, args = _tiny(T=T)
    batch = MR.make_batch(2, T, n_img, 64, seed=3, device="cuda", img_dtype=torch.bfloat16, human_tokens=4)
    logits, targets = m(batch)
    loss = m.training_step(batch)
    loss.backward()
    P, rlogits, rtargets = _oracle(m, batch, args, n_img)
    rloss = MR.training_loss(rlogits, rtargets)
    rloss.backward()
    assert logits.shape == rlogits.shape == (2, T, args.vocab_size)
    assert _rel(logits, rlogits) < 3e-2
    assert abs(float(loss) - float(rloss)) < 2e-2 * abs(float(rloss))
    checked = 0
    for n, p in m.named_parameters():
        if p.grad is None or P[n].grad is None:
            continue
        if p.dim() == 2 and min(p.shape) >= 32:  # weight matrices: clean signal
            assert _rel(p.grad, P[n].grad) < 0.12, n
            checked += 1
    assert checked >= 20
    assert m.proj.o_proj.weight.grad is not None and m.rwkv.emb.weight.grad is not None


def test_fused_head_loss_matches_unfused_graph():
    from visualrwkv_b200 import ops
    from visualrwkv_b200.model import IGNORE_INDEX, L2Wrap
    m, args = _tiny(seed=1)
    batch = MR.make_batch(2, 64, 16, 64, seed=5, device="cuda", img_dtype=torch.bfloat16, human_tokens=4)
    loss_f = m.training_step(batch)
    loss_f.backward()
    g_f = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    m.zero_grad()
    logits, targets = m(batch)
    loss_u = ops.training_loss(logits, targets, IGNORE_INDEX, L2Wrap)
    loss_u.backward()
    assert abs(float(loss_f) - float(loss_u)) < 1e-2 * abs(float(loss_u))
    for n, p in m.named_parameters():
        if p.grad is not None and p.dim() == 2 and min(p.shape) >= 32:
            assert _rel(g_f[n], p.grad) < 6e-2, n


def test_scatter_is_bit_exact_and_ordered():
    from visualrwkv_b200 import ops
    torch.manual_seed(0)
    emb = torch.randn(70000, 128, device="cuda").to(torch.bfloat16)
    ids = torch.randint(0, 65535, (3, 40), device="cuda")
    ids[0, 3:11] = 65535
    ids[2, 0:8] = 65535
    feats = torch.randn(2, 8, 128, device="cuda").to(torch.bfloat16)
    x = ops.embed_and_scatter(emb, ids, feats, 65535)
    ref = emb[ids].clone()
    ref[0, 3:11] = feats[0]
    ref[2, 0:8] = feats[1]
    assert torch.equal(x, ref)
    # fewer slots than features: the reference truncates the features (model.py:487-491)
    ids2 = ids.clone()
    ids2[2, 4:8] = 7
    x2 = ops.embed_and_scatter(emb, ids2, feats, 65535)
    assert torch.equal(x2[0, 3:11], feats[0]) and torch.equal(x2[2, 0:4], feats[1][:4])
    assert torch.equal(x2[2, 4:8], emb[ids2[2, 4:8]])
    with pytest.warns(UserWarning):  # the count check is deferred (no host sync inside the forward)
        ops.flush_checks()
    # more slots than features is an error in the reference (index_put shape mismatch); here it is raised late too
    ids3 = ids.clone()
    ids3[1, 0:4] = 65535
    ops.embed_and_scatter(emb, ids3, feats, 65535)
    with pytest.warns(UserWarning), pytest.raises(RuntimeError):
        ops.flush_checks()
    # gradient reaches exactly the gathered feature rows
    f = feats.clone().float().requires_grad_(True)
    ops.embed_and_scatter(emb.float(), ids, f, 65535).sum().backward()
    ops.flush_checks()
    assert torch.equal(f.grad, torch.full_like(f, 1.0))


def test_vit_matches_hf_siglip_on_gpu():
    transformers = pytest.importorskip("transformers")
    from visualrwkv_b200.vision import SiglipVisionTower
    cfg = transformers.SiglipVisionConfig(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                                          image_size=64, patch_size=16)
    torch.manual_seed(0)
    hf = transformers.SiglipVisionModel(cfg).eval()
    tower = SiglipVisionTower("siglip-tiny-test")
    missing, unexpected = tower.load_state_dict(hf.state_dict(), strict=False)
    assert not missing, missing  # every tower parameter exists in the HF checkpoint layout
    px = torch.randn(3, 3, 64, 64)
    with torch.no_grad():
        ref = hf(pixel_values=px).last_hidden_state
    out = tower.to("cuda", torch.bfloat16)(px.to("cuda", torch.bfloat16))
    assert out.shape == ref.shape
    assert _rel(out, ref) < 3e-2


def test_bench_line_keys():
    """The GPU arm of bench.py on a reduced configuration (2 layers, ctx 1024, batch 2): one JSON line carrying the
    contract's keys, a non-zero launch count of this library's kernels and the roofline object of the WKV7 pair."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "2", "--warmup", "3", "--no-cpu-baseline",
                          "--layers", "2", "--ctx", "1024", "--batch", "2"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert k in d, k
    assert d["gpu_launches"] > 0 and d["value"] > 0 and d["e2e"]["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] == 4
    r = d["roofline"]
    assert r["bound"] == "hbm" and 0 < r["frac"] < 1 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert "x3_bwd" in r["kernel"] and d["config"]["wkv"] == "x6"  # default: the chunked kernels that pass the strict parity tests
