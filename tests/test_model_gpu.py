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


@pytest.mark.parametrize("name,hf_kw,N", [
    ("siglip-tiny-test", dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, image_size=64, patch_size=16), 3),
    # the real tower of BASELINE cfg2: SigLIP-B/16 @ 224 -> 196 patches (random weights: no checkpoints offline)
    ("siglip-base-patch16-224", dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072, image_size=224, patch_size=16), 2),
    ("siglip-base-patch16-256", dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072, image_size=256, patch_size=16), 1),
])
def test_vit_matches_hf_siglip_on_gpu(name, hf_kw, N):
    """Our tower (im2col + CTA-pair GEMMs with bias/GELU/residual epilogues + the tcgen05 attention kernel + LayerNorm
    kernels) against transformers' SiglipVisionModel in fp32 on the CPU, same weights."""
    transformers = pytest.importorskip("transformers")
    from visualrwkv_b200 import fused
    from visualrwkv_b200.vision import SiglipVisionTower
    cfg = transformers.SiglipVisionConfig(**hf_kw)
    torch.manual_seed(0)
    hf = transformers.SiglipVisionModel(cfg).eval()
    tower = SiglipVisionTower(name)
    missing, unexpected = tower.load_state_dict(hf.state_dict(), strict=False)
    assert not missing, missing  # every tower parameter exists in the HF checkpoint layout
    px = torch.randn(N, 3, hf_kw["image_size"], hf_kw["image_size"])
    with torch.no_grad():
        ref = hf(pixel_values=px.to(torch.bfloat16).float()).last_hidden_state
    from visualrwkv_b200 import wkv7
    n0 = wkv7.launch_count()
    out = tower.to("cuda", torch.bfloat16)(px.to("cuda", torch.bfloat16))
    assert wkv7.launch_count() - n0 >= 2 + 7 * hf_kw["num_hidden_layers"]   # the tower ran on this library's kernels, not on a fallback
    assert out.shape == ref.shape
    assert _rel(out, ref) < 3e-2


def test_vit_attention_kernel_vs_sdpa():
    from visualrwkv_b200 import _lib, fused
    torch.manual_seed(1)
    for (N, S, H) in [(2, 196, 12), (1, 256, 4), (3, 16, 2), (1, 100, 1)]:
        q, k, v = [(torch.randn(N * S, 64 * H, device="cuda")).to(torch.bfloat16) for _ in range(3)]
        o = torch.empty_like(q)
        fused._chk(_lib.lib().vrwkv_vit_attention(N, S, H, fused._p(q), fused._p(k), fused._p(v), fused._p(o), _lib.cur_stream()), "attn")
        f = lambda t: t.view(N, S, H, 64).transpose(1, 2).float()
        ref = torch.nn.functional.scaled_dot_product_attention(f(q), f(k), f(v)).transpose(1, 2).reshape(N * S, 64 * H)
        assert _rel(o, ref) < 6e-3, (N, S, H)


def test_adaptive_pool_and_projector_kernels():
    from visualrwkv_b200 import ops
    from visualrwkv_b200.model import MLPWithContextGating
    torch.manual_seed(2)
    x = torch.randn(3, 196, 768, device="cuda").to(torch.bfloat16)
    for out_hw in (24, 7, 14):
        y = ops.adaptive_pooling(x, out_hw)
        ref = torch.nn.functional.adaptive_avg_pool2d(x.float().view(3, 14, 14, 768).permute(0, 3, 1, 2), out_hw).reshape(3, 768, -1).permute(0, 2, 1)
        assert y.shape == ref.shape and _rel(y, ref) < 3e-3
    m = MLPWithContextGating(768, 768).to("cuda", torch.bfloat16)
    xin = torch.randn(2, 576, 768, device="cuda").to(torch.bfloat16)
    g = torch.randn(2, 576, 768, device="cuda").to(torch.bfloat16)
    xa = xin.clone().requires_grad_(True)
    ya = ops.projector_forward(m, xa)
    ya.backward(g)
    mine = {n: p.grad.clone() for n, p in m.named_parameters()}
    m.zero_grad()
    xb = xin.clone().float().requires_grad_(True)
    P = {n: p.detach().float().requires_grad_(True) for n, p in m.named_parameters()}
    hb = xb * torch.sigmoid(xb @ P["gate.weight"].t())
    yb = torch.nn.functional.layer_norm(hb @ P["o_proj.weight"].t(), (768,), P["ln_v.weight"], P["ln_v.bias"], m.ln_v.eps)
    yb.backward(g.float())
    assert _rel(ya, yb) < 1e-2 and _rel(xa.grad, xb.grad) < 3e-2
    for n in mine:
        assert _rel(mine[n], P[n].grad) < 3e-2, n


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


def test_graphed_train_step_matches_eager():
    """visualrwkv_b200.graph.GraphedTrainStep: the whole step replayed as one CUDA graph (native kernels, their
    stream-ordered workspaces and the flag-chained WKV7 kernels included) gives the eager step's loss and gradients."""
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

    def eager(batch):
        for p in model.parameters():
            p.grad = None
        loss = model.training_step(batch)
        loss.backward()
        return float(loss), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}

    l1, g1 = eager(b1)
    l2, g2 = eager(b2)
    step = GraphedTrainStep(model, b1)
    for batch, lref, gref in ((b1, l1, g1), (b2, l2, g2), (b1, l1, g1)):
        loss = step(batch)
        torch.cuda.synchronize()
        assert abs(float(loss) - lref) < 1e-3 * max(1.0, abs(lref))
        for n, p in model.named_parameters():
            if n in gref:
                assert _rel(p.grad, gref[n]) < 1e-3 or float(gref[n].float().abs().max()) == 0.0, n


def test_fused_adamw_matches_torch_adamw():
    """csrc/optim.cu against torch.optim.AdamW on fp32 master copies: odd sizes (scalar tails, unaligned slices), weight
    decay, three steps, and the same three steps replayed from a CUDA graph."""
    from visualrwkv_b200.optim import FusedAdamW
    torch.manual_seed(0)
    shapes = [(768, 768), (1, 1, 768), (12, 64), (1000, 7), (5,), (65536, 16), (3, 33)]
    params = [torch.nn.Parameter((0.1 * torch.randn(*s, device="cuda")).to(torch.bfloat16)) for s in shapes]
    ref = [p.detach().float().clone().requires_grad_(True) for p in params]
    kw = dict(lr=3e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.01)
    opt = FusedAdamW(params, **kw)
    topt = torch.optim.AdamW(ref, **kw)
    grads = [[(0.05 * torch.randn(*s, device="cuda")).to(torch.bfloat16) for s in shapes] for _ in range(3)]
    for gs in grads:
        for p, r, g in zip(params, ref, gs):
            p.grad = g.clone()
            r.grad = g.float()
        opt.step()
        topt.step()
    torch.cuda.synchronize()
    for i, (p, r) in enumerate(zip(params, ref)):
        assert torch.allclose(opt.master_of(i), r.detach(), rtol=2e-5, atol=1e-7), shapes[i]
        assert torch.equal(p.detach(), opt.master_of(i).to(torch.bfloat16)), shapes[i]
    assert int(opt.step_count) == 3
    # graph replay: static gradient buffers, the step counter advances on the device
    for p in params:
        p.grad = torch.zeros_like(p)
    opt.step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        opt.step()
    before = int(opt.step_count)
    g.replay()
    torch.cuda.synchronize()
    assert int(opt.step_count) == before + 1


def test_recurrent_forward_split_sequence_equivalence():
    """prefill + stateful continuation equals one pass over the whole sequence (no reference oracle for
    this row): 128 tokens at once (chunked tensor-core WKV7) vs 64 + 64 vs 127 + 1 (step-by-step kernel, T = 1 decode with
    M = 2 GEMMs), and vs the stateless training-path forward."""
    from visualrwkv_b200.recurrent import rwkv_forward_recurrent
    m, args = _tiny(seed=2, T=128)
    torch.manual_seed(5)
    x = (0.5 * torch.randn(2, 128, args.n_embd, device="cuda")).to(torch.bfloat16)
    with torch.no_grad():
        ref = m.rwkv(x)                                              # stateless path, [2,128,V]
    full, st_full = rwkv_forward_recurrent(m.rwkv, x, last_only=False)
    assert _rel(full, ref) < 2e-2
    for cut in (64, 127, 16):
        a, st = rwkv_forward_recurrent(m.rwkv, x[:, :cut].contiguous(), last_only=False)
        b, st = rwkv_forward_recurrent(m.rwkv, x[:, cut:].contiguous(), st, last_only=False)
        assert st.tokens_seen == 128
        assert _rel(torch.cat([a, b], dim=1), full) < 2e-2, cut
        for l0, l1 in zip(st.layers, st_full.layers):
            assert _rel(l0.wkv, l1.wkv) < 2e-2, cut
            assert _rel(l0.att_prev, l1.att_prev) < 2e-2 and _rel(l0.ffn_prev, l1.ffn_prev) < 2e-2
    last, _ = rwkv_forward_recurrent(m.rwkv, x)
    assert last.shape == (2, args.vocab_size) and _rel(last, full[:, -1]) < 1e-3


def test_recurrent_generate_matches_stateless_loop():
    """VisualRWKV.generate(recurrent=True): first token from the same prompt logits as the reference-style loop (prompt
    length a multiple of 16, so the loop's left padding is empty), then five single-token steps (B = 1, T = 1)."""
    m, args = _tiny(seed=3, T=64)
    batch = MR.make_batch(1, 64, 16, 64, seed=7, device="cuda", img_dtype=torch.bfloat16, human_tokens=4)
    ids, imgs = batch["input_ids"], batch["images"]
    with torch.no_grad():
        t_loop, l_loop, p_loop = m.generate(ids, imgs, False, 1.0, 1.0, 1, -1)
    t_rec, l_rec, p_rec = m.generate(ids, imgs, False, 1.0, 1.0, 6, -1, recurrent=True)
    assert len(t_rec) == 6 and all(0 <= t < args.vocab_size for t in t_rec)
    assert abs(l_rec[0] - l_loop[0]) < 2e-2 * max(1.0, abs(l_loop[0]))
    assert t_rec[0] == t_loop[0] or abs(l_rec[0] - l_loop[0]) < 1e-2


def test_graphed_decoder_matches_eager_decode():
    """GraphedDecoder (one CUDA graph per T = 1 step, state updated in place) against eager recurrent steps."""
    from visualrwkv_b200.recurrent import GraphedDecoder, rwkv_forward_recurrent
    m, args = _tiny(seed=4, T=64)
    torch.manual_seed(6)
    x = (0.5 * torch.randn(3, 48, args.n_embd, device="cuda")).to(torch.bfloat16)
    toks = torch.randint(0, 60000, (4, 3, 1), device="cuda")
    _, st_e = rwkv_forward_recurrent(m.rwkv, x)
    _, st_g = rwkv_forward_recurrent(m.rwkv, x)
    dec = GraphedDecoder(m, st_g, 3)
    for t in toks:
        le, st_e = rwkv_forward_recurrent(m.rwkv, m.rwkv.emb(t).to(torch.bfloat16), st_e)
        lg = dec.step(t)
        assert _rel(lg, le) < 1e-3
    assert st_g.tokens_seen == st_e.tokens_seen == 52
    for a, b in zip(st_g.layers, st_e.layers):
        assert _rel(a.wkv, b.wkv) < 1e-3 and torch.equal(a.att_prev, b.att_prev)


@pytest.mark.parametrize("V", [64, 512, 2048])
def test_head_loss_small_vocab_is_finite_and_matches(V):
    """lanes without elements (V <= 2032) used to merge -inf maxima into NaN in the CE reduction."""
    from visualrwkv_b200 import fused
    torch.manual_seed(V)
    B, T, C = 2, 32, 128
    x = (0.5 * torch.randn(B, T, C, device="cuda")).to(torch.bfloat16).requires_grad_(True)
    w = (0.05 * torch.randn(V, C, device="cuda")).to(torch.bfloat16).requires_grad_(True)
    labels = torch.randint(0, V, (B, T), device="cuda")
    labels[0, :5] = -100
    loss = fused.HeadLossFn.apply(x, w, labels, -100)
    loss.backward()
    assert torch.isfinite(loss) and torch.isfinite(x.grad.float()).all() and torch.isfinite(w.grad.float()).all()
    logits = (x.detach().float().reshape(-1, C) @ w.detach().float().t()).view(B, T, V)
    ce = torch.nn.functional.cross_entropy(logits[:, :-1].reshape(-1, V), labels[:, 1:].reshape(-1), ignore_index=-100, reduction="none").view(B, T - 1)
    valid = (labels[:, 1:] != -100).sum(1).clamp(min=1)
    ref = (ce.sum(1) / valid).mean()
    assert abs(float(loss) - float(ref)) < 2e-2 * float(ref)
    l2 = fused.HeadLossFn.apply(x, w, labels, -100)
    l2.backward(retain_graph=True)
    with pytest.raises(RuntimeError):    # the logits buffer now holds the gradient: a second backward must not silently reuse it
        l2.backward()
