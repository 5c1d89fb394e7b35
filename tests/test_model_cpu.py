"""CPU tests of the host-side mirror: state-dict keys / shapes / init as the reference defines them, the model-level
oracle against transformers' SigLIP, and the data-parallel gradient reducer over gloo (world_size 2)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

from oracle import model_ref as MR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_state_dict_keys_and_shapes_follow_reference():
    from visualrwkv_b200.model import VisualRWKV, default_args
    args = default_args(n_embd=256, n_layer=3, dim_att=256, vision_tower_path="siglip-tiny-test", num_token_per_image=16)
    m = VisualRWKV(args)
    sd = m.state_dict()
    C, H = 256, 4
    # reference names (v7.00/src/model.py:96-154,212-215,240-242,278-281,331-333)
    exp = {"rwkv.emb.weight": (65536, C), "rwkv.head.weight": (65536, C), "rwkv.ln_out.weight": (C,),
           "rwkv.blocks.0.ln0.weight": (C,), "rwkv.blocks.1.ln1.bias": (C,), "rwkv.blocks.2.ln2.weight": (C,),
           "rwkv.blocks.1.att.x_r": (1, 1, C), "rwkv.blocks.1.att.x_g": (1, 1, C), "rwkv.blocks.1.att.w0": (1, 1, C),
           "rwkv.blocks.1.att.w1": (C, 32), "rwkv.blocks.1.att.w2": (32, C), "rwkv.blocks.1.att.a1": (C, 32),
           "rwkv.blocks.1.att.v0": (1, 1, C), "rwkv.blocks.1.att.v1": (C, 32), "rwkv.blocks.1.att.v2": (32, C),
           "rwkv.blocks.1.att.g1": (C, 64), "rwkv.blocks.1.att.g2": (64, C), "rwkv.blocks.1.att.k_k": (1, 1, C),
           "rwkv.blocks.1.att.k_a": (1, 1, C), "rwkv.blocks.1.att.r_k": (H, 64), "rwkv.blocks.1.att.receptance.weight": (C, C),
           "rwkv.blocks.1.att.key.weight": (C, C), "rwkv.blocks.1.att.value.weight": (C, C),
           "rwkv.blocks.1.att.output.weight": (C, C), "rwkv.blocks.1.att.ln_x.weight": (C,), "rwkv.blocks.1.att.ln_x.bias": (C,),
           "rwkv.blocks.1.ffn.x_k": (1, 1, C), "rwkv.blocks.1.ffn.key.weight": (4 * C, C), "rwkv.blocks.1.ffn.value.weight": (C, 4 * C),
           "proj.gate.weight": (128, 128), "proj.o_proj.weight": (C, 128), "proj.ln_v.weight": (C,)}
    for k, shp in exp.items():
        assert k in sd, k
        assert tuple(sd[k].shape) == shp, (k, tuple(sd[k].shape), shp)
    assert "rwkv.blocks.0.att.v1" not in sd and "rwkv.blocks.1.ln0.weight" not in sd  # model.py:134,240
    assert any(k.startswith("vit.vision_model.encoder.layers.1.self_attn.q_proj") for k in sd)


def test_lora_ranks_and_init_values():
    from visualrwkv_b200.model import RWKV, default_args
    for C, ranks in [(768, (64, 64, 32, 128)), (2048, (96, 96, 64, 256))]:
        args = default_args(n_embd=C, n_layer=2, dim_att=C, vocab_size=16)
        att = RWKV(args).blocks[1].att
        assert (att.w1.shape[1], att.a1.shape[1], att.v1.shape[1], att.g1.shape[1]) == ranks
    args = default_args(n_embd=128, n_layer=4, dim_att=128, vocab_size=16)
    m = RWKV(args)
    att = m.blocks[2].att
    C, lid, L = 128, 2, 4
    ddd = torch.arange(C, dtype=torch.float32) / C
    r10, r01 = 1 - lid / L, lid / (L - 1)
    assert torch.allclose(att.x_r.flatten(), 1 - ddd ** (0.2 * r10))
    assert torch.allclose(att.x_k.flatten(), 1 - (ddd ** (0.9 * r10) + 0.4 * r01))
    n = torch.arange(C, dtype=torch.float32)
    assert torch.allclose(att.w0.flatten(), -7 + 5 * (n / (C - 1)) ** (0.85 + r01 ** 0.5) + 0.5)
    assert float(att.output.weight.abs().max()) == 0.0 and float(m.blocks[2].ffn.value.weight.abs().max()) == 0.0
    assert abs(float(att.k_k.mean()) - 0.85) < 1e-6 and att.ln_x.eps == pytest.approx(64e-5)
    assert float(att.receptance.weight.abs().max()) <= 0.5 / C ** 0.5 + 1e-6
    assert torch.allclose(m.blocks[2].ffn.x_k.flatten(), 1 - ddd ** (r10 ** 4))


def test_oracle_siglip_matches_transformers():
    transformers = pytest.importorskip("transformers")
    cfg = transformers.SiglipVisionConfig(hidden_size=64, num_hidden_layers=2, num_attention_heads=2, intermediate_size=128,
                                          image_size=32, patch_size=16)
    torch.manual_seed(0)
    hf = transformers.SiglipVisionModel(cfg).eval().double()
    P = {"vit." + k: v for k, v in hf.state_dict().items()}
    px = torch.randn(2, 3, 32, 32, dtype=torch.double)
    with torch.no_grad():
        ref = hf(pixel_values=px).last_hidden_state
        out = MR.siglip_forward(P, px, dict(hidden=64, layers=2, heads=2, patch=16, eps=cfg.layer_norm_eps))
    assert torch.allclose(out, ref, rtol=1e-9, atol=1e-10)


def test_oracle_scatter_pool_loss_semantics():
    torch.manual_seed(0)
    P = {"rwkv.emb.weight": torch.randn(70000, 8, dtype=torch.double)}
    ids = torch.randint(0, 65535, (2, 12))
    ids[1, 2:6] = MR.IMAGE_TOKEN_INDEX
    feats = torch.randn(1, 4, 8, dtype=torch.double)
    x = MR.scatter_image_features(P, ids, feats)
    assert torch.equal(x[1, 2:6], feats[0]) and torch.equal(x[0], P["rwkv.emb.weight"][ids[0]])
    f = torch.arange(2 * 16 * 3, dtype=torch.double).view(2, 16, 3)
    assert torch.equal(MR.adaptive_pooling(f, 16), f)                # identity when the grid already matches
    assert MR.adaptive_pooling(f, 4).shape == (2, 4, 3)
    # loss: shifted, per-sample mean over valid labels, then batch mean; L2Wrap adds 1e-4/(B*T) * max at the argmax
    logits = torch.randn(2, 5, 11, dtype=torch.double, requires_grad=True)
    tg = torch.tensor([[-100, 3, 4, -100, 2], [1, 2, 3, 4, 5]])
    loss = MR.training_loss(logits, tg)
    lp = torch.log_softmax(logits, -1)
    s0 = -(lp[0, 0, 3] + lp[0, 1, 4] + lp[0, 3, 2]) / 3
    s1 = -(lp[1, 0, 2] + lp[1, 1, 3] + lp[1, 2, 4] + lp[1, 3, 5]) / 4
    assert torch.allclose(loss, (s0 + s1) / 2)
    loss.backward()
    g_plain = torch.autograd.grad((s0 + s1) / 2, logits)[0]
    extra = logits.grad - g_plain
    mx, ix = logits.detach().max(-1)
    exp = torch.zeros_like(extra).scatter_(-1, ix.unsqueeze(-1), (mx * 1e-4 / 10).unsqueeze(-1))
    assert torch.allclose(extra, exp, atol=1e-12)


def test_synthetic_batch_contract():
    from visualrwkv_b200.synthetic import make_batch
    b = make_batch(2, 2048, 576, 224, seed=1)
    assert b["input_ids"].shape == (2, 2048) and b["input_ids"].dtype == torch.int64
    assert int((b["input_ids"] == 65535).sum()) == 2 * 576
    assert (b["labels"][b["input_ids"] == 65535] == -100).all() and b["images"].shape == (2, 3, 224, 224)


def test_ddp_bucket_reducer_gloo_world2(tmp_path):
    script = tmp_path / "ddp_worker.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, torch, torch.distributed as dist
        sys.path.insert(0, {ROOT!r})
        from visualrwkv_b200.ddp import GradBucketReducer
        dist.init_process_group("gloo")
        rank = dist.get_rank()
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 8), torch.nn.Linear(8, 4))
        net[3].weight.requires_grad_(False)
        ref = [p.detach().clone() for p in net.parameters()]
        red = GradBucketReducer(net.parameters(), bucket_bytes=1024)
        assert len(red.buckets) >= 3
        for step in range(2):
            red.reset()
            x = torch.randn(5, 16, generator=torch.Generator().manual_seed(100 * step + rank))
            net(x).pow(2).sum().backward()
            red.finish()
        # reference: average of the two ranks' local gradients of the last step
        grads = []
        for r in range(2):
            n2 = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 8), torch.nn.Linear(8, 4))
            for p, q in zip(n2.parameters(), ref):
                p.data.copy_(q)
            x = torch.randn(5, 16, generator=torch.Generator().manual_seed(100 + r))
            n2(x).pow(2).sum().backward()
            grads.append([p.grad for p in n2.parameters()])
        for p, g0, g1 in zip(net.parameters(), *grads):
            if p.requires_grad:
                assert torch.allclose(p.grad, (g0 + g1) / 2, atol=1e-6), "bucketed all-reduce mismatch"
        # hook-free mode (what the CUDA-graphed step uses: collectives outside the graph) + the set_to_none guard
        red.overlap = False
        for p in net.parameters():
            p.grad = None                      # what optimizer.zero_grad(set_to_none=True) does: reset() must re-attach the views
        red.reset()
        red.zero_buckets()
        x = torch.randn(5, 16, generator=torch.Generator().manual_seed(100 + rank))
        net(x).pow(2).sum().backward()
        red.allreduce_all()
        for p, g0, g1 in zip(net.parameters(), *grads):
            if p.requires_grad:
                assert torch.allclose(p.grad, (g0 + g1) / 2, atol=1e-6), "allreduce_all mismatch"
        # a rank-dependent set of parameters with gradients must not reorder the collectives: rank 1 skips the last layer's use
        red.overlap = True
        red.reset()
        y = net[:2](x) if rank == 1 else net[:3](x)
        y.pow(2).sum().backward()
        red.finish()
        dist.destroy_process_group()
        print("ok", rank)
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29533", str(script)], capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("ok") == 2


def test_wgrad_tiling_heuristic_and_model_bounds():
    """Host logic only: when a weight gradient is left unsplit (the C library then tiles it 128 wide) and when the
    contraction is sliced; the n_embd bound of the fused row kernels is enforced at construction."""
    from visualrwkv_b200 import fused
    from visualrwkv_b200.model import RWKV, default_args
    rows = 16384
    assert fused._ksplit(4, 768, 768, rows) == 1        # 36 tiles at N = 256 -> 72 tiles at N = 128: no split
    assert fused._ksplit(2, 768, 3072, rows) == 2       # 72 tiles < 74 pairs and already >= 60: sliced once
    assert fused._ksplit(1, 65536, 768, rows) == 1      # head weight gradient: 768 tiles
    assert fused._ksplit(4, 768, 128, rows) > 1         # LoRA down-weight gradients: 12 tiles, sliced
    assert fused._ksplit(4, 768, 128, 192) in (1, 2)    # never slices below whole k-blocks
    assert fused.gemm2_supported(1, 768, 768) is False and fused.gemm2_supported(8, 768, 768) is True
    import pytest
    with pytest.raises(ValueError):
        RWKV(default_args(n_embd=2560, n_layer=1, dim_att=2560))
