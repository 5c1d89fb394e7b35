"""Parity of the CTA-pair tcgen05 GEMM (csrc/gemm2_sm100.cu) against fp32 torch: every operand layout, epilogue, the
grouped launch, the split contraction with its self-cleaning workspace, and the transposed store."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(c, ref):
    return ((c.float() - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()


def mk(shape, scale=1.0):
    return (scale * torch.randn(*shape, device="cuda")).to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K", [(256, 128, 64), (384, 768, 768), (1568, 768, 768), (512, 256, 1024)])
@pytest.mark.parametrize("layout", ["tn", "nn", "tt"])
def test_layouts_and_epilogues(M, N, K, layout):
    from visualrwkv_b200 import fused
    torch.manual_seed(M + N + K)
    lay = {"tn": fused.G2_TN, "nn": fused.G2_NN, "tt": fused.G2_TT}[layout]
    a = mk((K, M) if layout == "tt" else (M, K), 0.5)
    b = mk((N, K) if layout == "tn" else (K, N), K ** -0.5)
    A = a.t() if layout == "tt" else a
    B = b.t() if layout == "tn" else b
    ref = A.float() @ B.float()
    assert rel(fused.gemm2(a, b, lay), ref) < 2.5e-3                      # bf16 output rounding: 1.65e-3 RMS
    r = mk((M, N))
    assert rel(fused.gemm2(a, b, lay, fused.EPI_ADD, r), ref + r.float()) < 2.5e-3
    if layout == "tn":
        assert rel(fused.gemm2(a, b, lay, fused.EPI_RELU_SQ), torch.relu(ref.to(torch.bfloat16).float()) ** 2) < 3e-3
    if layout == "nn":
        act = mk((M, N)).abs()
        want = ref.to(torch.bfloat16).float() * 2 * act.float().sqrt()
        assert rel(fused.gemm2(a, b, lay, fused.EPI_RELUSQ_BWD, act), want) < 3e-3
    if layout == "tt" and K % 256 == 0:
        for ks in (2, 4):
            assert rel(fused.gemm2(a, b, lay, ksplit=ks), ref) < 2.5e-3
        assert rel(fused.gemm2(a, b, lay, ksplit=2), ref) < 2.5e-3         # the workspace was left zeroed by the calls above


def test_grouped_and_transposed():
    from visualrwkv_b200 import fused
    torch.manual_seed(3)
    xs = [mk((1024, 256), 0.5) for _ in range(4)]
    ws = [mk((256, 256), 1 / 16) for _ in range(4)]
    for c, x, w in zip(fused.gemm2_grouped(xs[:3], ws[:3]), xs, ws):
        assert rel(c, x.float() @ w.float().t()) < 2.5e-3
    gs = fused.gemm2_grouped(xs, xs[::-1], fused.G2_TT, ksplit=2)
    for i, g in enumerate(gs):
        assert rel(g, xs[i].float().t() @ xs[3 - i].float()) < 2.5e-3
    c0, c1 = fused.gemm2_grouped(xs[:2], xs[2:], fused.G2_TT, transposed=[0, 1])
    assert c1.shape == (256, 256)
    assert rel(c0, xs[0].float().t() @ xs[2].float()) < 2.5e-3
    assert rel(c1, (xs[1].float().t() @ xs[3].float()).t()) < 2.5e-3


def test_argument_checks():
    from visualrwkv_b200 import fused
    with pytest.raises(RuntimeError):
        fused.gemm2(mk((128, 64)), mk((100, 64)))        # N % 8 != 0
    with pytest.raises(RuntimeError):
        fused.gemm2_grouped([mk((128, 100))], [mk((128, 100))])   # ragged extents must be multiples of 8 (16-byte TMA strides)
    a, b = mk((128, 96), 0.5), mk((72, 96), 0.1)
    (c,) = fused.gemm2_grouped([a], [b])                 # K = 96, N = 72: tiled as (128, 128), zero-filled / clipped by TMA
    assert c.shape == (128, 72) and rel(c, a.float() @ b.float().t()) < 2.5e-3


@pytest.mark.parametrize("rows,C,ranks", [(1024, 768, (64, 64, 128, 32)), (512, 2048, (96, 96, 256, 64)), (256, 768, (64, 32, 160))])
def test_lora_branch_shapes(rows, C, ranks):
    """The LoRA launches of the time-mix block: groups of different rank in one launch (rank as N, as K and as M, tensors
    unpadded), activation epilogues and their backward from the saved output, weight gradients under a split contraction."""
    from visualrwkv_b200 import fused
    torch.manual_seed(rows + C)
    acts = [fused.ACT_TANH, fused.ACT_NONE, fused.ACT_SIGMOID, fused.ACT_NONE][:len(ranks)]
    fn = [torch.tanh, lambda t: t, torch.sigmoid, lambda t: t]
    xs = [mk((rows, C), 0.5) for _ in ranks]
    downs = [mk((C, r), C ** -0.5) for r in ranks]
    ups = [mk((r, C), r ** -0.5) for r in ranks]
    hs = fused.gemm2_grouped(xs, downs, fused.G2_NN, fused.EPI_ACT, acts=acts)
    for h, x, d, f, r in zip(hs, xs, downs, fn, ranks):
        assert h.shape == (rows, r)
        assert rel(h, f((x.float() @ d.float()).to(torch.bfloat16).float())) < 2.5e-3
    outs = fused.gemm2_grouped(hs, ups, fused.G2_NN)
    for o, h, u in zip(outs, hs, ups):
        assert rel(o, h.float() @ u.float()) < 2.5e-3
    douts = [mk((rows, C), 0.5) for _ in ranks]
    dps = fused.gemm2_grouped(douts, ups, fused.G2_TN, fused.EPI_ACT_BWD, residuals=hs, acts=acts)
    for dp, do, u, h, ac in zip(dps, douts, ups, hs, acts):
        dh = (do.float() @ u.float().t()).to(torch.bfloat16).float()
        hf = h.float()
        want = dh * {fused.ACT_TANH: 1 - hf * hf, fused.ACT_SIGMOID: hf * (1 - hf), fused.ACT_NONE: torch.ones_like(hf)}[ac]
        assert dp.shape == h.shape and rel(dp, want) < 3e-3
    dxs = fused.gemm2_grouped(dps, downs, fused.G2_TN)
    for dx, dp, d in zip(dxs, dps, downs):
        assert rel(dx, dp.float() @ d.float().t()) < 2.5e-3
    Rp = max(ranks)
    for ks in (1, 2):   # transposed store from a split contraction (the LoRA up-weight gradient as (dout^T h)^T)
        dUt = fused.gemm2_grouped(douts, hs, fused.G2_TT, ksplit=ks, transposed=[1] * len(ranks))
        for g, h, do, r in zip(dUt, hs, douts, ranks):
            assert g.shape == (r, C) and rel(g, h.float().t() @ do.float()) < 2.5e-3
    for ks in (1, fused._ksplit(len(ranks), Rp, C, rows)):
        dU = fused.gemm2_grouped(hs, douts, fused.G2_TT, ksplit=ks)
        dD = fused.gemm2_grouped(xs, dps, fused.G2_TT, ksplit=ks)
        for g, h, do, r in zip(dU, hs, douts, ranks):
            assert g.shape == (r, C) and rel(g, h.float().t() @ do.float()) < 2.5e-3
        for g, x, dp, r in zip(dD, xs, dps, ranks):
            assert g.shape == (C, r) and rel(g, x.float().t() @ dp.float()) < 2.5e-3


@pytest.mark.parametrize("layout,M,N,K", [("tt", 65536, 128, 256), ("tn", 16384, 128, 128), ("nn", 32768, 128, 64)])
def test_many_tiles_per_cta_with_one_store_chunk_each(layout, M, N, K):
    """N = 128 tiles are one 64-column store chunk per epilogue warp; several tiles per CTA then reuse the staging buffers
    back to back (the head weight gradient of a C = 128 model, the LoRA launches).  Checked element-wise: a staging
    buffer rewritten under a bulk store in flight shows up as whole wrong 32 x 64 blocks."""
    from visualrwkv_b200 import fused
    torch.manual_seed(7)
    lay = {"tn": fused.G2_TN, "nn": fused.G2_NN, "tt": fused.G2_TT}[layout]
    a = mk((K, M) if layout == "tt" else (M, K), 0.5)
    b = mk((N, K) if layout == "tn" else (K, N), K ** -0.5)
    ref = (a.t() if layout == "tt" else a).float() @ (b.t() if layout == "tn" else b).float()
    for _ in range(5):
        c = fused.gemm2(a, b, lay)
        assert (c.float() - ref).abs().max().item() < 0.02 * ref.abs().max().item() + 1e-3
