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
        fused.gemm2(mk((128, 96)), mk((128, 96)))        # K % 64 != 0
    with pytest.raises(RuntimeError):
        fused.gemm2(mk((128, 64)), mk((100, 64)))        # N % 128 != 0
