"""tcgen05 GEMM (visualrwkv_b200/csrc/gemm_sm100.cu) against a plain PyTorch fp32 reference of the same op:
C = epilogue(A @ W^T) with bf16 operands, fp32 accumulation, bf16 output (the Linear layers of model.py:157-160,
216-217 and their fused relu^2 / residual epilogues, model.py:225-226,194)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, w, epi, res):
    c = a.float() @ w.float().t()
    if epi == 1:  # relu(key(x))**2 in bf16 eager: the Linear output is rounded to bf16 before the square (model.py:225)
        c = torch.relu(c.to(torch.bfloat16).float()) ** 2
    if epi == 2:
        c = c + res.float()
    return c


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 768), (1000, 768, 768), (4096, 3072, 768),
                                   (16384, 768, 3072), (77, 128, 128)])
@pytest.mark.parametrize("epi", [0, 1, 2])
def test_gemm_matches_fp32_reference(M, N, K, epi):
    from visualrwkv_b200 import fused
    assert fused.gemm_supported(M, N, K)
    g = torch.Generator(device="cuda").manual_seed(M + N + K + epi)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) * (K ** -0.5)).to(torch.bfloat16)
    res = torch.randn(M, N, device="cuda", generator=g).to(torch.bfloat16) if epi == 2 else None
    c = fused.gemm_tn(a, w, epi, res)
    ref = _ref(a, w, epi, res)
    # one bf16 rounding of an fp32-accumulated result: |err| <= 2^-8 |ref| (+ a little for accumulation order)
    err = (c.float() - ref).abs()
    # (relu^2: a pre-activation that sits on a bf16 rounding boundary may round the other way (accumulation order) ->
    #  one bf16 ulp of c, i.e. 2 ulp-fractions of c^2, on top of the final rounding; rare, so the RMS bound below stays tight)
    assert float((err - (2.0 ** -8) * ref.abs() * (6.0 if epi == 1 else 1.0)).max()) <= 2e-3
    assert float(err.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt().clamp_min(1e-12)) < 2.5e-3


def test_gemm_rejects_unsupported_shapes():
    from visualrwkv_b200 import fused
    a = torch.zeros(64, 96, device="cuda", dtype=torch.bfloat16)  # K not a multiple of 64
    w = torch.zeros(128, 96, device="cuda", dtype=torch.bfloat16)
    assert not fused.gemm_supported(64, 128, 96)
    with pytest.raises(RuntimeError):
        fused.gemm_tn(a, w)
