"""CPU tests pinning the oracle: spec loop, exact adjoint, and the reference-kernel
goldens minted on a B200 (tests/golden/README.md)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import wkv7 as O

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "wkv7_ref_*.npz")))


def _spec_loop(w, q, k, v, a, b):
    """fp64 loop in the form of VisualRWKV-v6/v6.xx/RWKV-v7_simple.py:20-32 (matmul form)."""
    B, T, H, N = w.shape
    wd = torch.exp(-torch.exp(w.double()))
    st = torch.zeros(B, H, N, N, dtype=torch.double)
    out = torch.zeros(B, T, H, N, dtype=torch.double)
    for t in range(T):
        sa = st @ a[:, t].double().unsqueeze(-1)
        st = st * wd[:, t, :, None, :] + sa @ b[:, t].double().unsqueeze(-2) \
            + v[:, t].double().unsqueeze(-1) @ k[:, t].double().unsqueeze(-2)
        out[:, t] = (st @ q[:, t].double().unsqueeze(-1)).squeeze(-1)
    return out.numpy(), st.numpy()


@pytest.mark.parametrize("kind", ["realistic", "stress"])
def test_forward_matches_spec_loop(kind):
    inp = O.make_inputs(2, 32, 3, 64, seed=5, kind=kind)
    y, s, sa, sT = O.forward(*inp[:6], want_final_state=True)
    ys, st = _spec_loop(*inp[:6])
    np.testing.assert_allclose(y, ys, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(sT, st, rtol=1e-12, atol=1e-13)
    # checkpoint layout: s[b,h,c,j,i] = S_ij  (wkv7_cuda.cu:44-50)
    np.testing.assert_allclose(s[:, :, -1].transpose(0, 1, 3, 2), st, rtol=1e-12, atol=1e-13)


def test_small_head_size_and_non64():
    # the oracle is generic in N (the spec script uses N=4)
    g = torch.Generator().manual_seed(0)
    w, q, k, v, a, b = [torch.randn(1, 16, 2, 4, generator=g).to(torch.bfloat16) for _ in range(6)]
    w = -torch.nn.functional.softplus(-w.float()).to(torch.bfloat16) - 0.5
    y, s, sa = O.forward(w, q, k, v, a, b)
    ys, _ = _spec_loop(w, q, k, v, a, b)
    np.testing.assert_allclose(y, ys, rtol=1e-11, atol=1e-12)


def test_f32_tracks_f64():
    inp = O.make_inputs(2, 64, 2, 64, seed=9)
    y64, s64, sa64 = O.forward(*inp[:6])
    y32, s32, sa32 = O.forward(*inp[:6], precision="f32")
    assert O.err_ratio(y32, y64) < 2e-6 and O.err_ratio(s32, s64) < 2e-6 and O.err_ratio(sa32, sa64) < 2e-6


@pytest.mark.parametrize("kind", ["realistic", "stress"])
def test_reference_backward_equals_exact_adjoint(kind):
    inp = O.make_inputs(2, 48, 2, 64, seed=3, kind=kind)
    y, s, sa = O.forward(*inp[:6])
    g = O.backward(*inp, s, sa)
    ge = O.backward_exact(*inp)
    for x, e in zip(g, ge):
        assert O.err_ratio(x, e) < 1e-12


def test_backward_matches_torch_autograd():
    inp = O.make_inputs(1, 16, 2, 64, seed=1)
    w, q, k, v, a, b, dy = [x.double().requires_grad_(True) for x in inp]
    B, T, H, N = w.shape
    st = torch.zeros(B, H, N, N, dtype=torch.double)
    ys = []
    for t in range(T):
        wd = torch.exp(-torch.exp(w[:, t]))
        sa = st @ a[:, t].unsqueeze(-1)
        st = st * wd[:, :, None, :] + sa @ b[:, t].unsqueeze(-2) + v[:, t].unsqueeze(-1) @ k[:, t].unsqueeze(-2)
        ys.append((st @ q[:, t].unsqueeze(-1)).squeeze(-1))
    y = torch.stack(ys, 1)
    y.backward(dy.detach())
    _, s, sa_ = O.forward(*inp[:6])
    g = O.backward(*inp, s, sa_)
    for x, t in zip(g, (w, q, k, v, a, b)):
        np.testing.assert_allclose(x, t.grad.numpy(), rtol=1e-9, atol=1e-11)


def test_state_carry_equals_one_shot():
    inp = O.make_inputs(1, 64, 2, 64, seed=2)
    y, s, sa, sT = O.forward(*inp[:6], want_final_state=True)
    h1 = [x[:, :32].contiguous() for x in inp[:6]]
    h2 = [x[:, 32:].contiguous() for x in inp[:6]]
    y1, _, _, s1 = O.forward(*h1, want_final_state=True)
    y2, _, _, s2 = O.forward(*h2, s0=s1, want_final_state=True)
    np.testing.assert_allclose(np.concatenate([y1, y2], 1), y, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(s2, sT, rtol=1e-12, atol=1e-13)


def test_input_generator_is_deterministic_and_in_domain():
    a1 = O.make_inputs(2, 32, 3, 64, seed=7)
    a2 = O.make_inputs(2, 32, 3, 64, seed=7)
    for x, y in zip(a1, a2):
        assert torch.equal(x, y)
    w = a1[0].float()
    assert (w <= -0.5).all()  # decay in (0.545, 1): model.py:176
    assert all(x.dtype == torch.bfloat16 and x.is_contiguous() for x in a1)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_reproduces_reference_kernel_goldens(path):
    """The oracle is pinned against outputs of the reference kernel itself (run on a B200)."""
    z = np.load(path)
    B, T, H = int(z["B"]), int(z["T"]), int(z["H"])
    inp = O.make_inputs(B, T, H, 64, seed=int(z["seed"]), kind=str(z["kind"]))
    bf = lambda name: torch.from_numpy(z[name]).view(torch.bfloat16).float().numpy()
    # fp32 emulation of the reference arithmetic: fp32 quantities at rtol 1e-3 / atol 1e-5 (north star)
    y32, s32, sa32 = O.forward(*inp[:6], precision="f32")
    np.testing.assert_allclose(sa32, z["sa"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(s32[:, :, -1], z["s_last"], rtol=1e-3, atol=1e-5)
    # bf16 outputs: <= 1 bf16 ulp except where summation order flips a rounding (rare), RMS tiny
    ulp = O.bf16_ulp_diff(O.to_bf16_f32(y32), bf("y"))
    assert (ulp <= 1).mean() > 0.999 and O.err_ratio(O.to_bf16_f32(y32), bf("y")) < 5e-4
    # fp64 ground truth vs the reference's bf16 outputs: only bf16 rounding separates them
    y64, s64, sa64 = O.forward(*inp[:6])
    assert O.err_ratio(bf("y"), y64) < 2.5e-3
    g64 = O.backward(*inp, s64, sa64)
    for name, ref in zip(["dw", "dq", "dk", "dv", "da", "db"], g64):
        assert O.err_ratio(bf(name), ref) < 2.5e-3, name


def test_goldens_present():
    assert len(GOLDEN) >= 2


@pytest.mark.parametrize("L", [16, 64])
def test_chunked_restatement_equals_step_by_step_oracle(L):
    """oracle/wkv7_chunked.py (the algebra of the tensor-core kernels) against the step-by-step fp64 oracle."""
    import torch
    from oracle import wkv7_chunked as C
    cpu = O.make_inputs(2, 128, 2, 64, seed=13)
    y64, s64, sa64 = O.forward(*cpu[:6])
    g64 = O.backward_exact(*cpu)
    y, sa, s_start, s_end = C.chunk_forward(*cpu[:6], L=L, dtype=torch.float64)
    assert O.err_ratio(y.numpy(), y64) < 1e-12 and O.err_ratio(sa.numpy(), sa64) < 1e-12
    # state at the start of chunk c = transposed checkpoint after step c*L - 1
    for c in range(1, 128 // L):
        ck = s64[:, :, c * L // 16 - 1].transpose(0, 1, 3, 2)
        assert O.err_ratio(s_start[:, :, c].numpy(), ck) < 1e-12
    g = C.chunk_backward(*cpu, L=L, dtype=torch.float64)
    for x, r in zip(g, g64):
        assert O.err_ratio(x.numpy(), np.asarray(r)) < 1e-11
    # TF32 operand rounding (what the kernels do) stays well inside the output rounding of bf16
    yt, sat, _, _ = C.chunk_forward(*cpu[:6], L=64, dtype=torch.float32, rnd=C.tf32_round)
    assert O.err_ratio(yt.numpy(), y64) < 8e-4 and O.err_ratio(sat.numpy(), sa64) < 8e-4
