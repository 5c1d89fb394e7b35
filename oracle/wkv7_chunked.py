"""Chunk-parallel restatement of the WKV7 recurrence (TEST INFRASTRUCTURE ONLY).

The CUDA tensor-core kernels (visualrwkv_b200/csrc/wkv7_chunk_*.cu) evaluate the reference recurrence
(VisualRWKV-v7/v7.00/cuda/wkv7_cuda.cu:17-51) chunk by chunk with matrix products; this file states exactly that
algebra in torch so that every intermediate the kernels exchange can be checked, and checks it against the step-by-step
oracle (oracle/wkv7.py).  Nothing here is imported by the product.

Per head, S[i][j] (i: value index, j: key index), d_t = exp(-exp(w_t)):
    sa_t = S_{t-1} a_t ;  S_t = S_{t-1} diag(d_t) + sa_t b_t^T + v_t k_t^T ;  y_t = S_t q_t
Within a chunk of L steps, with G_t = sum_{s<=t} -exp(w_s) (G_0 = 0):
    At = a_t exp(G_{t-1})   Qt = q_t exp(G_t)   Kt = k_t exp(-G_t)   Bt = b_t exp(-G_t)
    A_ab = stril(At Bt^T)  A_ak = stril(At Kt^T)  A_qb = tril(Qt Bt^T)  A_qk = tril(Qt Kt^T)
    U = At S_0^T + A_ab U + A_ak V            (rows of U are the sa_t)
    Y = Qt S_0^T + A_qb U + A_qk V
    S_L = (S_0 + U^T Bt + V^T Kt) diag(exp(G_L))
"""
from __future__ import annotations

import torch


def tf32_round(x: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest-even to a 10-bit mantissa (what the tensor core keeps of an fp32 operand is a truncation on
    some parts; rounding is the optimistic model, `tf32_trunc` the pessimistic one)."""
    xi = x.float().contiguous().view(torch.int32)
    r = ((xi >> 13) & 1) + 0xFFF
    return ((xi + r) & ~0x1FFF).view(torch.float32).to(x.dtype)


def tf32_trunc(x: torch.Tensor) -> torch.Tensor:
    xi = x.float().contiguous().view(torch.int32)
    return (xi & ~0x1FFF).view(torch.float32).to(x.dtype)


def _ident(x):
    return x


def chunk_forward(w, q, k, v, a, b, L=64, dtype=torch.float64, rnd=_ident, s0=None, want=False):
    """w..b: [B,T,H,N] (any float dtype).  Returns y, sa [B,T,H,N] and S at every chunk start [B,H,T/L,N,N] (not
    transposed), plus the final state.  `rnd` is applied to every matrix-product operand."""
    w, q, k, v, a, b = [x.to(dtype).permute(0, 2, 1, 3) for x in (w, q, k, v, a, b)]  # [B,H,T,N]
    B, H, T, N = w.shape
    nc = T // L
    S = torch.zeros(B, H, N, N, dtype=dtype) if s0 is None else s0.to(dtype).clone()
    ys, sas, s_start = [], [], []
    mm = lambda x, y: rnd(x) @ rnd(y)
    inter = []
    for c in range(nc):
        sl = slice(c * L, (c + 1) * L)
        g = -torch.exp(w[:, :, sl])
        G = torch.cumsum(g, dim=2)
        Gm = G - g  # G_{t-1}
        At, Qt = a[:, :, sl] * torch.exp(Gm), q[:, :, sl] * torch.exp(G)
        Kt, Bt = k[:, :, sl] * torch.exp(-G), b[:, :, sl] * torch.exp(-G)
        V = v[:, :, sl]
        A_ab = torch.tril(mm(At, Bt.transpose(-1, -2)), -1)
        A_ak = torch.tril(mm(At, Kt.transpose(-1, -2)), -1)
        A_qb = torch.tril(mm(Qt, Bt.transpose(-1, -2)))
        A_qk = torch.tril(mm(Qt, Kt.transpose(-1, -2)))
        I = torch.eye(L, dtype=dtype)
        # the S_0-independent part
        AV = mm(A_ak, V)
        Ahat = torch.linalg.solve_triangular(I - A_ab, At, upper=False)   # T^-1 At
        Uhat = torch.linalg.solve_triangular(I - A_ab, AV, upper=False)   # T^-1 A_ak V
        Qp = Qt + mm(A_qb, Ahat)
        Yi = mm(A_qk, V) + mm(A_qb, Uhat)
        EL = torch.exp(G[:, :, -1:, :])                                    # [B,H,1,N]
        Bb, Kb = Bt * EL, Kt * EL                                          # b_t exp(G_L - G_t) <= |b_t|
        # the sequential part
        s_start.append(S.clone())
        U = mm(Ahat, S.transpose(-1, -2)) + Uhat
        Y = mm(Qp, S.transpose(-1, -2)) + Yi
        S = S * EL + mm(U.transpose(-1, -2), Bb) + mm(V.transpose(-1, -2), Kb)
        ys.append(Y)
        sas.append(U)
        if want:
            inter.append(dict(A_ab=A_ab, A_ak=A_ak, A_qb=A_qb, A_qk=A_qk, Ahat=Ahat, Uhat=Uhat, Qp=Qp, Yi=Yi, Bb=Bb,
                              Kb=Kb, EL=EL, At=At, Qt=Qt, Kt=Kt, Bt=Bt, G=G))
    y = torch.cat(ys, dim=2).permute(0, 2, 1, 3)
    sa = torch.cat(sas, dim=2).permute(0, 2, 1, 3)
    out = (y, sa, torch.stack(s_start, dim=2), S)
    return out + (inter,) if want else out


def chunk_backward(w, q, k, v, a, b, dy, L=64, dtype=torch.float64, rnd=_ident):
    """Hand-derived reverse pass of `chunk_forward` (the algebra the backward kernels follow).  Returns
    dw,dq,dk,dv,da,db [B,T,H,N]."""
    _, sa, s_start, _ = chunk_forward(w, q, k, v, a, b, L=L, dtype=dtype, rnd=rnd)
    w, q, k, v, a, b, dy, sa = [x.to(dtype).permute(0, 2, 1, 3) for x in (w, q, k, v, a, b, dy, sa)]
    B, H, T, N = w.shape
    nc = T // L
    mm = lambda x, y: rnd(x) @ rnd(y)
    tr = lambda x: x.transpose(-1, -2)
    dS = torch.zeros(B, H, N, N, dtype=dtype)
    grads = {n: [None] * nc for n in "wqkvab"}
    I = torch.eye(L, dtype=dtype)
    for c in reversed(range(nc)):
        sl = slice(c * L, (c + 1) * L)
        g = -torch.exp(w[:, :, sl])
        G = torch.cumsum(g, dim=2)
        Gm = G - g
        E, Em, F = torch.exp(G), torch.exp(Gm), torch.exp(-G)
        At, Qt, Kt, Bt = a[:, :, sl] * Em, q[:, :, sl] * E, k[:, :, sl] * F, b[:, :, sl] * F
        V, U, dY, S0 = v[:, :, sl], sa[:, :, sl], dy[:, :, sl], s_start[:, :, c]
        EL = E[:, :, -1:, :]
        A_ab = torch.tril(mm(At, tr(Bt)), -1)
        A_ak = torch.tril(mm(At, tr(Kt)), -1)
        A_qb = torch.tril(mm(Qt, tr(Bt)))
        A_qk = torch.tril(mm(Qt, tr(Kt)))
        # S_L = Z diag(E_L),  Z = S_0 + U^T Bt + V^T Kt
        Z = S0 + mm(tr(U), Bt) + mm(tr(V), Kt)
        dZ = dS * EL
        dGL = (dS * Z * EL).sum(dim=2, keepdim=True)          # [B,H,1,N]: d/dG_L
        dU = mm(Bt, tr(dZ))
        dBt = mm(U, dZ)
        dV = mm(Kt, tr(dZ))
        dKt = mm(V, dZ)
        dS0 = dZ.clone()
        # Y = Qt S_0^T + A_qb U + A_qk V
        dQt = mm(dY, S0)
        dS0 = dS0 + mm(tr(dY), Qt)
        dA_qb = torch.tril(mm(dY, tr(U)))
        dA_qk = torch.tril(mm(dY, tr(V)))
        dU = dU + mm(tr(A_qb), dY)
        dV = dV + mm(tr(A_qk), dY)
        # U = T^-1 R,  R = At S_0^T + A_ak V
        dR = torch.linalg.solve_triangular(tr(I - A_ab), dU, upper=True)
        dA_ab = torch.tril(mm(dR, tr(U)), -1)
        dA_ak = torch.tril(mm(dR, tr(V)), -1)
        dAt = mm(dR, S0)
        dS0 = dS0 + mm(tr(dR), At)
        dV = dV + mm(tr(A_ak), dR)
        # score matrices
        dAt = dAt + mm(dA_ab, Bt) + mm(dA_ak, Kt)
        dBt = dBt + mm(tr(dA_ab), At) + mm(tr(dA_qb), Qt)
        dKt = dKt + mm(tr(dA_ak), At) + mm(tr(dA_qk), Qt)
        dQt = dQt + mm(dA_qb, Bt) + mm(dA_qk, Kt)
        # scalings
        da, dq, dk, db = dAt * Em, dQt * E, dKt * F, dBt * F
        dG = dQt * Qt - dKt * Kt - dBt * Bt                     # d/dG_t through E_t, F_t
        dGm = dAt * At                                          # d/dG_{t-1}
        dG[:, :, :-1] += dGm[:, :, 1:]                          # (G_0 = 0 is a constant)
        dG[:, :, -1:] += dGL
        dg = torch.flip(torch.cumsum(torch.flip(dG, dims=[2]), dim=2), dims=[2])
        dw = dg * g
        for n, t in zip("wqkvab", (dw, dq, dk, dV, da, db)):
            grads[n][c] = t
        dS = dS0
    return tuple(torch.cat(grads[n], dim=2).permute(0, 2, 1, 3) for n in "wqkvab")
