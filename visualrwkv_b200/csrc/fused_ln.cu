// fused_ln.cu — LayerNorm fused with the RWKV token-shift mixes (forward + backward).
//
// Replaces, per Block (VisualRWKV-v7/v7.00/src/model.py):
//   ln1 -> time_shift -> six lerps xr,xw,xk,xv,xa,xg      (:250, :149, :166-173)     NMIX = 6
//   ln2 -> time_shift -> one lerp k                         (:252, :205, :222-224)     NMIX = 1
//   plain LayerNorm (ln0, ln_out, proj.ln_v)                (:248, :323, :338)         NMIX = 0
// which the reference runs as ~20 separate bf16 eager kernels, each streaming a (B,T,C) tensor.
// Here: one pass reads x once and writes the NMIX mixed streams (and mean/rstd for the backward).
// bf16 rounding points of the eager graph are reproduced (h, xx = shift(h)-h, xx*c, h + xx*c each rounded to
// bf16) so the outputs equal the reference module's bit for bit up to LayerNorm's fp32 summation order.
//
// Mapping: see rowops.cuh (CTA = run of RUN rows, thread = 8 channels).  HBM bytes per element (bf16):
// forward 2 + 2*NMIX, backward 2*NMIX + 2 (x) [+ 2 residual grad] + 2 (dx).
#include "host_util.h"
#include "rowops.cuh"

namespace vrwkv {

constexpr int LN_MAXMIX = 6;
constexpr int LN_RUN = 32;

struct LnMixFwdArgs {
    int rows, T, C, nmix;
    float eps;
    const uint16_t* x;
    const uint16_t *gamma, *beta;
    const uint16_t* coef[LN_MAXMIX];
    uint16_t* out[LN_MAXMIX];
    uint16_t* h_out;  // optional: LN output itself
    float* stats;     // [rows][2] mean, rstd
};

// mean / rstd of NB rows at once: two block reductions for NB rows instead of two per row
template <int NB>
__device__ __forceinline__ void ln_stats(const F8 (&x)[NB], bool active, float inv_c, float eps, float* red, int& phase,
                                         int nwarps, float (&mean)[NB], float (&rstd)[NB]) {
    float s[NB];
#pragma unroll
    for (int i = 0; i < NB; i++) {
        s[i] = 0.f;
#pragma unroll
        for (int e = 0; e < 8; e++) s[i] += x[i].v[e];
    }
    block_sum<NB>(s, red, phase, nwarps);
    float q[NB];
#pragma unroll
    for (int i = 0; i < NB; i++) {
        mean[i] = s[i] * inv_c;
        q[i] = 0.f;
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const float d = active ? x[i].v[e] - mean[i] : 0.f;  // padding threads (c0 >= C) must not add mean^2
            q[i] += d * d;
        }
    }
    block_sum<NB>(q, red, phase, nwarps);
#pragma unroll
    for (int i = 0; i < NB; i++) rstd[i] = rsqrtf(q[i] * inv_c + eps);
}

template <int NMIX>
__global__ void __launch_bounds__(256) ln_mix_fwd_kernel(const LnMixFwdArgs a) {
    __shared__ float red[2 * 4 * 32];
    int phase = 0;
    const int tid = threadIdx.x, nwarps = (blockDim.x + 31) >> 5;
    const int c0 = tid * 8;
    const bool active = c0 < a.C;
    const float inv_c = 1.f / a.C;
    const bool do_ln = a.gamma != nullptr;  // gamma == NULL: the input is already normalised (mix only)
    const F8 g = ldz(active && do_ln, a.gamma + c0), b = ldz(active && do_ln, a.beta + c0);
    const int row0 = blockIdx.x * LN_RUN;
    const int row1 = min(row0 + LN_RUN, a.rows);

    auto normalize = [&](const F8& x, float mean, float rstd) {
        if (!do_ln) return x;
        F8 h;
#pragma unroll
        for (int e = 0; e < 8; e++) h.v[e] = rb((x.v[e] - mean) * rstd * g.v[e] + b.v[e]);
        return h;
    };
    F8 cf[NMIX > 0 ? NMIX : 1];
#pragma unroll
    for (int m = 0; m < NMIX; m++) cf[m] = ldz(active, a.coef[m] + c0);

    F8 hprev = zero8();
    if (NMIX > 0 && row0 < a.rows && (row0 % a.T) != 0) {  // halo: LN of the row before the run
        F8 xp[1] = {ldz(active, a.x + (size_t)(row0 - 1) * a.C + c0)};
        float mean[1] = {0.f}, rstd[1] = {1.f};
        if (do_ln) ln_stats<1>(xp, active, inv_c, a.eps, red, phase, nwarps, mean, rstd);
        hprev = normalize(xp[0], mean[0], rstd[0]);
    }
    for (int row = row0; row < row1; row += 4) {
        F8 x[4];
#pragma unroll
        for (int i = 0; i < 4; i++) x[i] = ldz(active && row + i < row1, a.x + (size_t)(row + i) * a.C + c0);
        float mean[4] = {0.f, 0.f, 0.f, 0.f}, rstd[4] = {1.f, 1.f, 1.f, 1.f};
        if (do_ln) {
            ln_stats<4>(x, active, inv_c, a.eps, red, phase, nwarps, mean, rstd);
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (tid == i && row + i < row1 && a.stats) {
                    a.stats[2 * (row + i)] = mean[i];
                    a.stats[2 * (row + i) + 1] = rstd[i];
                }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (row + i >= row1) break;
            const F8 h = normalize(x[i], mean[i], rstd[i]);
            if (a.h_out) stz(active, a.h_out + (size_t)(row + i) * a.C + c0, h);
            if constexpr (NMIX > 0) {
                if (((row + i) % a.T) == 0) hprev = zero8();  // time_shift pads with zeros at t = 0
                F8 xx;
#pragma unroll
                for (int e = 0; e < 8; e++) xx.v[e] = rb(hprev.v[e] - h.v[e]);
#pragma unroll
                for (int m = 0; m < NMIX; m++) {
                    F8 o;
#pragma unroll
                    for (int e = 0; e < 8; e++) o.v[e] = h.v[e] + rb(xx.v[e] * cf[m].v[e]);
                    stz(active, a.out[m] + (size_t)(row + i) * a.C + c0, o);
                }
                hprev = h;
            }
        }
    }
}

struct LnMixBwdArgs {
    int rows, T, C, nmix;
    const uint16_t* x;
    const float* stats;
    const uint16_t *gamma, *beta;
    const uint16_t* coef[LN_MAXMIX];
    const uint16_t* dout[LN_MAXMIX];
    const uint16_t* dh;       // nmix == 0: gradient of the LN output
    const uint16_t* dresid;   // optional: added to dx (residual stream gradient)
    uint16_t* dx;
    float* partial;           // [gridDim.x][2 + nmix][C]: dgamma, dbeta, dcoef[m]
};

template <int NMIX>
__global__ void __launch_bounds__(256) ln_mix_bwd_kernel(const LnMixBwdArgs a) {
    __shared__ float red[2 * 4 * 32];
    int phase = 0;
    const int tid = threadIdx.x, nwarps = (blockDim.x + 31) >> 5;
    const int c0 = tid * 8;
    const bool active = c0 < a.C;
    const float inv_c = 1.f / a.C;
    const bool do_ln = a.gamma != nullptr;
    F8 g = ldz(active && do_ln, a.gamma + c0);
    const F8 b = ldz(active && do_ln, a.beta + c0);
    if (!do_ln) {
#pragma unroll
        for (int e = 0; e < 8; e++) g.v[e] = 1.f;
    }
    const int row0 = blockIdx.x * LN_RUN;
    const int row1 = min(row0 + LN_RUN, a.rows);

    F8 dgam = zero8(), dbet = zero8();
    F8 dco[NMIX > 0 ? NMIX : 1];
#pragma unroll
    for (int m = 0; m < NMIX; m++) dco[m] = zero8();

    struct Row {  // what is needed to finish a row once the gradient of its LN output is known
        F8 xh, dh;
        float rstd;
        int row;
        bool valid;
    };
    auto xhat_of = [&](int row, F8& xh) {
        const F8 x = ldz(active, a.x + (size_t)row * a.C + c0);
        if (!do_ln) {
            xh = x;
            return 1.f;
        }
        const float mean = a.stats[2 * row], rstd = a.stats[2 * row + 1];
#pragma unroll
        for (int e = 0; e < 8; e++) xh.v[e] = (x.v[e] - mean) * rstd;
        return rstd;
    };
    // LayerNorm backward of up to two rows with ONE block reduction (4 sums)
    auto finish2 = [&](const Row& A, const Row& B) {
        const Row* R[2] = {&A, &B};
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        F8 dxh[2];
#pragma unroll
        for (int i = 0; i < 2; i++) {
            if (!R[i]->valid) continue;
#pragma unroll
            for (int e = 0; e < 8; e++) {
                dxh[i].v[e] = R[i]->dh.v[e] * g.v[e];
                if (do_ln) {
                    s[2 * i] += dxh[i].v[e];
                    s[2 * i + 1] += dxh[i].v[e] * R[i]->xh.v[e];
                    dgam.v[e] += R[i]->dh.v[e] * R[i]->xh.v[e];
                    dbet.v[e] += R[i]->dh.v[e];
                }
            }
        }
        if (do_ln) block_sum<4>(s, red, phase, nwarps);
#pragma unroll
        for (int i = 0; i < 2; i++) {
            if (!R[i]->valid) continue;
            F8 dx;
            if (do_ln) {
                const float m1 = s[2 * i] * inv_c, m2 = s[2 * i + 1] * inv_c;
#pragma unroll
                for (int e = 0; e < 8; e++) dx.v[e] = R[i]->rstd * (dxh[i].v[e] - m1 - R[i]->xh.v[e] * m2);
            } else {
                dx = R[i]->dh;
            }
            if (a.dresid) {
                const F8 r = ldz(active, a.dresid + (size_t)R[i]->row * a.C + c0);
#pragma unroll
                for (int e = 0; e < 8; e++) dx.v[e] += r.v[e];
            }
            stz(active, a.dx + (size_t)R[i]->row * a.C + c0, dx);
        }
    };

    if constexpr (NMIX == 0) {
        for (int row = row0; row < row1; row += 2) {
            Row A, B;
            A.row = row; A.valid = true; A.rstd = xhat_of(row, A.xh); A.dh = ldz(active, a.dh + (size_t)row * a.C + c0);
            B.row = row + 1; B.valid = row + 1 < row1;
            if (B.valid) { B.rstd = xhat_of(row + 1, B.xh); B.dh = ldz(active, a.dh + (size_t)(row + 1) * a.C + c0); }
            finish2(A, B);
        }
    } else {
        // out_m[t] = h[t] + (h[t-1] - h[t]) c_m  =>  dh[t] = P[t] - Q[t] + Q[t+1],  P = sum_m dout_m, Q = sum_m dout_m c_m.
        // Rows are visited two at a time; a row is finished one iteration late, when Q of its successor is known, so each
        // iteration finishes (previous pending row, first row of the pair) with one block reduction.
        F8 cf[NMIX];
#pragma unroll
        for (int m = 0; m < NMIX; m++) cf[m] = ldz(active, a.coef[m] + c0);
        F8 hprev = zero8();
        if (row0 < a.rows && (row0 % a.T) != 0) {
            F8 xh;
            xhat_of(row0 - 1, xh);
#pragma unroll
            for (int e = 0; e < 8; e++) hprev.v[e] = rb(xh.v[e] * g.v[e] + b.v[e]);
        }
        // P and Q of one row (and, for rows of the run, x-hat / rstd / the dcoef accumulation)
        auto visit = [&](int row, bool in_run, F8& P, F8& Q, F8& xh, float& rstd) {
            P = zero8(); Q = zero8();
            const bool exists = row < a.rows && (in_run || (row % a.T) != 0);
            F8 xx = zero8();
            if (in_run) {
                rstd = xhat_of(row, xh);
                if ((row % a.T) == 0) hprev = zero8();
                F8 h;
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    h.v[e] = rb(xh.v[e] * g.v[e] + b.v[e]);
                    xx.v[e] = rb(hprev.v[e] - h.v[e]);
                }
                hprev = h;
            }
            if (exists) {
#pragma unroll
                for (int m = 0; m < NMIX; m++) {
                    const F8 d = ldz(active, a.dout[m] + (size_t)row * a.C + c0);
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        P.v[e] += d.v[e];
                        Q.v[e] += d.v[e] * cf[m].v[e];
                        if (in_run) dco[m].v[e] += d.v[e] * xx.v[e];
                    }
                }
            }
            return exists;
        };
        Row pend;
        pend.valid = false;
        F8 Dp = zero8();  // P - Q of the pending row
        for (int row = row0; row < row1; row += 2) {
            F8 P0, Q0, P1, Q1;
            Row A, B;  // A: pending row (finished now), B: first row of the pair (finished now); second row becomes pending
            float rstd1 = 1.f;
            F8 xh1 = zero8();
            B.row = row; B.valid = true;
            visit(row, true, P0, Q0, B.xh, B.rstd);
            const bool in1 = row + 1 < row1;
            const bool ex1 = visit(row + 1, in1, P1, Q1, xh1, rstd1);
            A = pend;
            if (A.valid) {
                const bool same = (row % a.T) != 0;
#pragma unroll
                for (int e = 0; e < 8; e++) A.dh.v[e] = Dp.v[e] + (same ? Q0.v[e] : 0.f);
            }
            const bool same1 = ex1 && ((row + 1) % a.T) != 0;
#pragma unroll
            for (int e = 0; e < 8; e++) B.dh.v[e] = P0.v[e] - Q0.v[e] + (same1 ? Q1.v[e] : 0.f);
            finish2(A, B);
            pend.valid = in1;
            if (in1) {
                pend.row = row + 1; pend.xh = xh1; pend.rstd = rstd1;
#pragma unroll
                for (int e = 0; e < 8; e++) Dp.v[e] = P1.v[e] - Q1.v[e];
            }
        }
        if (pend.valid) {  // last row of the run: needs Q of the halo row after the run
            F8 P, Q, xh;
            float rstd;
            const bool ex = visit(row1, false, P, Q, xh, rstd);
            const bool same = ex && (row1 % a.T) != 0;
#pragma unroll
            for (int e = 0; e < 8; e++) pend.dh.v[e] = Dp.v[e] + (same ? Q.v[e] : 0.f);
            Row none;
            none.valid = false;
            finish2(pend, none);
        }
    }
    if (!active) return;
    float* dst = a.partial + (size_t)blockIdx.x * (2 + NMIX) * a.C + c0;
#pragma unroll
    for (int e = 0; e < 8; e++) {
        dst[e] = dgam.v[e];
        dst[a.C + e] = dbet.v[e];
    }
#pragma unroll
    for (int m = 0; m < NMIX; m++)
#pragma unroll
        for (int e = 0; e < 8; e++) dst[(size_t)(2 + m) * a.C + e] = dco[m].v[e];
}

}  // namespace vrwkv

using namespace vrwkv;

static int ln_check(int rows, int T, int C, int nmix) {
    if (rows <= 0 || T <= 0 || C <= 0 || rows % T) return vrwkv_fail(VRWKV_EINVAL, "ln_mix: bad rows/T (%d,%d)", rows, T);
    if (C % 8 || C / 8 > 256) return vrwkv_fail(VRWKV_EUNSUP, "ln_mix: C=%d must be a multiple of 8 and <= 2048", C);
    if (nmix < 0 || nmix > LN_MAXMIX) return vrwkv_fail(VRWKV_EINVAL, "ln_mix: nmix=%d out of range", nmix);
    return VRWKV_OK;
}

extern "C" int vrwkv_ln_mix_blocks(int rows) { return (rows + LN_RUN - 1) / LN_RUN; }

extern "C" int vrwkv_ln_mix_forward(int rows, int T, int C, int nmix, float eps, const uint16_t* x, const uint16_t* gamma,
                                    const uint16_t* beta, const uint16_t* const* coef, uint16_t* const* out,
                                    uint16_t* h_out, float* stats, void* stream) {
    int rc = ln_check(rows, T, C, nmix);
    if (rc) return rc;
    if (!x || ((gamma == nullptr) != (beta == nullptr)) || (nmix == 0 && !h_out))
        return vrwkv_fail(VRWKV_EINVAL, "ln_mix_forward: null pointer");
    LnMixFwdArgs a{};
    a.rows = rows; a.T = T; a.C = C; a.nmix = nmix; a.eps = eps; a.x = x; a.gamma = gamma; a.beta = beta;
    for (int m = 0; m < nmix; m++) {
        if (!coef[m] || !out[m]) return vrwkv_fail(VRWKV_EINVAL, "ln_mix_forward: null mix pointer %d", m);
        a.coef[m] = coef[m];
        a.out[m] = out[m];
    }
    a.h_out = h_out;
    a.stats = stats;
    const dim3 grid(vrwkv_ln_mix_blocks(rows)), block(row_threads(C));
    switch (nmix) {
        case 0: ln_mix_fwd_kernel<0><<<grid, block, 0, (cudaStream_t)stream>>>(a); break;
        case 1: ln_mix_fwd_kernel<1><<<grid, block, 0, (cudaStream_t)stream>>>(a); break;
        case 6: ln_mix_fwd_kernel<6><<<grid, block, 0, (cudaStream_t)stream>>>(a); break;
        default: return vrwkv_fail(VRWKV_EUNSUP, "ln_mix_forward: nmix must be 0, 1 or 6");
    }
    VRWKV_CUDA(cudaGetLastError());
    vrwkv_count_launch(1);
    return VRWKV_OK;
}

extern "C" int vrwkv_ln_mix_backward(int rows, int T, int C, int nmix, const uint16_t* x, const float* stats,
                                     const uint16_t* gamma, const uint16_t* beta, const uint16_t* const* coef,
                                     const uint16_t* const* dout, const uint16_t* dh, const uint16_t* dresid, uint16_t* dx,
                                     float* partial, void* stream) {
    int rc = ln_check(rows, T, C, nmix);
    if (rc) return rc;
    if (!x || (gamma && !stats) || ((gamma == nullptr) != (beta == nullptr)) || !dx || !partial || (nmix == 0 && !dh))
        return vrwkv_fail(VRWKV_EINVAL, "ln_mix_backward: null pointer");
    LnMixBwdArgs a{};
    a.rows = rows; a.T = T; a.C = C; a.nmix = nmix; a.x = x; a.stats = stats; a.gamma = gamma; a.beta = beta;
    for (int m = 0; m < nmix; m++) {
        if (!coef[m] || !dout[m]) return vrwkv_fail(VRWKV_EINVAL, "ln_mix_backward: null mix pointer %d", m);
        a.coef[m] = coef[m];
        a.dout[m] = dout[m];
    }
    a.dh = dh; a.dresid = dresid; a.dx = dx; a.partial = partial;
    const dim3 grid(vrwkv_ln_mix_blocks(rows)), block(row_threads(C));
    switch (nmix) {
        case 0: ln_mix_bwd_kernel<0><<<grid, block, 0, (cudaStream_t)stream>>>(a); break;
        case 1: ln_mix_bwd_kernel<1><<<grid, block, 0, (cudaStream_t)stream>>>(a); break;
        case 6: ln_mix_bwd_kernel<6><<<grid, block, 0, (cudaStream_t)stream>>>(a); break;
        default: return vrwkv_fail(VRWKV_EUNSUP, "ln_mix_backward: nmix must be 0, 1 or 6");
    }
    VRWKV_CUDA(cudaGetLastError());
    vrwkv_count_launch(1);
    return VRWKV_OK;
}
