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
// Vector width of the LayerNorm kernels: 4 channels per thread (C/4 threads per row).  With 8 the backward needs
// ~230 registers per thread and only 6 warps fit on an SM (ncu: profiles/r1a_fused_ln_bwd); 4 halves every
// per-thread array and triples the resident warps.
constexpr int LN_VW = 4;
struct FV {
    float v[LN_VW];
};
__device__ __forceinline__ FV zerov() {
    FV r;
#pragma unroll
    for (int e = 0; e < LN_VW; e++) r.v[e] = 0.f;
    return r;
}
__device__ __forceinline__ FV ldv(bool active, const uint16_t* p) {
    if (!active) return zerov();
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    FV r;
    r.v[0] = bf16lo_to_f32(u.x); r.v[1] = bf16hi_to_f32(u.x);
    r.v[2] = bf16lo_to_f32(u.y); r.v[3] = bf16hi_to_f32(u.y);
    return r;
}
__device__ __forceinline__ void stv(bool active, uint16_t* p, const FV& r) {
    if (!active) return;
    uint2 u;
    u.x = pack_bf16x2(r.v[0], r.v[1]);
    u.y = pack_bf16x2(r.v[2], r.v[3]);
    *reinterpret_cast<uint2*>(p) = u;
}
__host__ __device__ inline int ln_threads(int C) { return ((C / LN_VW + 31) / 32) * 32; }
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
__device__ __forceinline__ void ln_stats(const FV (&x)[NB], bool active, float inv_c, float eps, float* red, int& phase,
                                         int nwarps, float (&mean)[NB], float (&rstd)[NB]) {
    float s[NB];
#pragma unroll
    for (int i = 0; i < NB; i++) {
        s[i] = 0.f;
#pragma unroll
        for (int e = 0; e < LN_VW; e++) s[i] += x[i].v[e];
    }
    block_sum<NB>(s, red, phase, nwarps);
    float q[NB];
#pragma unroll
    for (int i = 0; i < NB; i++) {
        mean[i] = s[i] * inv_c;
        q[i] = 0.f;
#pragma unroll
        for (int e = 0; e < LN_VW; e++) {
            const float d = active ? x[i].v[e] - mean[i] : 0.f;  // padding threads (c0 >= C) must not add mean^2
            q[i] += d * d;
        }
    }
    block_sum<NB>(q, red, phase, nwarps);
#pragma unroll
    for (int i = 0; i < NB; i++) rstd[i] = rsqrtf(q[i] * inv_c + eps);
}

template <int NMIX>
__global__ void __launch_bounds__(512) ln_mix_fwd_kernel(const LnMixFwdArgs a) {
    __shared__ float red[2 * 4 * 32];
    int phase = 0;
    const int tid = threadIdx.x, nwarps = (blockDim.x + 31) >> 5;
    const int c0 = tid * LN_VW;
    const bool active = c0 < a.C;
    const float inv_c = 1.f / a.C;
    const bool do_ln = a.gamma != nullptr;  // gamma == NULL: the input is already normalised (mix only)
    const FV g = ldv(active && do_ln, a.gamma + c0), b = ldv(active && do_ln, a.beta + c0);
    const int row0 = blockIdx.x * LN_RUN;
    const int row1 = min(row0 + LN_RUN, a.rows);

    auto normalize = [&](const FV& x, float mean, float rstd) {
        if (!do_ln) return x;
        FV h;
#pragma unroll
        for (int e = 0; e < LN_VW; e++) h.v[e] = rb((x.v[e] - mean) * rstd * g.v[e] + b.v[e]);
        return h;
    };
    FV cf[NMIX > 0 ? NMIX : 1];
#pragma unroll
    for (int m = 0; m < NMIX; m++) cf[m] = ldv(active, a.coef[m] + c0);

    FV hprev = zerov();
    if (NMIX > 0 && row0 < a.rows && (row0 % a.T) != 0) {  // halo: LN of the row before the run
        FV xp[1] = {ldv(active, a.x + (size_t)(row0 - 1) * a.C + c0)};
        float mean[1] = {0.f}, rstd[1] = {1.f};
        if (do_ln) ln_stats<1>(xp, active, inv_c, a.eps, red, phase, nwarps, mean, rstd);
        hprev = normalize(xp[0], mean[0], rstd[0]);
    }
    for (int row = row0; row < row1; row += 4) {
        FV x[4];
#pragma unroll
        for (int i = 0; i < 4; i++) x[i] = ldv(active && row + i < row1, a.x + (size_t)(row + i) * a.C + c0);
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
            const FV h = normalize(x[i], mean[i], rstd[i]);
            if (a.h_out) stv(active, a.h_out + (size_t)(row + i) * a.C + c0, h);
            if constexpr (NMIX > 0) {
                if (((row + i) % a.T) == 0) hprev = zerov();  // time_shift pads with zeros at t = 0
                FV xx;
#pragma unroll
                for (int e = 0; e < LN_VW; e++) xx.v[e] = rb(hprev.v[e] - h.v[e]);
#pragma unroll
                for (int m = 0; m < NMIX; m++) {
                    FV o;
#pragma unroll
                    for (int e = 0; e < LN_VW; e++) o.v[e] = h.v[e] + rb(xx.v[e] * cf[m].v[e]);
                    stv(active, a.out[m] + (size_t)(row + i) * a.C + c0, o);
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

// MODE 0: dx and parameter-gradient partials in one pass.  MODE 1: dx only.  MODE 2: parameter partials only (no block
// reductions, no barriers: rows stream through).  With six mixes the one-pass kernel is latency-bound (one CTA per SM at
// 171 registers, two barriers per row pair); running MODE 1 + MODE 2 re-reads the inputs once but both passes stream.
template <int NMIX, int MAXT, int MODE>
__global__ void __launch_bounds__(MAXT) ln_mix_bwd_kernel(const LnMixBwdArgs a) {
    __shared__ float red[2 * 4 * 32];
    int phase = 0;
    const int tid = threadIdx.x, nwarps = (blockDim.x + 31) >> 5;
    const int c0 = tid * LN_VW;
    const bool active = c0 < a.C;
    const float inv_c = 1.f / a.C;
    const bool do_ln = a.gamma != nullptr;
    FV g = ldv(active && do_ln, a.gamma + c0);
    const FV b = ldv(active && do_ln, a.beta + c0);
    if (!do_ln) {
#pragma unroll
        for (int e = 0; e < LN_VW; e++) g.v[e] = 1.f;
    }
    const int row0 = blockIdx.x * LN_RUN;
    const int row1 = min(row0 + LN_RUN, a.rows);

    FV dgam = zerov(), dbet = zerov();
    FV dco[NMIX > 0 ? NMIX : 1];
#pragma unroll
    for (int m = 0; m < NMIX; m++) dco[m] = zerov();

    struct Row {  // what is needed to finish a row once the gradient of its LN output is known
        FV xh, dh, res;
        float rstd;
        int row;
        bool valid;
    };
    // x-hat of a row from its (already loaded) values and statistics
    auto xhat = [&](const FV& x, float mean, float rstd) {
        if (!do_ln) return x;
        FV xh;
#pragma unroll
        for (int e = 0; e < LN_VW; e++) xh.v[e] = (x.v[e] - mean) * rstd;
        return xh;
    };
    // LayerNorm backward of up to two rows with ONE block reduction (4 sums); residual rows are preloaded in R.res
    auto finish2 = [&](const Row& A, const Row& B) {
        const Row* R[2] = {&A, &B};
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        FV dxh[2];
#pragma unroll
        for (int i = 0; i < 2; i++) {
#pragma unroll
            for (int e = 0; e < LN_VW; e++) {
                // an invalid row (the empty pending slot of the first iteration, the clamped row past the end) contributes
                // exact zeros: its x-hat may be anything, and 0 * inf/NaN is not 0
                const float dh = R[i]->valid ? R[i]->dh.v[e] : 0.f;
                const float xh = R[i]->valid ? R[i]->xh.v[e] : 0.f;
                dxh[i].v[e] = dh * g.v[e];
                if (do_ln) {
                    s[2 * i] += dxh[i].v[e];
                    s[2 * i + 1] += dxh[i].v[e] * xh;
                    if (MODE != 1) {
                        dgam.v[e] += dh * xh;
                        dbet.v[e] += dh;
                    }
                }
            }
        }
        if (MODE == 2) return;
        if (do_ln) block_sum<4>(s, red, phase, nwarps);
#pragma unroll
        for (int i = 0; i < 2; i++) {
            if (!R[i]->valid) continue;
            FV dx;
            if (do_ln) {
                const float m1 = s[2 * i] * inv_c, m2 = s[2 * i + 1] * inv_c;
#pragma unroll
                for (int e = 0; e < LN_VW; e++) dx.v[e] = R[i]->rstd * (dxh[i].v[e] - m1 - R[i]->xh.v[e] * m2);
            } else {
                dx = R[i]->dh;
            }
#pragma unroll
            for (int e = 0; e < LN_VW; e++) dx.v[e] += R[i]->res.v[e];
            stv(active, a.dx + (size_t)R[i]->row * a.C + c0, dx);
        }
    };
    const int last = a.rows - 1;
    auto ld_res = [&](int row) { return ldv(active && MODE != 2 && a.dresid != nullptr, a.dresid + (size_t)min(row, last) * a.C + c0); };
    auto ld_stat = [&](int row, float& mean, float& rstd) {
        mean = 0.f; rstd = 1.f;
        if (do_ln) {
            const float2 st = *reinterpret_cast<const float2*>(a.stats + 2 * (size_t)min(row, last));
            mean = st.x; rstd = st.y;
        }
    };

    if constexpr (NMIX == 0) {
        for (int row = row0; row < row1; row += 2) {
            // ---- load phase: everything the iteration needs, issued back to back (clamped addresses, no branches)
            const int r1 = min(row + 1, last);
            const FV x0 = ldv(active, a.x + (size_t)row * a.C + c0), x1 = ldv(active, a.x + (size_t)r1 * a.C + c0);
            Row A, B;
            A.dh = ldv(active, a.dh + (size_t)row * a.C + c0);
            B.dh = ldv(active, a.dh + (size_t)r1 * a.C + c0);
            A.res = ld_res(row);
            B.res = ld_res(r1);
            float m0, m1;
            ld_stat(row, m0, A.rstd);
            ld_stat(r1, m1, B.rstd);
            // ---- compute
            A.row = row; A.valid = true; A.xh = xhat(x0, m0, A.rstd);
            B.row = row + 1; B.valid = row + 1 < row1; B.xh = xhat(x1, m1, B.rstd);
            finish2(A, B);
        }
    } else {
        // out_m[t] = h[t] + (h[t-1] - h[t]) c_m  =>  dh[t] = P[t] - Q[t] + Q[t+1],  P = sum_m dout_m, Q = sum_m dout_m c_m.
        // Rows are visited two at a time; a row is finished one iteration late, when Q of its successor is known, so each
        // iteration finishes (previous pending row, first row of the pair) with one block reduction.  All global loads of
        // an iteration are issued up front with clamped addresses (ncu r1b: 8 serialized load round trips per iteration
        // when they were predicated and consumed one by one).
        FV cf[NMIX];
#pragma unroll
        for (int m = 0; m < NMIX; m++) cf[m] = ldv(active, a.coef[m] + c0);
        FV hprev = zerov();
        if (row0 < a.rows && (row0 % a.T) != 0) {
            float mean, rstd;
            ld_stat(row0 - 1, mean, rstd);
            const FV xh = xhat(ldv(active, a.x + (size_t)(row0 - 1) * a.C + c0), mean, rstd);
#pragma unroll
            for (int e = 0; e < LN_VW; e++) hprev.v[e] = rb(xh.v[e] * g.v[e] + b.v[e]);
        }
        Row pend;
        pend.valid = false;
        pend.row = row0;
        pend.xh = zerov();
        pend.dh = zerov();
        pend.res = zerov();
        pend.rstd = 0.f;
        FV Dp = zerov();  // P - Q of the pending row
        const int niter = (row1 - row0 + 2) / 2;  // pairs, plus a final pass for the halo row when the run length is even
        for (int it = 0; it < niter; it++) {
            const int row = row0 + 2 * it;
            const bool in0 = row < row1, in1 = row + 1 < row1;
            const bool ex0 = row < a.rows && (in0 || (row % a.T) != 0);
            const bool ex1 = row + 1 < a.rows && (in1 || ((row + 1) % a.T) != 0);
            if (!in0 && !pend.valid) break;
            // ---- load phase
            const int c0r = min(row, last), c1r = min(row + 1, last);
            const FV x0 = ldv(active, a.x + (size_t)c0r * a.C + c0), x1 = ldv(active, a.x + (size_t)c1r * a.C + c0);
            FV d0[NMIX], d1[NMIX];
#pragma unroll
            for (int m = 0; m < NMIX; m++) {
                d0[m] = ldv(active, a.dout[m] + (size_t)c0r * a.C + c0);
                d1[m] = ldv(active, a.dout[m] + (size_t)c1r * a.C + c0);
            }
            Row A = pend, B;
            A.res = ld_res(pend.row);
            B.res = ld_res(c0r);
            float m0, m1, rstd1;
            ld_stat(c0r, m0, B.rstd);
            ld_stat(c1r, m1, rstd1);
            // ---- compute: P, Q of both rows; dcoef accumulation for the rows of the run
            FV P0 = zerov(), Q0 = zerov(), P1 = zerov(), Q1 = zerov();
            const FV xh0 = xhat(x0, m0, B.rstd), xh1 = xhat(x1, m1, rstd1);
            FV xx0 = zerov(), xx1 = zerov(), h0 = zerov(), h1;
            if (in0) {
                if ((row % a.T) == 0) hprev = zerov();
#pragma unroll
                for (int e = 0; e < LN_VW; e++) {
                    h0.v[e] = rb(xh0.v[e] * g.v[e] + b.v[e]);
                    xx0.v[e] = rb(hprev.v[e] - h0.v[e]);
                }
                hprev = h0;
            }
            if (in1) {
                if (((row + 1) % a.T) == 0) hprev = zerov();
#pragma unroll
                for (int e = 0; e < LN_VW; e++) {
                    h1.v[e] = rb(xh1.v[e] * g.v[e] + b.v[e]);
                    xx1.v[e] = rb(hprev.v[e] - h1.v[e]);
                }
                hprev = h1;
            }
#pragma unroll
            for (int m = 0; m < NMIX; m++) {
#pragma unroll
                for (int e = 0; e < LN_VW; e++) {
                    const float a0 = ex0 ? d0[m].v[e] : 0.f, a1 = ex1 ? d1[m].v[e] : 0.f;
                    P0.v[e] += a0; Q0.v[e] += a0 * cf[m].v[e];
                    P1.v[e] += a1; Q1.v[e] += a1 * cf[m].v[e];
                    if (MODE != 1) dco[m].v[e] += (in0 ? a0 * xx0.v[e] : 0.f) + (in1 ? a1 * xx1.v[e] : 0.f);
                }
            }
            if (A.valid) {
                const bool same = ex0 && (row % a.T) != 0;
#pragma unroll
                for (int e = 0; e < LN_VW; e++) A.dh.v[e] = Dp.v[e] + (same ? Q0.v[e] : 0.f);
            }
            B.row = row; B.valid = in0; B.xh = xh0;
            const bool same1 = ex1 && ((row + 1) % a.T) != 0;
#pragma unroll
            for (int e = 0; e < LN_VW; e++) B.dh.v[e] = P0.v[e] - Q0.v[e] + (same1 ? Q1.v[e] : 0.f);
            finish2(A, B);
            pend.valid = in1;
            pend.row = in1 ? row + 1 : pend.row;
            pend.xh = xh1;
            pend.rstd = rstd1;
#pragma unroll
            for (int e = 0; e < LN_VW; e++) Dp.v[e] = P1.v[e] - Q1.v[e];
        }
    }
    if (!active || MODE == 1) return;
    float* dst = a.partial + (size_t)blockIdx.x * (2 + NMIX) * a.C + c0;
#pragma unroll
    for (int e = 0; e < LN_VW; e++) {
        dst[e] = dgam.v[e];
        dst[a.C + e] = dbet.v[e];
    }
#pragma unroll
    for (int m = 0; m < NMIX; m++)
#pragma unroll
        for (int e = 0; e < LN_VW; e++) dst[(size_t)(2 + m) * a.C + e] = dco[m].v[e];
}

}  // namespace vrwkv

using namespace vrwkv;

static int ln_check(int rows, int T, int C, int nmix) {
    if (rows <= 0 || T <= 0 || C <= 0 || rows % T) return vrwkv_fail(VRWKV_EINVAL, "ln_mix: bad rows/T (%d,%d)", rows, T);
    if (C % 8 || C / LN_VW > 512) return vrwkv_fail(VRWKV_EUNSUP, "ln_mix: C=%d must be a multiple of 8 and <= 2048", C);
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
    const dim3 grid(vrwkv_ln_mix_blocks(rows)), block(ln_threads(C));
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
    const dim3 grid(vrwkv_ln_mix_blocks(rows)), block(ln_threads(C));
    switch (nmix) {
#define VRWKV_LN_BWD(nm, mode)                                                                           \
    do {                                                                                                \
        if (block.x <= 256) ln_mix_bwd_kernel<nm, 256, mode><<<grid, block, 0, (cudaStream_t)stream>>>(a); \
        else ln_mix_bwd_kernel<nm, 512, mode><<<grid, block, 0, (cudaStream_t)stream>>>(a);             \
    } while (0)
        case 0: VRWKV_LN_BWD(0, 0); break;
        case 1: VRWKV_LN_BWD(1, 1); VRWKV_LN_BWD(1, 2); vrwkv_count_launch(1); break;
        case 6: VRWKV_LN_BWD(6, 1); VRWKV_LN_BWD(6, 2); vrwkv_count_launch(1); break;
#undef VRWKV_LN_BWD
        default: return vrwkv_fail(VRWKV_EUNSUP, "ln_mix_backward: nmix must be 0, 1 or 6");
    }
    VRWKV_CUDA(cudaGetLastError());
    vrwkv_count_launch(1);
    return VRWKV_OK;
}
