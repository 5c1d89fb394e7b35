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


// =====================================================================================================================
// Warp-per-row kernels (C <= 1024).  ncu (profiles/r2_rows_*): the CTA-per-row kernels above run at 1.2-1.8 TB/s of
// algorithmic bytes with the issue slots 45-60 % busy at 17-30 % occupancy — instruction-bound, not memory-bound: 4
// channels per thread amortise the per-row work (two block reductions through shared memory, address arithmetic, the
// t % T test) over too few elements, and every bf16 rounding point of the eager graph costs a convert + shift.
// Here a WARP owns a run of rows: lane l holds the NV 16-byte vectors at columns (j*32 + l)*8, the row sums are five
// shuffles (no shared memory, no barrier), and the token-shift mixes run in packed bf16x2 arithmetic
// (sub/mul/add.rn.bf16x2 round once per operation, exactly as the eager graph's bf16 ops do), two elements per instruction.
// =====================================================================================================================
constexpr int WR_RUN = 8;      // rows per warp (16: half the warps, and the kernels are latency-bound at ~7 warps per SM)
constexpr int WR_WARPS = 4;    // warps per CTA  -> 32 rows per CTA

__device__ __forceinline__ uint32_t& w4(uint4& v, int i) { return reinterpret_cast<uint32_t*>(&v)[i]; }
__device__ __forceinline__ uint32_t w4(const uint4& v, int i) { return reinterpret_cast<const uint32_t*>(&v)[i]; }

// shared-memory parameter table: [which][NV*32] 16-byte vectors, vector (j*32 + lane) = columns (j*32 + lane)*8 ..+8
template <int NV>
__device__ __forceinline__ void wr_stage_params(uint4* sm, int nwhich, const uint16_t* const* src, int C) {
    for (int i = threadIdx.x; i < nwhich * NV * 32; i += blockDim.x) {
        const int which = i / (NV * 32), col = (i % (NV * 32)) * 8;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (col < C && src[which]) v = __ldg(reinterpret_cast<const uint4*>(src[which] + col));
        sm[i] = v;
    }
    __syncthreads();
}

// LayerNorm of one row held as NV packed vectors per lane: statistics by warp shuffles, output packed bf16
template <int NV>
__device__ __forceinline__ void wr_ln_row(const uint4 (&x)[NV], const bool (&act)[NV], bool do_ln, float inv_c, float eps,
                                          const uint4* sm_g, const uint4* sm_b, int lane, float& mean, float& rstd, uint4 (&h)[NV]) {
    if (!do_ln) {
#pragma unroll
        for (int j = 0; j < NV; j++) h[j] = x[j];
        mean = 0.f; rstd = 1.f;
        return;
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NV; j++)
#pragma unroll
        for (int i = 0; i < 4; i++) s += bf16lo_to_f32(w4(x[j], i)) + bf16hi_to_f32(w4(x[j], i));
    mean = warp_sum(s) * inv_c;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NV; j++)
        if (act[j]) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float d0 = bf16lo_to_f32(w4(x[j], i)) - mean, d1 = bf16hi_to_f32(w4(x[j], i)) - mean;
                q += d0 * d0 + d1 * d1;
            }
        }
    rstd = rsqrtf(warp_sum(q) * inv_c + eps);
#pragma unroll
    for (int j = 0; j < NV; j++) {
        const uint4 g = sm_g[j * 32 + lane], b = sm_b[j * 32 + lane];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float h0 = (bf16lo_to_f32(w4(x[j], i)) - mean) * rstd * bf16lo_to_f32(w4(g, i)) + bf16lo_to_f32(w4(b, i));
            const float h1 = (bf16hi_to_f32(w4(x[j], i)) - mean) * rstd * bf16hi_to_f32(w4(g, i)) + bf16hi_to_f32(w4(b, i));
            w4(h[j], i) = pack_bf16x2(h0, h1);
        }
    }
}

template <int NMIX, int NV>
__global__ void __launch_bounds__(WR_WARPS * 32) ln_mix_fwd_wr_kernel(const LnMixFwdArgs a) {
    extern __shared__ uint4 wr_sm[];   // gamma, beta, coef[0..NMIX)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {
        const uint16_t* src[2 + LN_MAXMIX] = {a.gamma, a.beta};
#pragma unroll
        for (int m = 0; m < NMIX; m++) src[2 + m] = a.coef[m];
        wr_stage_params<NV>(wr_sm, 2 + NMIX, src, a.C);
    }
    const uint4 *sm_g = wr_sm, *sm_b = wr_sm + NV * 32;
    const bool do_ln = a.gamma != nullptr;
    const float inv_c = 1.f / a.C;
    bool act[NV];
    int col[NV];
#pragma unroll
    for (int j = 0; j < NV; j++) {
        col[j] = (j * 32 + lane) * 8;
        act[j] = col[j] < a.C;
    }
    const int row0 = (blockIdx.x * WR_WARPS + warp) * WR_RUN;
    const int row1 = min(row0 + WR_RUN, a.rows);
    if (row0 >= a.rows) return;
    auto load_row = [&](int row, uint4 (&x)[NV]) {
#pragma unroll
        for (int j = 0; j < NV; j++)
            x[j] = act[j] ? *reinterpret_cast<const uint4*>(a.x + (size_t)row * a.C + col[j]) : make_uint4(0u, 0u, 0u, 0u);
    };
    uint4 hprev[NV];
#pragma unroll
    for (int j = 0; j < NV; j++) hprev[j] = make_uint4(0u, 0u, 0u, 0u);
    int tpos = row0 % a.T;
    if (NMIX > 0 && tpos != 0) {   // halo: LN of the row before the run
        uint4 xp[NV];
        load_row(row0 - 1, xp);
        float mean, rstd;
        wr_ln_row<NV>(xp, act, do_ln, inv_c, a.eps, sm_g, sm_b, lane, mean, rstd, hprev);
    }
    uint4 xn[NV];
    load_row(row0, xn);
    for (int row = row0; row < row1; row++) {
        uint4 x[NV];
#pragma unroll
        for (int j = 0; j < NV; j++) x[j] = xn[j];
        if (row + 1 < row1) load_row(row + 1, xn);   // next row in flight while this one is normalised and mixed
        float mean, rstd;
        uint4 h[NV];
        wr_ln_row<NV>(x, act, do_ln, inv_c, a.eps, sm_g, sm_b, lane, mean, rstd, h);
        if (do_ln && a.stats && lane == 0) *reinterpret_cast<float2*>(a.stats + 2 * (size_t)row) = make_float2(mean, rstd);
        const size_t roff = (size_t)row * a.C;
        if (a.h_out) {
#pragma unroll
            for (int j = 0; j < NV; j++)
                if (act[j]) *reinterpret_cast<uint4*>(a.h_out + roff + col[j]) = h[j];
        }
        if constexpr (NMIX > 0) {
            if (tpos == 0) {   // time_shift pads with zeros at t = 0
#pragma unroll
                for (int j = 0; j < NV; j++) hprev[j] = make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int j = 0; j < NV; j++) {
                uint4 xx;
#pragma unroll
                for (int i = 0; i < 4; i++) w4(xx, i) = bf2_sub(w4(hprev[j], i), w4(h[j], i));
#pragma unroll
                for (int m = 0; m < NMIX; m++) {
                    const uint4 c = wr_sm[(2 + m) * NV * 32 + j * 32 + lane];
                    uint4 o;
#pragma unroll
                    for (int i = 0; i < 4; i++) w4(o, i) = bf2_add(w4(h[j], i), bf2_mul(w4(xx, i), w4(c, i)));
                    if (act[j]) *reinterpret_cast<uint4*>(a.out[m] + roff + col[j]) = o;
                }
                hprev[j] = h[j];
            }
        }
        tpos = (tpos + 1 == a.T) ? 0 : tpos + 1;
    }
}

// Backward, pass 1: dx (and, for the plain LayerNorm, the parameter partials).  A warp walks its run once; a row is
// finished one step late, when Q of its successor is known (out_m[t] = h[t] + (h[t-1] - h[t]) c_m  =>
// dh[t] = P[t] - Q[t] + Q[t+1],  P = sum_m dout_m, Q = sum_m dout_m c_m).
template <int NMIX, int NV>
__global__ void __launch_bounds__(WR_WARPS * 32) ln_mix_bwd_wr_kernel(const LnMixBwdArgs a) {
    extern __shared__ uint4 wr_sm[];   // gamma, beta, coef[0..NMIX); reused at the end for the CTA's partial sums
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {
        const uint16_t* src[2 + LN_MAXMIX] = {a.gamma, a.beta};
#pragma unroll
        for (int m = 0; m < NMIX; m++) src[2 + m] = a.coef[m];
        wr_stage_params<NV>(wr_sm, 2 + NMIX, src, a.C);
    }
    const bool do_ln = a.gamma != nullptr;
    const float inv_c = 1.f / a.C;
    bool act[NV];
    int col[NV];
#pragma unroll
    for (int j = 0; j < NV; j++) {
        col[j] = (j * 32 + lane) * 8;
        act[j] = col[j] < a.C;
    }
    const int row0 = (blockIdx.x * WR_WARPS + warp) * WR_RUN;
    const int row1 = min(row0 + WR_RUN, a.rows);
    float dgam[NV][8], dbet[NV][8];
#pragma unroll
    for (int j = 0; j < NV; j++)
#pragma unroll
        for (int e = 0; e < 8; e++) dgam[j][e] = dbet[j][e] = 0.f;

    if (row0 < a.rows) {
        float Dp[NV][8];       // P - Q of the pending row
        uint4 xpend[NV];       // its x (packed), statistics
        float mean_p = 0.f, rstd_p = 1.f;
        bool pending = false;
        int tpos = row0 % a.T;
        auto ld = [&](const uint16_t* base, int row, int j) {
            return act[j] ? *reinterpret_cast<const uint4*>(base + (size_t)row * a.C + col[j]) : make_uint4(0u, 0u, 0u, 0u);
        };
        // finish row r: dh known -> LayerNorm backward, residual gradient, store
        auto finish = [&](int r, const float (&dh)[NV][8], const uint4 (&xr)[NV], float mean, float rstd, const uint4 (&res)[NV]) {
            float dxh[NV][8], xh[NV][8];
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int j = 0; j < NV; j++) {
                const uint4 g = wr_sm[j * 32 + lane];
#pragma unroll
                for (int i = 0; i < 4; i++) {
#pragma unroll
                    for (int hl = 0; hl < 2; hl++) {
                        const int e = 2 * i + hl;
                        const float xv = hl ? bf16hi_to_f32(w4(xr[j], i)) : bf16lo_to_f32(w4(xr[j], i));
                        const float gv = do_ln ? (hl ? bf16hi_to_f32(w4(g, i)) : bf16lo_to_f32(w4(g, i))) : 1.f;
                        xh[j][e] = do_ln ? (xv - mean) * rstd : xv;
                        dxh[j][e] = dh[j][e] * gv;
                        if (do_ln) {
                            s1 += dxh[j][e];
                            s2 += dxh[j][e] * xh[j][e];
                            if constexpr (NMIX == 0) {   // with mixes the parameter sums are the second kernel's
                                dgam[j][e] += dh[j][e] * xh[j][e];
                                dbet[j][e] += dh[j][e];
                            }
                        }
                    }
                }
            }
            float m1 = 0.f, m2 = 0.f;
            if (do_ln) {
                m1 = warp_sum(s1) * inv_c;
                m2 = warp_sum(s2) * inv_c;
            }
#pragma unroll
            for (int j = 0; j < NV; j++) {
                uint4 o;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    float d0, d1;
                    if (do_ln) {
                        d0 = rstd * (dxh[j][2 * i] - m1 - xh[j][2 * i] * m2);
                        d1 = rstd * (dxh[j][2 * i + 1] - m1 - xh[j][2 * i + 1] * m2);
                    } else {
                        d0 = dh[j][2 * i];
                        d1 = dh[j][2 * i + 1];
                    }
                    d0 += bf16lo_to_f32(w4(res[j], i));
                    d1 += bf16hi_to_f32(w4(res[j], i));
                    w4(o, i) = pack_bf16x2(d0, d1);
                }
                if (act[j]) *reinterpret_cast<uint4*>(a.dx + (size_t)r * a.C + col[j]) = o;
            }
        };
        if constexpr (NMIX == 0) {
            for (int row = row0; row < row1; row++) {
                uint4 x[NV], d[NV], res[NV];
#pragma unroll
                for (int j = 0; j < NV; j++) {
                    x[j] = ld(a.x, row, j);
                    d[j] = ld(a.dh, row, j);
                    res[j] = a.dresid ? ld(a.dresid, row, j) : make_uint4(0u, 0u, 0u, 0u);
                }
                float mean = 0.f, rstd = 1.f;
                if (do_ln) {
                    const float2 st = *reinterpret_cast<const float2*>(a.stats + 2 * (size_t)row);
                    mean = st.x; rstd = st.y;
                }
                float dh[NV][8];
#pragma unroll
                for (int j = 0; j < NV; j++)
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        dh[j][2 * i] = bf16lo_to_f32(w4(d[j], i));
                        dh[j][2 * i + 1] = bf16hi_to_f32(w4(d[j], i));
                    }
                finish(row, dh, x, mean, rstd, res);
            }
        } else {
            // steps row0 .. row1 (the last one only supplies Q of the row after the run)
            for (int row = row0; row <= row1; row++) {
                const bool have = row < a.rows && (row < row1 || tpos != 0);   // this row's douts are needed
                if (!have && !pending) break;
                const int rr = min(row, a.rows - 1);
                uint4 d[NMIX][NV], x[NV], res[NV];
#pragma unroll
                for (int j = 0; j < NV; j++) {
#pragma unroll
                    for (int m = 0; m < NMIX; m++) d[m][j] = ld(a.dout[m], rr, j);
                    x[j] = ld(a.x, rr, j);
                    res[j] = (a.dresid && pending) ? ld(a.dresid, row - 1, j) : make_uint4(0u, 0u, 0u, 0u);
                }
                if (!have) {   // the row after the run belongs to another sequence (or does not exist): Q = 0
#pragma unroll
                    for (int j = 0; j < NV; j++)
#pragma unroll
                        for (int m = 0; m < NMIX; m++) d[m][j] = make_uint4(0u, 0u, 0u, 0u);
                }
                float mean = 0.f, rstd = 1.f;
                if (do_ln) {
                    const float2 st = *reinterpret_cast<const float2*>(a.stats + 2 * (size_t)rr);
                    mean = st.x; rstd = st.y;
                }
                float PQ[NV][8], Q[NV][8];
#pragma unroll
                for (int j = 0; j < NV; j++) {
#pragma unroll
                    for (int e = 0; e < 8; e++) PQ[j][e] = Q[j][e] = 0.f;
#pragma unroll
                    for (int m = 0; m < NMIX; m++) {
                        const uint4 c = wr_sm[(2 + m) * NV * 32 + j * 32 + lane];
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const float d0 = bf16lo_to_f32(w4(d[m][j], i)), d1 = bf16hi_to_f32(w4(d[m][j], i));
                            const float q0 = d0 * bf16lo_to_f32(w4(c, i)), q1 = d1 * bf16hi_to_f32(w4(c, i));
                            Q[j][2 * i] += q0; Q[j][2 * i + 1] += q1;
                            PQ[j][2 * i] += d0; PQ[j][2 * i + 1] += d1;
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 8; e++) PQ[j][e] -= Q[j][e];
                }
                if (pending) {
                    const bool same = have && tpos != 0;   // the successor belongs to the same sequence
                    float dh[NV][8];
#pragma unroll
                    for (int j = 0; j < NV; j++)
#pragma unroll
                        for (int e = 0; e < 8; e++) dh[j][e] = Dp[j][e] + (same ? Q[j][e] : 0.f);
                    finish(row - 1, dh, xpend, mean_p, rstd_p, res);
                }
                pending = row < row1;
#pragma unroll
                for (int j = 0; j < NV; j++) {
                    xpend[j] = x[j];
#pragma unroll
                    for (int e = 0; e < 8; e++) Dp[j][e] = PQ[j][e];
                }
                mean_p = mean; rstd_p = rstd;
                tpos = (tpos + 1 == a.T) ? 0 : tpos + 1;
            }
        }
    }
    if constexpr (NMIX != 0) return;
    // CTA partial of dgamma / dbeta: the warps' sums meet in shared memory (the parameter table is no longer needed)
    __syncthreads();
    float* acc = reinterpret_cast<float*>(wr_sm);   // [2][NV*256]
    for (int i = threadIdx.x; i < 2 * NV * 256; i += blockDim.x) acc[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NV; j++)
#pragma unroll
        for (int e = 0; e < 8; e++) {
            atomicAdd(&acc[(j * 32 + lane) * 8 + e], dgam[j][e]);
            atomicAdd(&acc[NV * 256 + (j * 32 + lane) * 8 + e], dbet[j][e]);
        }
    __syncthreads();
    float* dst = a.partial + (size_t)blockIdx.x * (2 + NMIX) * a.C;
    for (int i = threadIdx.x; i < 2 * NV * 256; i += blockDim.x) {
        const int which = i / (NV * 256), c = i % (NV * 256);
        if (c < a.C) dst[(size_t)which * a.C + c] = acc[i];
    }
}

// Backward, pass 2 (NMIX > 0): every parameter partial — dgamma = sum_t dh xhat, dbeta = sum_t dh,
// dcoef_m = sum_t dout_m[t] (h[t-1] - h[t]).  No row reductions (the statistics come from the forward), so a thread owns
// 8 columns; the CTA's 64 rows are split over G thread groups (each walks 64/G rows, a row's dh again finished one step
// late) whose sums meet in shared memory.  Partials: [gridDim.x][2 + NMIX][C].
template <int VW>
struct PV {   // VW packed bf16 columns
    uint32_t w[VW / 2];
};
template <int VW>
__device__ __forceinline__ PV<VW> pv_ld(const uint16_t* p) {
    PV<VW> r;
    if constexpr (VW == 8) {
        const uint4 u = *reinterpret_cast<const uint4*>(p);
        r.w[0] = u.x; r.w[1] = u.y; r.w[2] = u.z; r.w[3] = u.w;
    } else {
        const uint2 u = *reinterpret_cast<const uint2*>(p);
        r.w[0] = u.x; r.w[1] = u.y;
    }
    return r;
}

// VW columns per thread: 8 for the one-mix layer, 4 for the six-mix one (64 instead of 32 accumulators would leave one
// 6-warp CTA per SM at 224 registers).
template <int NMIX, int G, int VW, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) ln_mix_bwd_params_kernel(const LnMixBwdArgs a) {
    extern __shared__ float pk_sm[];   // [(2 + NMIX) * VW][tpg]
    constexpr int NACC = 2 + NMIX, GROWS = WR_RUN * WR_WARPS / G, NW = VW / 2;
    const int tpg = blockDim.x / G, grp = threadIdx.x / tpg, ct = threadIdx.x % tpg;
    const int c0 = ct * VW;
    const bool active = c0 < a.C;
    const bool do_ln = a.gamma != nullptr;
    const int b1 = min((int)(blockIdx.x + 1) * (WR_RUN * WR_WARPS), a.rows);
    const int r0 = blockIdx.x * (WR_RUN * WR_WARPS) + grp * GROWS, r1 = min(r0 + GROWS, b1);
    float acc[NACC][VW];
#pragma unroll
    for (int k = 0; k < NACC; k++)
#pragma unroll
        for (int e = 0; e < VW; e++) acc[k][e] = 0.f;
    if (active && r0 < r1) {
        PV<VW> gp, bp, cp[NMIX];   // parameters stay packed (unpacked where used)
#pragma unroll
        for (int i = 0; i < NW; i++) gp.w[i] = bp.w[i] = 0u;
        if (do_ln) {
            gp = pv_ld<VW>(a.gamma + c0);
            bp = pv_ld<VW>(a.beta + c0);
        }
#pragma unroll
        for (int m = 0; m < NMIX; m++) cp[m] = pv_ld<VW>(a.coef[m] + c0);
        // x-hat (fp32) and the packed LN output of a row's columns
        auto ln = [&](int row, float (&xh)[VW]) {
            const PV<VW> x = pv_ld<VW>(a.x + (size_t)row * a.C + c0);
            float2 st = make_float2(0.f, 1.f);
            if (do_ln) st = __ldg(reinterpret_cast<const float2*>(a.stats + 2 * (size_t)row));
            PV<VW> h;
#pragma unroll
            for (int i = 0; i < NW; i++) {
                const float x0 = bf16lo_to_f32(x.w[i]), x1 = bf16hi_to_f32(x.w[i]);
                if (do_ln) {
                    xh[2 * i] = (x0 - st.x) * st.y;
                    xh[2 * i + 1] = (x1 - st.x) * st.y;
                    h.w[i] = pack_bf16x2((x0 - st.x) * st.y * bf16lo_to_f32(gp.w[i]) + bf16lo_to_f32(bp.w[i]),
                                         (x1 - st.x) * st.y * bf16hi_to_f32(gp.w[i]) + bf16hi_to_f32(bp.w[i]));
                } else {
                    xh[2 * i] = x0; xh[2 * i + 1] = x1;
                    h.w[i] = x.w[i];
                }
            }
            return h;
        };
        int tpos = r0 % a.T;
        PV<VW> hprev;
#pragma unroll
        for (int i = 0; i < NW; i++) hprev.w[i] = 0u;
        float tmp[VW];
        if (tpos != 0) hprev = ln(r0 - 1, tmp);
        float Dp[VW], xhp[VW];
        bool pending = false;
        for (int row = r0; row <= r1; row++) {
            const bool have = row < a.rows && (row < r1 || tpos != 0);
            if (!have && !pending) break;
            const int rr = min(row, a.rows - 1);
            PV<VW> d[NMIX];
#pragma unroll
            for (int m = 0; m < NMIX; m++) d[m] = pv_ld<VW>(a.dout[m] + (size_t)rr * a.C + c0);
            if (!have) {
#pragma unroll
                for (int m = 0; m < NMIX; m++)
#pragma unroll
                    for (int i = 0; i < NW; i++) d[m].w[i] = 0u;
            }
            float xh[VW];
            const PV<VW> h = ln(rr, xh);
            if (tpos == 0) {
#pragma unroll
                for (int i = 0; i < NW; i++) hprev.w[i] = 0u;
            }
            float P[VW], Q[VW];
#pragma unroll
            for (int e = 0; e < VW; e++) P[e] = Q[e] = 0.f;
            const bool inrun = row < r1;
#pragma unroll
            for (int i = 0; i < NW; i++) {
                const uint32_t xx = bf2_sub(hprev.w[i], h.w[i]);
                const float x0 = bf16lo_to_f32(xx), x1 = bf16hi_to_f32(xx);
#pragma unroll
                for (int m = 0; m < NMIX; m++) {
                    const float d0 = bf16lo_to_f32(d[m].w[i]), d1 = bf16hi_to_f32(d[m].w[i]);
                    P[2 * i] += d0; P[2 * i + 1] += d1;
                    Q[2 * i] += d0 * bf16lo_to_f32(cp[m].w[i]); Q[2 * i + 1] += d1 * bf16hi_to_f32(cp[m].w[i]);
                    if (inrun) {
                        acc[2 + m][2 * i] += d0 * x0;
                        acc[2 + m][2 * i + 1] += d1 * x1;
                    }
                }
            }
            if (pending) {
                const bool same = have && tpos != 0;
#pragma unroll
                for (int e = 0; e < VW; e++) {
                    const float dh = Dp[e] + (same ? Q[e] : 0.f);
                    acc[0][e] += dh * xhp[e];
                    acc[1][e] += dh;
                }
            }
            pending = inrun;
#pragma unroll
            for (int e = 0; e < VW; e++) {
                Dp[e] = P[e] - Q[e];
                xhp[e] = xh[e];
            }
            hprev = h;
            tpos = (tpos + 1 == a.T) ? 0 : tpos + 1;
        }
    }
    // the groups' sums meet in shared memory, one group per round
    for (int gi = 0; gi < G; gi++) {
        if (grp == gi) {
#pragma unroll
            for (int k = 0; k < NACC; k++)
#pragma unroll
                for (int e = 0; e < VW; e++) {
                    float* p = &pk_sm[(k * VW + e) * tpg + ct];
                    *p = (gi == 0) ? acc[k][e] : *p + acc[k][e];
                }
        }
        __syncthreads();
    }
    if (grp == 0 && active) {
        float* dst = a.partial + (size_t)blockIdx.x * NACC * a.C + c0;
#pragma unroll
        for (int k = 0; k < NACC; k++)
#pragma unroll
            for (int e4 = 0; e4 < VW; e4 += 4)
                *reinterpret_cast<float4*>(dst + (size_t)k * a.C + e4) =
                    make_float4(pk_sm[(k * VW + e4) * tpg + ct], pk_sm[(k * VW + e4 + 1) * tpg + ct], pk_sm[(k * VW + e4 + 2) * tpg + ct],
                                pk_sm[(k * VW + e4 + 3) * tpg + ct]);
    }
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
// number of partial rows the backward writes for this shape (warp-per-row kernels: 64 rows per CTA)
static bool ln_warp_rows(int C) { return C <= 1024; }        // backward (and its partial layout): up to 4 vectors per lane
static bool ln_warp_rows_fwd(int C) { return C <= 2048; }    // forward: up to 8 vectors per lane (C = 2048: 1.5B model)
extern "C" int vrwkv_ln_mix_blocks2(int rows, int C) {
    return ln_warp_rows(C) ? (rows + WR_RUN * WR_WARPS - 1) / (WR_RUN * WR_WARPS) : vrwkv_ln_mix_blocks(rows);
}
static size_t wr_smem(int nmix, int nv, bool bwd) {
    const size_t table = (size_t)(2 + nmix) * nv * 32 * sizeof(uint4), acc = (size_t)2 * nv * 256 * sizeof(float);
    return bwd && acc > table ? acc : table;
}
// NMIX x NV dispatch of the warp-per-row kernels
#define VRWKV_WR_NV(KERNEL, nm, nv, grid, smem, st, args)                                \
    switch (nv) {                                                                        \
        case 1: KERNEL<nm, 1><<<grid, WR_WARPS * 32, smem, st>>>(args); break;           \
        case 2: KERNEL<nm, 2><<<grid, WR_WARPS * 32, smem, st>>>(args); break;           \
        case 3: KERNEL<nm, 3><<<grid, WR_WARPS * 32, smem, st>>>(args); break;           \
        default: KERNEL<nm, 4><<<grid, WR_WARPS * 32, smem, st>>>(args); break;          \
    }
#define VRWKV_WR(KERNEL, nmix, nv, grid, smem, st, args)                                 \
    switch (nmix) {                                                                      \
        case 0: VRWKV_WR_NV(KERNEL, 0, nv, grid, smem, st, args) break;                  \
        case 1: VRWKV_WR_NV(KERNEL, 1, nv, grid, smem, st, args) break;                  \
        default: VRWKV_WR_NV(KERNEL, 6, nv, grid, smem, st, args) break;                 \
    }

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
    if (nmix != 0 && nmix != 1 && nmix != 6) return vrwkv_fail(VRWKV_EUNSUP, "ln_mix_forward: nmix must be 0, 1 or 6");
    if (ln_warp_rows_fwd(C)) {
        int nv = (C + 255) / 256;
        nv = nv <= 4 ? nv : (nv <= 6 ? 6 : 8);
        const int nblk = (rows + WR_RUN * WR_WARPS - 1) / (WR_RUN * WR_WARPS);
        const size_t smem = wr_smem(nmix, nv, false);
        if (nv <= 4) {
            VRWKV_WR(ln_mix_fwd_wr_kernel, nmix, nv, nblk, smem, (cudaStream_t)stream, a)
        } else {
#define VRWKV_WR_BIG(nm)                                                                                        \
    if (nv == 6) ln_mix_fwd_wr_kernel<nm, 6><<<nblk, WR_WARPS * 32, smem, (cudaStream_t)stream>>>(a);           \
    else ln_mix_fwd_wr_kernel<nm, 8><<<nblk, WR_WARPS * 32, smem, (cudaStream_t)stream>>>(a);
            if (nmix == 0) { VRWKV_WR_BIG(0) } else if (nmix == 1) { VRWKV_WR_BIG(1) } else { VRWKV_WR_BIG(6) }
#undef VRWKV_WR_BIG
        }
        VRWKV_CUDA(cudaGetLastError());
        vrwkv_count_launch(1);
        return VRWKV_OK;
    }
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
    if (nmix != 0 && nmix != 1 && nmix != 6) return vrwkv_fail(VRWKV_EUNSUP, "ln_mix_backward: nmix must be 0, 1 or 6");
    if (ln_warp_rows(C)) {
        const int nv = (C + 255) / 256, nblk = vrwkv_ln_mix_blocks2(rows, C);
        VRWKV_WR(ln_mix_bwd_wr_kernel, nmix, nv, nblk, wr_smem(nmix, nv, true), (cudaStream_t)stream, a)
        if (nmix == 1) {
            const int tpg = ((C / 8 + 31) / 32) * 32;   // threads per row group (8 columns each), <= 128
            const size_t sm = (size_t)3 * 8 * tpg * sizeof(float);
            if (4 * tpg <= 384) ln_mix_bwd_params_kernel<1, 4, 8, 384, 2><<<nblk, 4 * tpg, sm, (cudaStream_t)stream>>>(a);
            else ln_mix_bwd_params_kernel<1, 4, 8, 512, 1><<<nblk, 4 * tpg, sm, (cudaStream_t)stream>>>(a);
        } else if (nmix == 6) {
            const int tpg = ((C / 4 + 31) / 32) * 32;   // 4 columns each, <= 256
            const size_t sm = (size_t)8 * 4 * tpg * sizeof(float);
            if (2 * tpg <= 384) ln_mix_bwd_params_kernel<6, 2, 4, 384, 2><<<nblk, 2 * tpg, sm, (cudaStream_t)stream>>>(a);
            else ln_mix_bwd_params_kernel<6, 2, 4, 512, 1><<<nblk, 2 * tpg, sm, (cudaStream_t)stream>>>(a);
        }
        VRWKV_CUDA(cudaGetLastError());
        vrwkv_count_launch(nmix ? 2 : 1);
        return VRWKV_OK;
    }
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
