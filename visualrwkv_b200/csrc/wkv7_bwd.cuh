// wkv7_bwd.cuh — WKV7 backward (reverse-time) for sm_100a.
//
// Same algorithm and same saved-tensor contract as backward_kernel of the reference
// (VisualRWKV-v7/v7.00/cuda/wkv7_cuda.cu:54-130): walk t = T-1..0, reload the transposed fp32
// state checkpoint at chunk ends, un-step the state by dividing by the decay, keep dS in both
// orientations so every reduction is along a thread's own registers.  Machine mapping:
//
//   * one CTA per (head, batch); index i (row of dS / column of S) is spread over L lanes, each
//     lane owning 64/L of the j-range (interleaved float4 groups, conflict-free LDS.128), so the
//     3 x 64 fp32 registers per i of the reference become 3 x 64/L per thread and the CTA has
//     64*L compute threads instead of 64;
//   * the seven bf16 streams (w,q,k,v,a,b,dy) and the fp32 `sa` stream are staged by TMA
//     ([16 x 64] tiles, mbarrier completion) in reverse chunk order; converter warps expand them
//     to fp32 once per head (decay and its derivative factor evaluated once per (t,j));
//   * the seven per-step dot products are reduced with warp shuffles; the only cross-warp
//     exchange (dSb) goes through a double-buffered 64-float shared array guarded by a
//     split-phase mbarrier (arrive right after the critical reduction, wait just before use);
//   * state checkpoints are prefetched into registers one chunk ahead straight from L2/HBM;
//   * elementwise math is packed fp32x2 (FFMA2).
//
// Algorithmic HBM bytes: 26 B per (b,t,c) element (7 bf16 reads + 6 bf16 writes); the reference
// contract adds 4 B (sa) + 16 B (s) of reads per element.
#pragma once
#include "common.cuh"
#include "wkv7_fwd.cuh"

namespace vrwkv {

struct Wkv7BwdArgs {
    int B, T, H;
    const float* s;
    uint16_t *dw, *dq, *dk, *dv, *da, *db;
};

template <int NSTAGE>
struct alignas(128) Wkv7BwdSmem {
    uint16_t raw[NSTAGE][7][WKV_TC][WKV_N];  // w,q,k,v,a,b,dy
    float raw_sa[NSTAGE][WKV_TC][WKV_N];
    float f[2][9][WKV_TC][WKV_N];            // decay,q,k,v,a,b,dy,sa,wfac
    float dsb[2][WKV_N];
    uint64_t full_raw[NSTAGE], empty_raw[NSTAGE], full_f[2], empty_f[2], dsb_bar[2];
};

template <int L, int NCONV, int NSTAGE>
__global__ void __launch_bounds__(WKV_N* L + NCONV * 32)
wkv7_bwd_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_q,
                const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                const __grid_constant__ CUtensorMap tm_dy, const __grid_constant__ CUtensorMap tm_sa,
                const Wkv7BwdArgs p) {
    constexpr int N = WKV_N, TC = WKV_TC;
    constexpr int COLS = N / L, M = COLS / 4;
    constexpr int NCOMP = N * L, NCW = NCOMP / 32;
    constexpr int TS = TC * N;

    extern __shared__ __align__(128) uint8_t smem_bytes[];
    Wkv7BwdSmem<NSTAGE>& sm = *reinterpret_cast<Wkv7BwdSmem<NSTAGE>*>(smem_bytes);

    const int hh = blockIdx.x, bb = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for ptxas (see wkv7_fwd.cuh)
    const int T = p.T, H = p.H;
    const int nchunks = T / TC;

    if (tid == 0) {
        for (int i = 0; i < NSTAGE; i++) {
            mbar_init(&sm.full_raw[i], 1);
            mbar_init(&sm.empty_raw[i], NCONV);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&sm.full_f[i], NCONV);
            mbar_init(&sm.empty_f[i], NCW);
            mbar_init(&sm.dsb_bar[i], NCW);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp >= NCW) {
        // =============================== converter / producer warps ===========================
        const int cw = warp - NCW;
        const bool producer = (cw == 0 && lane == 0);
        auto issue = [&](int n) {  // n-th chunk in processing order = chunk nchunks-1-n
            const int stage = n % NSTAGE;
            uint64_t* bar = &sm.full_raw[stage];
            mbar_arrive_expect_tx(bar, 7 * TC * N * 2 + TC * N * 4);
            const int x0 = hh * N, y0 = bb * T + (nchunks - 1 - n) * TC;
            tma_load_2d(&sm.raw[stage][0][0][0], &tm_w, x0, y0, bar);
            tma_load_2d(&sm.raw[stage][1][0][0], &tm_q, x0, y0, bar);
            tma_load_2d(&sm.raw[stage][2][0][0], &tm_k, x0, y0, bar);
            tma_load_2d(&sm.raw[stage][3][0][0], &tm_v, x0, y0, bar);
            tma_load_2d(&sm.raw[stage][4][0][0], &tm_a, x0, y0, bar);
            tma_load_2d(&sm.raw[stage][5][0][0], &tm_b, x0, y0, bar);
            tma_load_2d(&sm.raw[stage][6][0][0], &tm_dy, x0, y0, bar);
            tma_load_2d(&sm.raw_sa[stage][0][0], &tm_sa, x0, y0, bar);
        };
        if (producer) {
            tma_prefetch_desc(&tm_w); tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k);
            tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_a); tma_prefetch_desc(&tm_b);
            tma_prefetch_desc(&tm_dy); tma_prefetch_desc(&tm_sa);
            for (int n = 0; n < NSTAGE && n < nchunks; n++) issue(n);
        }
        for (int n = 0; n < nchunks; n++) {
            const int stage = n % NSTAGE;
            const uint32_t rpar = (n / NSTAGE) & 1;
            const int buf = n & 1;
            mbar_wait(&sm.full_raw[stage], rpar);
            if (n >= 2) mbar_wait(&sm.empty_f[buf], ((n >> 1) - 1) & 1);
            // 7 bf16 tensors x 8 row-pairs + fp32 sa x 8 row-pairs
#pragma unroll 4
            for (int g = cw; g < 64; g += NCONV) {
                const int tensor = g >> 3, off = (g & 7) * 128 + lane * 4;
                if (tensor < 7) {
                    const uint2 u = *reinterpret_cast<const uint2*>(&sm.raw[stage][tensor][0][0] + off);
                    float4 o;
                    o.x = bf16lo_to_f32(u.x); o.y = bf16hi_to_f32(u.x);
                    o.z = bf16lo_to_f32(u.y); o.w = bf16hi_to_f32(u.y);
                    if (tensor == 0) {  // wfac = -exp(w), decay = exp(wfac)  (wkv7_cuda.cu:67-68)
                        float4 wf;
                        wf.x = -__expf(o.x); wf.y = -__expf(o.y); wf.z = -__expf(o.z); wf.w = -__expf(o.w);
                        *reinterpret_cast<float4*>(&sm.f[buf][8][0][0] + off) = wf;
                        o.x = __expf(wf.x); o.y = __expf(wf.y); o.z = __expf(wf.z); o.w = __expf(wf.w);
                    }
                    *reinterpret_cast<float4*>(&sm.f[buf][tensor][0][0] + off) = o;
                } else {
                    *reinterpret_cast<float4*>(&sm.f[buf][7][0][0] + off) =
                        *reinterpret_cast<const float4*>(&sm.raw_sa[stage][0][0] + off);
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&sm.full_f[buf]);
                mbar_arrive(&sm.empty_raw[stage]);
            }
            if (producer && n + NSTAGE < nchunks) {
                mbar_wait(&sm.empty_raw[stage], rpar);
                issue(n + NSTAGE);
            }
        }
        return;
    }

    // ===================================== compute warps ======================================
    const int l = tid % L, i = tid / L;
    u64 ST[M][2], dS[M][2], dST[M][2];
#pragma unroll
    for (int m = 0; m < M; m++) ST[m][0] = ST[m][1] = dS[m][0] = dS[m][1] = dST[m][0] = dST[m][1] = 0ull;

    const float* sbase = p.s + ((size_t)bb * H + hh) * nchunks * N * N + (size_t)i * N;
    float4 pf[M];
#pragma unroll
    for (int m = 0; m < M; m++)
        pf[m] = __ldg(reinterpret_cast<const float4*>(sbase + (size_t)(nchunks - 1) * N * N + 4 * (l + L * m)));

    uint16_t* const pA = (l == 0 ? p.dq : p.dv) + i;
    uint16_t* const pB = (l == 0 ? p.dw : p.db) + i;
    auto reduce = [&](float x) {
#pragma unroll
        for (int o = 1; o < L; o <<= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        return x;
    };

    auto step = [&](const int buf, const int t, const size_t ind) {
        const float* f0 = &sm.f[buf][0][t][0];
        const int par = (t + 1) & 1;                     // dsb buffer of this step
        const uint32_t dpar = ((TC - 1 - t) >> 1) & 1;   // its mbarrier phase parity
        const float qi = f0[1 * TS + i], wi = f0[0 * TS + i], ki = f0[2 * TS + i], ai = f0[4 * TS + i],
                    bi = f0[5 * TS + i], dyi = f0[6 * TS + i], wfi = f0[8 * TS + i];
        // ---- dS += dy_i q_j ; dSb_i = sum_j dS_ij b_j   (critical path, wkv7_cuda.cu:96,105) ----
        float dSb;
        {
            const u64 dyi2 = pk2(dyi, dyi);
            u64 acc0 = 0ull, acc1 = 0ull;
#pragma unroll
            for (int m = 0; m < M; m++) {
                const int co = 4 * (l + L * m);
                const float4 q4 = *reinterpret_cast<const float4*>(f0 + 1 * TS + co);
                const float4 b4 = *reinterpret_cast<const float4*>(f0 + 5 * TS + co);
                dS[m][0] = ffma2(dyi2, pk2(q4.x, q4.y), dS[m][0]);
                dS[m][1] = ffma2(dyi2, pk2(q4.z, q4.w), dS[m][1]);
                acc0 = ffma2(dS[m][0], pk2(b4.x, b4.y), acc0);
                acc1 = ffma2(dS[m][1], pk2(b4.z, b4.w), acc1);
            }
            dSb = reduce(hsum2(fadd2(acc0, acc1)));
            if (l == 0) sm.dsb[par][i] = dSb;
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.dsb_bar[par]);
        }
        // ---- dq, un-step, dST += q_i dy_j, dw/dk/db/dv  (wkv7_cuda.cu:84-111) ----
        float dq, dw, dk, db, dv;
        {
            const float iwi = __frcp_rn(wi);
            const u64 nki2 = pk2(-ki, -ki), nbi2 = pk2(-bi, -bi), iwi2 = pk2(iwi, iwi), qi2 = pk2(qi, qi);
            u64 aq = 0ull, aw = 0ull, ak = 0ull, ab = 0ull, av = 0ull;
#pragma unroll
            for (int m = 0; m < M; m++) {
                const int co = 4 * (l + L * m);
                const float4 dy4 = *reinterpret_cast<const float4*>(f0 + 6 * TS + co);
                const float4 v4 = *reinterpret_cast<const float4*>(f0 + 3 * TS + co);
                const float4 sa4 = *reinterpret_cast<const float4*>(f0 + 7 * TS + co);
                const float4 k4 = *reinterpret_cast<const float4*>(f0 + 2 * TS + co);
                const u64 dy2[2] = {pk2(dy4.x, dy4.y), pk2(dy4.z, dy4.w)};
                const u64 v2[2] = {pk2(v4.x, v4.y), pk2(v4.z, v4.w)};
                const u64 sa2[2] = {pk2(sa4.x, sa4.y), pk2(sa4.z, sa4.w)};
                const u64 k2[2] = {pk2(k4.x, k4.y), pk2(k4.z, k4.w)};
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    aq = ffma2(ST[m][h], dy2[h], aq);
                    u64 st = ffma2(nki2, v2[h], ST[m][h]);
                    st = ffma2(nbi2, sa2[h], st);
                    st = fmul2(st, iwi2);
                    ST[m][h] = st;
                    const u64 dst = ffma2(qi2, dy2[h], dST[m][h]);
                    dST[m][h] = dst;
                    aw = ffma2(dst, st, aw);
                    ak = ffma2(dst, v2[h], ak);
                    ab = ffma2(dst, sa2[h], ab);
                    av = ffma2(dS[m][h], k2[h], av);
                }
            }
            dq = reduce(hsum2(aq));
            dw = reduce(hsum2(aw));
            dk = reduce(hsum2(ak));
            db = reduce(hsum2(ab));
            dv = reduce(hsum2(av));
        }
        // branch-free stores: lane l==0 writes dq/dw/dk, lane l==1 writes dv/db/da (all lanes hold the sums)
        {
            const float dwv = dw * wi * wfi;
            st_pred_b16(pA + ind, f32_to_bf16_bits(l == 0 ? dq : dv), l < 2);
            st_pred_b16(pB + ind, f32_to_bf16_bits(l == 0 ? dwv : db), l < 2);
            st_pred_b16(p.dk + ind + i, f32_to_bf16_bits(dk), l == 0);
        }
        // ---- da_i = sum_j S_ji dSb_j ; propagate dS  (wkv7_cuda.cu:117-128) ----
        mbar_wait(&sm.dsb_bar[par], dpar);
        {
            const u64 dSb2 = pk2(dSb, dSb), wi2 = pk2(wi, wi), ai2 = pk2(ai, ai);
            u64 aa = 0ull;
#pragma unroll
            for (int m = 0; m < M; m++) {
                const int co = 4 * (l + L * m);
                const float4 d4 = *reinterpret_cast<const float4*>(&sm.dsb[par][co]);
                const float4 w4 = *reinterpret_cast<const float4*>(f0 + 0 * TS + co);
                const float4 a4 = *reinterpret_cast<const float4*>(f0 + 4 * TS + co);
                const u64 d2[2] = {pk2(d4.x, d4.y), pk2(d4.z, d4.w)};
                const u64 w2[2] = {pk2(w4.x, w4.y), pk2(w4.z, w4.w)};
                const u64 a2[2] = {pk2(a4.x, a4.y), pk2(a4.z, a4.w)};
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    aa = ffma2(ST[m][h], d2[h], aa);
                    dS[m][h] = ffma2(dS[m][h], w2[h], fmul2(dSb2, a2[h]));
                    dST[m][h] = ffma2(dST[m][h], wi2, fmul2(ai2, d2[h]));
                }
            }
            const float da = reduce(hsum2(aa));
            st_pred_b16(p.da + ind + i, f32_to_bf16_bits(da), l == 1);
        }
    };

    for (int n = 0; n < nchunks; n++) {
        const int c = nchunks - 1 - n, buf = n & 1;
#pragma unroll
        for (int m = 0; m < M; m++) {
            ST[m][0] = pk2(pf[m].x, pf[m].y);
            ST[m][1] = pk2(pf[m].z, pf[m].w);
        }
        if (c > 0) {
#pragma unroll
            for (int m = 0; m < M; m++)
                pf[m] = __ldg(reinterpret_cast<const float4*>(sbase + (size_t)(c - 1) * N * N + 4 * (l + L * m)));
        }
        mbar_wait(&sm.full_f[buf], (n >> 1) & 1);
        __syncwarp();
        const size_t ind0 = (((size_t)bb * T + (size_t)c * TC) * H + hh) * N;
#pragma unroll 2
        for (int t = TC - 1; t >= 0; t--) step(buf, t, ind0 + (size_t)t * H * N);
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty_f[buf]);
    }
}

}  // namespace vrwkv
