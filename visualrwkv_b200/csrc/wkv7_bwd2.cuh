// wkv7_bwd2.cuh — WKV7 backward, second-generation mapping.
//
// Same algorithm / saved-tensor contract as the reference backward_kernel (VisualRWKV-v7/v7.00/cuda/wkv7_cuda.cu:
// 54-130): reverse time, reload the transposed fp32 checkpoint at chunk ends, un-step the state by dividing by the
// decay, keep dS in both orientations.  Design notes (see wkv7_fwd2.cuh for what ncu showed):
//
//   * a thread owns R i's x 8 j's of each of the three fp32 arrays (S^T, dS, dS^T); a head is (64/R) i-groups x 8
//     j-groups: R=2 -> 256 threads = 8 warps (two per SM sub-partition: one warp's shuffle / dependency stalls are
//     filled by the other), R=4 -> 128 threads;
//   * q,k,v,a,b,dy are read as bf16 straight from the TMA tiles (one LDS.128 per 8 columns) and expanded on the ALU
//     pipe; decay / decay-derivative factor / sa live in fp32 tiles written once per (t, column) by a converter
//     warp, column-permuted so that every LDS.128 is bank-conflict free;
//   * step order: [dS += dy q ; dSb = dS.b and dv = dS.k reduced together ; dSb posted to smem + mbarrier arrive ;
//     dS update] then the 128 packed ops that do not need the exchange (dq, un-step, dS^T +=, dw, dk, db — their
//     four reductions interleaved), only then the mbarrier wait + [da, dS^T update]: the exchange latency hides
//     behind ~260 FMA-pipe cycles;
//   * all reductions are transposing reduce-scatters with stage-interleaved shuffles (RowLanes<R>);
//   * each gradient row is stored by a distinct lane (branch-free, full sectors per warp).
//
// Per step and head (one SM): FMA pipe >= 512 cycles (16 FMA x 4096 / 128 lanes).  Algorithmic HBM bytes: 26 B per
// (b,t,c) element; the reference contract adds 4 B (sa) + 16 B (s) of reads.
#pragma once
#include "common.cuh"
#include "wkv7_fwd2.cuh"

namespace vrwkv {

template <int NSTAGE>
struct alignas(128) Wkv7Bwd2Smem {
    uint16_t raw[NSTAGE][7][WKV_TC][WKV_N];  // w,q,k,v,a,b,dy (bf16, TMA)
    float raw_sa[NSTAGE][WKV_TC][WKV_N];     // sa (fp32, TMA)
    float decay[2][WKV_TC][WKV_N];           // exp(-exp(w)), permuted columns
    float wfac[2][WKV_TC][WKV_N];            // -exp(w), permuted columns
    float sa[2][WKV_TC][WKV_N];              // sa, permuted columns
    float dsb[2][WKV_N];                     // exchanged dSb, permuted
    uint64_t full_raw[NSTAGE], empty_raw[NSTAGE], full_f[2], empty_f[2], dsb_bar[2];
};

// position of column j in the permuted fp32 tiles: a thread's columns 8l..8l+3 / 8l+4..8l+7 (and an i-group's
// consecutive indices) are 16-byte chunks that are contiguous across lanes.
__device__ __forceinline__ int perm_col(int j) { return ((j >> 2) & 1) * 32 + 4 * (j >> 3) + (j & 3); }

// Producer + fp32 converter warp shared by the backward kernels: issues the TMA tile loads in reverse chunk order and
// expands w into decay / decay-derivative factor and copies sa into the column-permuted fp32 tiles.
template <int NSTAGE, class Smem>
__device__ __forceinline__ void wkv7_bwd_producer_converter(Smem& sm, const CUtensorMap& tm_w, const CUtensorMap& tm_q,
                                                            const CUtensorMap& tm_k, const CUtensorMap& tm_v,
                                                            const CUtensorMap& tm_a, const CUtensorMap& tm_b,
                                                            const CUtensorMap& tm_dy, const CUtensorMap& tm_sa, const int hh,
                                                            const int bb, const int T, const int c_top, const int nchunks, const int lane) {
    // processes chunks c_top-1, c_top-2, ... (nchunks of them)
    constexpr int N = WKV_N, TC = WKV_TC;
        // ================= producer + fp32 converter warp =================
        auto issue = [&](int n) {  // n-th chunk in processing order = chunk c_top-1-n
            const int stage = n % NSTAGE;
            uint64_t* bar = &sm.full_raw[stage];
            mbar_arrive_expect_tx(bar, 7 * TC * N * 2 + TC * N * 4);
            const int x0 = hh * N, y0 = bb * T + (c_top - 1 - n) * TC;
            tma_load_2d(&sm.raw[stage][0][0][0], &tm_w, x0, y0, bar);
            tma_load_2d(&sm.raw[stage][1][0][0], &tm_q, x0, y0, bar);
            tma_load_2d(&sm.raw[stage][2][0][0], &tm_k, x0, y0, bar);
            tma_load_2d(&sm.raw[stage][3][0][0], &tm_v, x0, y0, bar);
            tma_load_2d(&sm.raw[stage][4][0][0], &tm_a, x0, y0, bar);
            tma_load_2d(&sm.raw[stage][5][0][0], &tm_b, x0, y0, bar);
            tma_load_2d(&sm.raw[stage][6][0][0], &tm_dy, x0, y0, bar);
            tma_load_2d(&sm.raw_sa[stage][0][0], &tm_sa, x0, y0, bar);
        };
        if (lane == 0) {
            tma_prefetch_desc(&tm_w); tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k);
            tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_a); tma_prefetch_desc(&tm_b);
            tma_prefetch_desc(&tm_dy); tma_prefetch_desc(&tm_sa);
            for (int n = 0; n < NSTAGE && n < nchunks; n++) issue(n);
        }
        for (int n = 0; n < nchunks; n++) {
            const int stage = n % NSTAGE, buf = n & 1;
            mbar_wait(&sm.full_raw[stage], (n / NSTAGE) & 1);
            if (n >= 2) mbar_wait(&sm.empty_f[buf], ((n >> 1) - 1) & 1);
            __syncwarp();
#pragma unroll
            for (int g = 0; g < 8; g++) {
                const int row = 2 * g + (lane >> 4), m = lane & 15;  // m: group of 4 columns 4m..4m+3
                const int pos = (m & 1) * 32 + 4 * (m >> 1);
                const uint2 u = *reinterpret_cast<const uint2*>(&sm.raw[stage][0][row][4 * m]);
                float4 wf, o;
                wf.x = -__expf(__uint_as_float(u.x << 16));          // wkv7_cuda.cu:67
                wf.y = -__expf(__uint_as_float(u.x & 0xffff0000u));
                wf.z = -__expf(__uint_as_float(u.y << 16));
                wf.w = -__expf(__uint_as_float(u.y & 0xffff0000u));
                o.x = __expf(wf.x); o.y = __expf(wf.y); o.z = __expf(wf.z); o.w = __expf(wf.w);  // :68
                *reinterpret_cast<float4*>(&sm.wfac[buf][row][pos]) = wf;
                *reinterpret_cast<float4*>(&sm.decay[buf][row][pos]) = o;
                *reinterpret_cast<float4*>(&sm.sa[buf][row][pos]) =
                    *reinterpret_cast<const float4*>(&sm.raw_sa[stage][row][4 * m]);
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&sm.full_f[buf]);
                mbar_arrive(&sm.empty_raw[stage]);
                if (n >= 1 && n - 1 + NSTAGE < nchunks) {
                    mbar_wait(&sm.empty_raw[(n - 1) % NSTAGE], ((n - 1) / NSTAGE) & 1);
                    issue(n - 1 + NSTAGE);
                }
            }
            __syncwarp();
        }
}

template <int R, int NSTAGE, int UNROLL = 2>
__global__ void __launch_bounds__((WKV_N / R) * 8 + 32)  // (capping R=4 at 192 registers for two CTAs per SM spills: slower)
wkv7_bwd2_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_q,
                 const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                 const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const __grid_constant__ CUtensorMap tm_dy, const __grid_constant__ CUtensorMap tm_sa,
                 const Wkv7BwdArgs p) {
    constexpr int N = WKV_N, TC = WKV_TC;
    constexpr int NCW = (N / R) * 8 / 32;
    constexpr int TS = TC * N;
    extern __shared__ __align__(128) uint8_t smem_bytes[];
    Wkv7Bwd2Smem<NSTAGE>& sm = *reinterpret_cast<Wkv7Bwd2Smem<NSTAGE>*>(smem_bytes);

    const int hh = blockIdx.x, bb = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int T = p.T, H = p.H;
    const int nchunks_all = T / TC;
    const int span = p.span > 0 ? p.span : nchunks_all;
    const int seg = blockIdx.z;
    const int c_lo = seg * span, c_top = min(nchunks_all, c_lo + span);
    const int nchunks = c_top - c_lo;  // 16-step chunks this CTA walks (from c_top-1 down to c_lo)

    if (tid == 0) {
        for (int i = 0; i < NSTAGE; i++) {
            mbar_init(&sm.full_raw[i], 1);
            mbar_init(&sm.empty_raw[i], NCW + 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&sm.full_f[i], 1);
            mbar_init(&sm.empty_f[i], NCW);
            mbar_init(&sm.dsb_bar[i], NCW);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == NCW) {
        wkv7_bwd_producer_converter<NSTAGE>(sm, tm_w, tm_q, tm_k, tm_v, tm_a, tm_b, tm_dy, tm_sa, hh, bb, T, c_top, nchunks, lane);
        return;
    }

    // ===================================== compute warps ======================================
    const int l = tid & 7;          // j block: 8l .. 8l+7
    const int i0 = (tid >> 3) * R;  // i block: i0 .. i0+R-1
    const RowLanes<R> L(l);
    const int myi = i0 + L.own_row;  // the i whose sums this lane owns after a reduce-scatter
    const int ipos = perm_col(i0);   // permuted position of i0 (R consecutive floats)
    const int mypos = perm_col(myi);

    u64 ST[R][4], dS[R][4], dST[R][4];
#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) ST[r][c] = dS[r][c] = dST[r][c] = 0ull;

    // checkpoint memory [row = i][col = j] holds S_{j,i} (stored transposed by the forward, wkv7_cuda.cu:44-50)
    const float* sbase = p.s + ((size_t)bb * H + hh) * nchunks_all * N * N + (size_t)i0 * N + 8 * l;
    float4 pf[R][2];
    auto prefetch = [&](int c) {
#pragma unroll
        for (int r = 0; r < R; r++) {
            pf[r][0] = __ldg(reinterpret_cast<const float4*>(sbase + (size_t)c * N * N + r * N));
            pf[r][1] = __ldg(reinterpret_cast<const float4*>(sbase + (size_t)c * N * N + r * N + 4));
        }
    };
    prefetch(c_top - 1);
    if (p.ds_in && c_top < nchunks_all) {
        // dL/dS at the end of this segment, in both orientations (dS_ij and dS_ji)
        const float* src = p.ds_in + (((size_t)bb * H + hh) * gridDim.z + seg) * N * N;
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int c2 = 0; c2 < 4; c2++) {
                const int j = 8 * l + 2 * c2;
                dS[r][c2] = pk2(__ldg(src + (i0 + r) * N + j), __ldg(src + (i0 + r) * N + j + 1));
                dST[r][c2] = pk2(__ldg(src + (size_t)j * N + i0 + r), __ldg(src + (size_t)(j + 1) * N + i0 + r));
            }
    }

    // which gradient rows this lane stores (see header): E after phase B1, F (R==4 only) with it, G after phase B2
    uint16_t *pE, *pF = nullptr, *pG;
    bool predF = false, predG = true;
    if constexpr (R == 4) {
        pE = (L.sub ? p.dw : p.dq) + myi;
        pF = (L.sub ? p.dv : p.dk) + myi;
        pG = (L.sub ? p.da : p.db) + myi;
        predF = true;
    } else {
        pE = (L.sub == 0 ? p.dq : L.sub == 1 ? p.dw : L.sub == 2 ? p.dk : p.db) + myi;
        pG = (L.sub == 0 ? p.da : p.dv) + myi;
        predG = L.sub < 2;
    }

    auto unpackR = [&](const uint16_t* ptr, float (&o)[R]) {
        if constexpr (R == 4) {
            const uint2 u = *reinterpret_cast<const uint2*>(ptr);
            o[0] = __uint_as_float(u.x << 16); o[1] = __uint_as_float(u.x & 0xffff0000u);
            o[2] = __uint_as_float(u.y << 16); o[3] = __uint_as_float(u.y & 0xffff0000u);
        } else {
            const uint32_t u = *reinterpret_cast<const uint32_t*>(ptr);
            o[0] = __uint_as_float(u << 16); o[1] = __uint_as_float(u & 0xffff0000u);
        }
    };

    auto step = [&](const int stage, const int buf, const int t, const size_t ind) {
        const uint16_t* r0 = &sm.raw[stage][0][t][0];
        const int par = (t + 1) & 1;                    // exchange buffer of this step
        const uint32_t dpar = ((TC - 1 - t) >> 1) & 1;  // its mbarrier phase parity
        // ---------------- phase A: dS chain ----------------
        float dvv;
        {
            u64 q2[4], b2v[4], k2[4], a2[4];
            unpack8(*reinterpret_cast<const uint4*>(r0 + 1 * TS + 8 * l), q2);
            unpack8(*reinterpret_cast<const uint4*>(r0 + 5 * TS + 8 * l), b2v);
            unpack8(*reinterpret_cast<const uint4*>(r0 + 2 * TS + 8 * l), k2);
            unpack8(*reinterpret_cast<const uint4*>(r0 + 4 * TS + 8 * l), a2);
            float dyi[R];
            unpackR(r0 + 6 * TS + i0, dyi);
            const float4 w0 = *reinterpret_cast<const float4*>(&sm.decay[buf][t][4 * l]);
            const float4 w1 = *reinterpret_cast<const float4*>(&sm.decay[buf][t][32 + 4 * l]);
            const u64 w2[4] = {pk2(w0.x, w0.y), pk2(w0.z, w0.w), pk2(w1.x, w1.y), pk2(w1.z, w1.w)};
            float x[2][R], z[2], dSb[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
                const u64 d2 = pk2(dyi[r], dyi[r]);
                u64 ab = 0ull, av = 0ull;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    dS[r][c] = ffma2(d2, q2[c], dS[r][c]);  // dS_ij += dy_i q_j   (wkv7_cuda.cu:96)
                    ab = ffma2(dS[r][c], b2v[c], ab);       // dSb_i               (:105)
                    av = ffma2(dS[r][c], k2[c], av);        // dv_i                (:104)
                }
                x[0][r] = hsum2(ab);
                x[1][r] = hsum2(av);
            }
            L.template reduce<2>(x, z);
            dvv = z[1];
            sm.dsb[par][mypos] = z[0];  // the NSUB lanes of a row hold (and write) the same value
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.dsb_bar[par]);
            L.allgather(z[0], dSb);
#pragma unroll
            for (int r = 0; r < R; r++) {
                const u64 s2 = pk2(dSb[r], dSb[r]);
#pragma unroll
                for (int c = 0; c < 4; c++) dS[r][c] = ffma2(dS[r][c], w2[c], fmul2(s2, a2[c]));  // (:125)
            }
        }
        // ---------------- phase B1: everything that does not need the exchanged dSb ----------------
        float dqv, dwv, dkv, dbv;
        float wi[R], ai[R];
        {
            u64 dy2[4], v2[4];
            unpack8(*reinterpret_cast<const uint4*>(r0 + 6 * TS + 8 * l), dy2);
            unpack8(*reinterpret_cast<const uint4*>(r0 + 3 * TS + 8 * l), v2);
            const float4 s0 = *reinterpret_cast<const float4*>(&sm.sa[buf][t][4 * l]);
            const float4 s1 = *reinterpret_cast<const float4*>(&sm.sa[buf][t][32 + 4 * l]);
            const u64 sa2[4] = {pk2(s0.x, s0.y), pk2(s0.z, s0.w), pk2(s1.x, s1.y), pk2(s1.z, s1.w)};
            float ki[R], bi[R], qi[R];
            unpackR(r0 + 2 * TS + i0, ki);
            unpackR(r0 + 5 * TS + i0, bi);
            unpackR(r0 + 1 * TS + i0, qi);
            unpackR(r0 + 4 * TS + i0, ai);
            if constexpr (R == 4) {
                const float4 w4 = *reinterpret_cast<const float4*>(&sm.decay[buf][t][ipos]);
                wi[0] = w4.x; wi[1] = w4.y; wi[2] = w4.z; wi[3] = w4.w;
            } else {
                const float2 w4 = *reinterpret_cast<const float2*>(&sm.decay[buf][t][ipos]);
                wi[0] = w4.x; wi[1] = w4.y;
            }
            float x[4][R], z[4];
#pragma unroll
            for (int r = 0; r < R; r++) {
                const float iw = __fdividef(1.f, wi[r]);  // the reference divides under --use_fast_math (:93)
                const u64 nk2 = pk2(-ki[r], -ki[r]), nb2 = pk2(-bi[r], -bi[r]), iw2 = pk2(iw, iw), q2i = pk2(qi[r], qi[r]);
                u64 aq = 0ull, aw = 0ull, ak = 0ull, ab = 0ull;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    aq = ffma2(ST[r][c], dy2[c], aq);              // dq_i = sum_j S_ji dy_j   (:84-89)
                    u64 st = ffma2(nk2, v2[c], ST[r][c]);          // un-step                 (:91-95)
                    st = ffma2(nb2, sa2[c], st);
                    st = fmul2(st, iw2);
                    ST[r][c] = st;
                    const u64 dt = ffma2(q2i, dy2[c], dST[r][c]);  // dS_ji += q_i dy_j        (:97)
                    dST[r][c] = dt;
                    aw = ffma2(dt, st, aw);                        // dw                      (:102)
                    ak = ffma2(dt, v2[c], ak);                     // dk                      (:103)
                    ab = ffma2(dt, sa2[c], ab);                    // db                      (:106)
                }
                x[0][r] = hsum2(aq); x[1][r] = hsum2(aw); x[2][r] = hsum2(ak); x[3][r] = hsum2(ab);
            }
            L.template reduce<4>(x, z);
            dqv = z[0]; dkv = z[2]; dbv = z[3];
            dwv = z[1] * (sm.decay[buf][t][mypos] * sm.wfac[buf][t][mypos]);  // (:108)
        }
        if constexpr (R == 4) {
            st_pred_b16(pE + ind, f32_to_bf16_bits(L.sub ? dwv : dqv), true);
            st_pred_b16(pF + ind, f32_to_bf16_bits(L.sub ? dvv : dkv), predF);
        } else {
            const float vE = L.sub == 0 ? dqv : L.sub == 1 ? dwv : L.sub == 2 ? dkv : dbv;
            st_pred_b16(pE + ind, f32_to_bf16_bits(vE), true);
        }
        // ---------------- phase B2: needs dSb of every i ----------------
        mbar_wait(&sm.dsb_bar[par], dpar);
        float dav;
        {
            const float4 d0 = *reinterpret_cast<const float4*>(&sm.dsb[par][4 * l]);
            const float4 d1 = *reinterpret_cast<const float4*>(&sm.dsb[par][32 + 4 * l]);
            const u64 d2[4] = {pk2(d0.x, d0.y), pk2(d0.z, d0.w), pk2(d1.x, d1.y), pk2(d1.z, d1.w)};
            float x[1][R], z[1];
#pragma unroll
            for (int r = 0; r < R; r++) {
                const u64 w2i = pk2(wi[r], wi[r]), a2i = pk2(ai[r], ai[r]);
                u64 aa = 0ull;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    aa = ffma2(ST[r][c], d2[c], aa);                           // da_i  (:117-122)
                    dST[r][c] = ffma2(dST[r][c], w2i, fmul2(a2i, d2[c]));      // (:126)
                }
                x[0][r] = hsum2(aa);
            }
            L.template reduce<1>(x, z);
            dav = z[0];
        }
        if constexpr (R == 4) {
            st_pred_b16(pG + ind, f32_to_bf16_bits(L.sub ? dav : dbv), true);
        } else {
            st_pred_b16(pG + ind, f32_to_bf16_bits(L.sub == 0 ? dav : dvv), predG);
        }
    };

    for (int n = 0; n < nchunks; n++) {
        const int c = c_top - 1 - n, buf = n & 1, stage = n % NSTAGE;
#pragma unroll
        for (int r = 0; r < R; r++) {
            ST[r][0] = pk2(pf[r][0].x, pf[r][0].y); ST[r][1] = pk2(pf[r][0].z, pf[r][0].w);
            ST[r][2] = pk2(pf[r][1].x, pf[r][1].y); ST[r][3] = pk2(pf[r][1].z, pf[r][1].w);
        }
        if (c > c_lo) prefetch(c - 1);
        mbar_wait(&sm.full_raw[stage], (n / NSTAGE) & 1);
        mbar_wait(&sm.full_f[buf], (n >> 1) & 1);
        __syncwarp();
        const size_t ind0 = (((size_t)bb * T + (size_t)c * TC) * H + hh) * N;
#pragma unroll UNROLL
        for (int t = TC - 1; t >= 0; t--) step(stage, buf, t, ind0 + (size_t)t * H * N);
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&sm.empty_f[buf]);
            mbar_arrive(&sm.empty_raw[stage]);
        }
    }
}

}  // namespace vrwkv
