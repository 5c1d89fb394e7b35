// wkv7_fwd.cuh — WKV7 forward recurrence for sm_100a.
//
// Replaces forward_kernel of the reference (VisualRWKV-v7/v7.00/cuda/wkv7_cuda.cu:10-52) with the
// same contract (y bf16, sa f32 per step, transposed f32 state checkpoint every 16 steps) but a
// different machine mapping:
//
//   * one CTA per (head, batch); the 64x64 fp32 state lives in registers, spread over
//     (64/R)*L compute threads: a thread owns R rows x (64/L) columns, columns interleaved in
//     groups of 4 so that every LDS.128 of a per-column vector is bank-conflict free;
//   * the six bf16 input streams are staged by TMA (cp.async.bulk.tensor.2d) as [16 x 64] tiles of
//     the (B*T, H*64) matrix into an NSTAGE ring, completion on mbarriers;
//   * NCONV converter warps (which run on otherwise idle issue slots) turn each raw tile into fp32
//     once per head (decay exp(-exp(w)) evaluated once per (t,j), not once per row) into a
//     double-buffered fp32 tile that the compute warps read with broadcast LDS.128;
//   * the row dot-products (sa = S.a, y = S.q) are reduced across the L lanes of a row with warp
//     shuffles; all elementwise state math is packed fp32x2 (FFMA2);
//   * y / sa / checkpoints are stored straight from registers (full 32 B sectors per warp).
//
// Algorithmic HBM bytes: 14 B per (b,t,c) element (6 bf16 reads + 1 bf16 write); the reference
// contract adds 4 B (sa) + 16 B (s) per element.
#pragma once
#include "common.cuh"

namespace vrwkv {

constexpr int WKV_N = 64;   // head size (model.py:69)
constexpr int WKV_TC = 16;  // steps per chunk == reference checkpoint interval (_CHUNK_LEN_)

struct Wkv7FwdArgs {
    int B, T, H;
    uint16_t* y;
    float* s;                // may be null
    float* sa;               // may be null
    const float* state_in;   // may be null (zeros)
    float* state_out;        // may be null
};

template <int NSTAGE>
struct alignas(128) Wkv7FwdSmem {
    uint16_t raw[NSTAGE][6][WKV_TC][WKV_N];  // TMA destination ring (w,q,k,v,a,b)
    float f[2][6][WKV_TC][WKV_N];            // converted fp32 tiles (w=decay,q,k,v,a,b)
    uint64_t full_raw[NSTAGE], empty_raw[NSTAGE], full_f[2], empty_f[2];
};

template <int L, int R, int NCONV, int NSTAGE>
__global__ void __launch_bounds__((WKV_N / R) * L + NCONV * 32)
wkv7_fwd_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_q,
                const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                const Wkv7FwdArgs p) {
    constexpr int N = WKV_N, TC = WKV_TC;
    constexpr int COLS = N / L;        // columns per thread
    constexpr int M = COLS / 4;        // float4 groups per thread
    constexpr int NCOMP = (N / R) * L; // compute threads
    constexpr int NCW = NCOMP / 32;    // compute warps
    static_assert(COLS % 4 == 0 && NCOMP % 32 == 0, "bad split");

    extern __shared__ __align__(128) uint8_t smem_bytes[];
    Wkv7FwdSmem<NSTAGE>& sm = *reinterpret_cast<Wkv7FwdSmem<NSTAGE>*>(smem_bytes);

    const int hh = blockIdx.x, bb = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31;
    // broadcast from lane 0 so that ptxas knows the role branch below is warp-uniform (no BRA.DIV
    // in front of every shuffle)
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int T = p.T, H = p.H;
    const int nchunks = (T + TC - 1) / TC;

    if (tid == 0) {
        for (int i = 0; i < NSTAGE; i++) {
            mbar_init(&sm.full_raw[i], 1);
            mbar_init(&sm.empty_raw[i], NCONV);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&sm.full_f[i], NCONV);
            mbar_init(&sm.empty_f[i], NCW);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp >= NCW) {
        // =============================== converter / producer warps ===========================
        const int cw = warp - NCW;
        const bool producer = (cw == 0 && lane == 0);
        auto issue = [&](int c) {
            const int stage = c % NSTAGE;
            uint64_t* bar = &sm.full_raw[stage];
            mbar_arrive_expect_tx(bar, 6 * TC * N * 2);
            const int x0 = hh * N, y0 = bb * T + c * TC;
            tma_load_2d(&sm.raw[stage][0][0][0], &tm_w, x0, y0, bar);
            tma_load_2d(&sm.raw[stage][1][0][0], &tm_q, x0, y0, bar);
            tma_load_2d(&sm.raw[stage][2][0][0], &tm_k, x0, y0, bar);
            tma_load_2d(&sm.raw[stage][3][0][0], &tm_v, x0, y0, bar);
            tma_load_2d(&sm.raw[stage][4][0][0], &tm_a, x0, y0, bar);
            tma_load_2d(&sm.raw[stage][5][0][0], &tm_b, x0, y0, bar);
        };
        if (producer) {
            tma_prefetch_desc(&tm_w); tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k);
            tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_a); tma_prefetch_desc(&tm_b);
            for (int c = 0; c < NSTAGE && c < nchunks; c++) issue(c);
        }
        for (int c = 0; c < nchunks; c++) {
            const int stage = c % NSTAGE;
            const uint32_t rpar = (c / NSTAGE) & 1;
            const int buf = c & 1;
            mbar_wait(&sm.full_raw[stage], rpar);
            if (c >= 2) mbar_wait(&sm.empty_f[buf], ((c >> 1) - 1) & 1);
            // 6 tensors x 8 row-pairs of 128 elements; one LDS.64 -> STS.128 per lane per group
#pragma unroll 4
            for (int g = cw; g < 48; g += NCONV) {
                const int tensor = g >> 3, off = (g & 7) * 128 + lane * 4;
                const uint2 u = *reinterpret_cast<const uint2*>(&sm.raw[stage][tensor][0][0] + off);
                float4 o;
                o.x = bf16lo_to_f32(u.x); o.y = bf16hi_to_f32(u.x);
                o.z = bf16lo_to_f32(u.y); o.w = bf16hi_to_f32(u.y);
                if (tensor == 0) {  // decay = exp(-exp(w))  (wkv7_cuda.cu:21)
                    o.x = __expf(-__expf(o.x)); o.y = __expf(-__expf(o.y));
                    o.z = __expf(-__expf(o.z)); o.w = __expf(-__expf(o.w));
                }
                *reinterpret_cast<float4*>(&sm.f[buf][tensor][0][0] + off) = o;
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&sm.full_f[buf]);
                mbar_arrive(&sm.empty_raw[stage]);
            }
            if (producer && c + NSTAGE < nchunks) {
                mbar_wait(&sm.empty_raw[stage], rpar);
                issue(c + NSTAGE);
            }
        }
        return;
    }

    // ===================================== compute warps ======================================
    const int l = tid % L;          // column group
    const int i0 = (tid / L) * R;   // first row
    u64 S[R][M][2];
    if (p.state_in) {
        const float* src = p.state_in + ((size_t)bb * H + hh) * N * N;
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int m = 0; m < M; m++) {
                const float4 x = *reinterpret_cast<const float4*>(src + (i0 + r) * N + 4 * (l + L * m));
                S[r][m][0] = pk2(x.x, x.y);
                S[r][m][1] = pk2(x.z, x.w);
            }
    } else {
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int m = 0; m < M; m++) S[r][m][0] = S[r][m][1] = 0ull;
    }

    const bool write_sa = p.sa != nullptr;
    auto step = [&](const int buf, const int t, const size_t ind) {
        const float* f0 = &sm.f[buf][0][t][0];
        constexpr int TS = TC * N;  // tensor stride in floats
        float sa[R];
        {
            u64 acc[R][4];
#pragma unroll
            for (int r = 0; r < R; r++) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0ull;
#pragma unroll
            for (int m = 0; m < M; m++) {
                const float4 a4 = *reinterpret_cast<const float4*>(f0 + 4 * TS + 4 * (l + L * m));
                const u64 a01 = pk2(a4.x, a4.y), a23 = pk2(a4.z, a4.w);
#pragma unroll
                for (int r = 0; r < R; r++) {
                    acc[r][(2 * m) & 3] = ffma2(a01, S[r][m][0], acc[r][(2 * m) & 3]);
                    acc[r][(2 * m + 1) & 3] = ffma2(a23, S[r][m][1], acc[r][(2 * m + 1) & 3]);
                }
            }
#pragma unroll
            for (int r = 0; r < R; r++) {
                float x = hsum2(fadd2(fadd2(acc[r][0], acc[r][1]), fadd2(acc[r][2], acc[r][3])));
#pragma unroll
                for (int o = 1; o < L; o <<= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
                sa[r] = x;
            }
        }
        float vv[R];
#pragma unroll
        for (int r = 0; r < R; r++) vv[r] = f0[3 * TS + i0 + r];
        u64 yacc[R][4];
#pragma unroll
        for (int r = 0; r < R; r++) yacc[r][0] = yacc[r][1] = yacc[r][2] = yacc[r][3] = 0ull;
#pragma unroll
        for (int m = 0; m < M; m++) {
            const int co = 4 * (l + L * m);
            const float4 w4 = *reinterpret_cast<const float4*>(f0 + 0 * TS + co);
            const float4 k4 = *reinterpret_cast<const float4*>(f0 + 2 * TS + co);
            const float4 b4 = *reinterpret_cast<const float4*>(f0 + 5 * TS + co);
            const float4 q4 = *reinterpret_cast<const float4*>(f0 + 1 * TS + co);
            const u64 w01 = pk2(w4.x, w4.y), w23 = pk2(w4.z, w4.w);
            const u64 k01 = pk2(k4.x, k4.y), k23 = pk2(k4.z, k4.w);
            const u64 b01 = pk2(b4.x, b4.y), b23 = pk2(b4.z, b4.w);
            const u64 q01 = pk2(q4.x, q4.y), q23 = pk2(q4.z, q4.w);
#pragma unroll
            for (int r = 0; r < R; r++) {
                const u64 v2 = pk2(vv[r], vv[r]), sa2 = pk2(sa[r], sa[r]);
                // S = S*w + k*v + sa*b   (wkv7_cuda.cu:39)
                u64 s0 = ffma2(S[r][m][0], w01, fmul2(k01, v2));
                u64 s1 = ffma2(S[r][m][1], w23, fmul2(k23, v2));
                s0 = ffma2(sa2, b01, s0);
                s1 = ffma2(sa2, b23, s1);
                S[r][m][0] = s0;
                S[r][m][1] = s1;
                yacc[r][(2 * m) & 3] = ffma2(s0, q01, yacc[r][(2 * m) & 3]);
                yacc[r][(2 * m + 1) & 3] = ffma2(s1, q23, yacc[r][(2 * m + 1) & 3]);
            }
        }
        float yy[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            float x = hsum2(fadd2(fadd2(yacc[r][0], yacc[r][1]), fadd2(yacc[r][2], yacc[r][3])));
#pragma unroll
            for (int o = 1; o < L; o <<= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            yy[r] = x;
        }
        // branch-free stores: lane l==0 of a row group writes y, lane l==1 writes sa (every lane holds
        // the full sums after the butterfly)
        if constexpr (R == 2) {
            st_pred_b32(p.y + ind + i0, pack_bf16x2(yy[0], yy[1]), l == 0);
            st_pred_f32x2(p.sa + ind + i0, sa[0], sa[1], (l == 1) & write_sa);
        } else {
#pragma unroll
            for (int r = 0; r < R; r++) {
                st_pred_b16(p.y + ind + i0 + r, f32_to_bf16_bits(yy[r]), l == 0);
                st_pred_f32(p.sa + ind + i0 + r, sa[r], (l == 1) & write_sa);
            }
        }
    };

    for (int c = 0; c < nchunks; c++) {
        const int buf = c & 1;
        mbar_wait(&sm.full_f[buf], (c >> 1) & 1);
        __syncwarp();
        const size_t ind0 = (((size_t)bb * T + (size_t)c * TC) * H + hh) * N;
        const int nsteps = min(TC, T - c * TC);
        if (nsteps == TC) {
#pragma unroll 8
            for (int t = 0; t < TC; t++) step(buf, t, ind0 + (size_t)t * H * N);
            if (p.s) {  // transposed checkpoint: s[b,h,c,j,i] = S_ij  (wkv7_cuda.cu:44-50)
                float* dst = p.s + (((size_t)bb * H + hh) * (T / TC) + c) * N * N;
#pragma unroll
                for (int m = 0; m < M; m++)
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int j = 4 * (l + L * m) + e;
                        float x[R];
#pragma unroll
                        for (int r = 0; r < R; r++) {
                            float lo, hi;
                            upk2(S[r][m][e >> 1], lo, hi);
                            x[r] = (e & 1) ? hi : lo;
                        }
                        if constexpr (R == 2) {
                            *reinterpret_cast<float2*>(dst + j * N + i0) = make_float2(x[0], x[1]);
                        } else {
#pragma unroll
                            for (int r = 0; r < R; r++) dst[j * N + i0 + r] = x[r];
                        }
                    }
            }
        } else {
            for (int t = 0; t < nsteps; t++) step(buf, t, ind0 + (size_t)t * H * N);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty_f[buf]);
    }
    if (p.state_out) {
        float* dst = p.state_out + ((size_t)bb * H + hh) * N * N;
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int m = 0; m < M; m++) {
                float4 x;
                upk2(S[r][m][0], x.x, x.y);
                upk2(S[r][m][1], x.z, x.w);
                *reinterpret_cast<float4*>(dst + (i0 + r) * N + 4 * (l + L * m)) = x;
            }
    }
}

}  // namespace vrwkv
