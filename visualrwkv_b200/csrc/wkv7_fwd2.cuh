// wkv7_fwd2.cuh — WKV7 forward, second-generation mapping.
//
// Same contract as the reference forward_kernel (VisualRWKV-v7/v7.00/cuda/wkv7_cuda.cu:10-52): y (bf16), sa (fp32,
// every step) and the transposed fp32 state checkpoint every 16 steps.  What ncu taught us about this loop
// (profiles/r1a, r1b): with one head per SM it is bound by (i) the shared-memory pipe if every thread re-reads
// fp32 per-column vectors, (ii) in-order issue — a warp stalls on every shuffle/LDS consumer, and nothing else
// runs on that SM sub-partition.  Hence:
//
//   * a thread owns an R-row x 8-column block of the 64x64 fp32 state; a head is (64/R) row groups x 8 column
//     groups: R=2 -> 256 threads = 8 warps (two per sub-partition, so one warp's stalls are filled by the
//     other), R=4 -> 128 threads;
//   * q,k,a,b,v are read as bf16 straight from the TMA-staged tile (one LDS.128 = 8 columns) and expanded on the
//     ALU pipe; only the decay exp(-exp(w)) is precomputed in fp32, once per (t, column), by a converter warp,
//     into a column-permuted tile whose LDS.128 are bank-conflict free;
//   * step t updates S, then computes y_t = S.q_t AND the next step's sa_{t+1} = S.a_{t+1} in the same pass and
//     reduces both with one interleaved transposing reduce-scatter over the 8 column-group lanes (their shuffle
//     latencies overlap), sa is all-gathered back; stores are branch-free from distinct lanes;
//   * packed fp32x2 math (FFMA2) everywhere.
//
// Per step and head (one SM): FMA pipe >= 160 cycles (5 FMA x 4096 / 128 lanes).  Algorithmic HBM bytes: 14 B
// per (b,t,c) element; the reference contract adds 4 B (sa) + 16 B (s).
#pragma once
#include "common.cuh"
#include "wkv7_fwd.cuh"

namespace vrwkv {

template <int NSTAGE>
struct alignas(128) Wkv7Fwd2Smem {
    uint16_t raw[NSTAGE][6][WKV_TC][WKV_N];  // TMA ring (w,q,k,v,a,b), 2 KB per tile
    float decay[2][WKV_TC][WKV_N];           // exp(-exp(w)), column-permuted
    uint64_t full_raw[NSTAGE], empty_raw[NSTAGE], full_w[2], empty_w[2];
};

__device__ __forceinline__ u64 bf2_to_f2(uint32_t x) {  // packed bf16x2 -> packed f32x2
    return pk2(__uint_as_float(x << 16), __uint_as_float(x & 0xffff0000u));
}
__device__ __forceinline__ void unpack8(const uint4 u, u64 (&o)[4]) {
    o[0] = bf2_to_f2(u.x); o[1] = bf2_to_f2(u.y); o[2] = bf2_to_f2(u.z); o[3] = bf2_to_f2(u.w);
}

// Lane geometry shared by the forward and backward kernels: the 8 lanes l = 0..7 (bits b2 b1 b0) of a row group
// hold the 8 column blocks of R rows.  A transposing reduce-scatter leaves on every lane the complete sum of row
// `own_row` = (R==4 ? 2*b2+b1 : b2); NSUB = 8/R lanes hold the same row and are told apart by `sub`.
template <int R>
struct RowLanes {
    static_assert(R == 2 || R == 4, "R must be 2 or 4");
    static constexpr int NSUB = 8 / R;
    bool b1, b2;
    int own_row, sub;
    __device__ __forceinline__ explicit RowLanes(int l) {
        b1 = (l >> 1) & 1;
        b2 = (l >> 2) & 1;
        own_row = (R == 4) ? (2 * (int)b2 + (int)b1) : (int)b2;
        sub = l & (NSUB - 1);
    }
    // K independent sets of R row sums, interleaved stage by stage so that the shuffle latencies overlap
    template <int K>
    __device__ __forceinline__ void reduce(const float (&x)[K][R], float (&z)[K]) const {
        float r[K];
        if constexpr (R == 4) {
            float s0[K], k0[K], s1[K], k1[K], r0[K], r1[K], s[K], kk[K];
#pragma unroll
            for (int k = 0; k < K; k++) {
                s0[k] = b2 ? x[k][0] : x[k][2]; k0[k] = b2 ? x[k][2] : x[k][0];
                s1[k] = b2 ? x[k][1] : x[k][3]; k1[k] = b2 ? x[k][3] : x[k][1];
            }
#pragma unroll
            for (int k = 0; k < K; k++) {
                r0[k] = __shfl_xor_sync(0xffffffffu, s0[k], 4);
                r1[k] = __shfl_xor_sync(0xffffffffu, s1[k], 4);
            }
#pragma unroll
            for (int k = 0; k < K; k++) {
                const float y0 = k0[k] + r0[k], y1 = k1[k] + r1[k];
                s[k] = b1 ? y0 : y1;
                kk[k] = b1 ? y1 : y0;
            }
#pragma unroll
            for (int k = 0; k < K; k++) r[k] = __shfl_xor_sync(0xffffffffu, s[k], 2);
#pragma unroll
            for (int k = 0; k < K; k++) z[k] = kk[k] + r[k];
        } else {
            float s[K], kk[K];
#pragma unroll
            for (int k = 0; k < K; k++) {
                s[k] = b2 ? x[k][0] : x[k][1];
                kk[k] = b2 ? x[k][1] : x[k][0];
            }
#pragma unroll
            for (int k = 0; k < K; k++) r[k] = __shfl_xor_sync(0xffffffffu, s[k], 4);
#pragma unroll
            for (int k = 0; k < K; k++) z[k] = kk[k] + r[k];
#pragma unroll
            for (int k = 0; k < K; k++) r[k] = __shfl_xor_sync(0xffffffffu, z[k], 2);
#pragma unroll
            for (int k = 0; k < K; k++) z[k] += r[k];
        }
#pragma unroll
        for (int k = 0; k < K; k++) r[k] = __shfl_xor_sync(0xffffffffu, z[k], 1);
#pragma unroll
        for (int k = 0; k < K; k++) z[k] += r[k];
    }
    // from "lane holds the sum of its own row" to "every lane holds all R row sums"
    __device__ __forceinline__ void allgather(float z, float (&o)[R]) const {
        if constexpr (R == 4) {
            const float t = __shfl_xor_sync(0xffffffffu, z, 2);
            const float lo = b1 ? t : z, hi = b1 ? z : t;  // rows 2*b2, 2*b2+1
            const float olo = __shfl_xor_sync(0xffffffffu, lo, 4), ohi = __shfl_xor_sync(0xffffffffu, hi, 4);
            o[0] = b2 ? olo : lo; o[1] = b2 ? ohi : hi; o[2] = b2 ? lo : olo; o[3] = b2 ? hi : ohi;
        } else {
            const float t = __shfl_xor_sync(0xffffffffu, z, 4);
            o[0] = b2 ? t : z;
            o[1] = b2 ? z : t;
        }
    }
};

template <int R, int NSTAGE, int UNROLL = 5>
__global__ void __launch_bounds__((WKV_N / R) * 8 + 32)
wkv7_fwd2_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_q,
                 const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                 const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const Wkv7FwdArgs p) {
    constexpr int N = WKV_N, TC = WKV_TC;
    constexpr int NCW = (N / R) * 8 / 32;  // compute warps
    constexpr int TS = TC * N;             // tensor stride (elements) inside a raw stage
    extern __shared__ __align__(128) uint8_t smem_bytes[];
    Wkv7Fwd2Smem<NSTAGE>& sm = *reinterpret_cast<Wkv7Fwd2Smem<NSTAGE>*>(smem_bytes);

    const int hh = blockIdx.x, bb = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for ptxas: no BRA.DIV before shuffles
    const int T = p.T, H = p.H;
    const int nchunks = (T + TC - 1) / TC;

    if (tid == 0) {
        for (int i = 0; i < NSTAGE; i++) {
            mbar_init(&sm.full_raw[i], 1);
            mbar_init(&sm.empty_raw[i], NCW + 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&sm.full_w[i], 1);
            mbar_init(&sm.empty_w[i], NCW);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == NCW) {
        // ================= producer + decay converter warp =================
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
        if (lane == 0) {
            tma_prefetch_desc(&tm_w); tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k);
            tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_a); tma_prefetch_desc(&tm_b);
            for (int c = 0; c < NSTAGE && c < nchunks; c++) issue(c);
        }
        for (int c = 0; c < nchunks; c++) {
            const int stage = c % NSTAGE, buf = c & 1;
            mbar_wait(&sm.full_raw[stage], (c / NSTAGE) & 1);
            if (c >= 2) mbar_wait(&sm.empty_w[buf], ((c >> 1) - 1) & 1);
            __syncwarp();
            // 16 rows x 64 columns, 4 consecutive columns of 2 rows per lane per pass.  Column j = 8*l + e goes to
            // position (e < 4 ? 4*l + e : 32 + 4*l + e - 4): a compute thread's two LDS.128 (columns 8l..8l+3 and
            // 8l+4..8l+7) then hit 32 distinct banks across the 8 lanes of a row group.
#pragma unroll
            for (int g = 0; g < 8; g++) {
                const int row = 2 * g + (lane >> 4), m = lane & 15;  // m: group of 4 columns
                const uint2 u = *reinterpret_cast<const uint2*>(&sm.raw[stage][0][row][4 * m]);
                float4 o;  // decay = exp(-exp(w))   (wkv7_cuda.cu:21)
                o.x = __expf(-__expf(__uint_as_float(u.x << 16)));
                o.y = __expf(-__expf(__uint_as_float(u.x & 0xffff0000u)));
                o.z = __expf(-__expf(__uint_as_float(u.y << 16)));
                o.w = __expf(-__expf(__uint_as_float(u.y & 0xffff0000u)));
                const int pos = (m & 1) * 32 + 4 * (m >> 1);
                *reinterpret_cast<float4*>(&sm.decay[buf][row][pos]) = o;
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&sm.full_w[buf]);
                mbar_arrive(&sm.empty_raw[stage]);
                // refill the stage of the PREVIOUS chunk (the compute warps have left it by now)
                if (c >= 1 && c - 1 + NSTAGE < nchunks) {
                    mbar_wait(&sm.empty_raw[(c - 1) % NSTAGE], ((c - 1) / NSTAGE) & 1);
                    issue(c - 1 + NSTAGE);
                }
            }
            __syncwarp();
        }
        return;
    }

    // ===================================== compute warps ======================================
    const int l = tid & 7;          // column group: columns 8l .. 8l+7
    const int i0 = (tid >> 3) * R;  // first row of this thread's row group
    const RowLanes<R> L(l);
    const int myrow = i0 + L.own_row;
    u64 S[R][4];  // [row][column pair]
    if (p.state_in) {
        const float* src = p.state_in + ((size_t)bb * H + hh) * N * N;
#pragma unroll
        for (int r = 0; r < R; r++) {
            const float4 x0 = *reinterpret_cast<const float4*>(src + (i0 + r) * N + 8 * l);
            const float4 x1 = *reinterpret_cast<const float4*>(src + (i0 + r) * N + 8 * l + 4);
            S[r][0] = pk2(x0.x, x0.y); S[r][1] = pk2(x0.z, x0.w);
            S[r][2] = pk2(x1.x, x1.y); S[r][3] = pk2(x1.z, x1.w);
        }
    } else {
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) S[r][c] = 0ull;
    }
    const bool st_y = L.sub == 0, st_sa = (L.sub == 1) && (p.sa != nullptr);

    float sa[R];  // sa_t of the step about to run, for this thread's rows (all-gathered)

    // One step: S <- S*w + k*v + sa*b ; y_t = S.q_t ; and, in the same pass and the same interleaved reduction, the
    // NEXT step's sa_{t+1} = S.a_{t+1}   (wkv7_cuda.cu:27-42).
    auto step = [&](const int stage, const int buf, const int t, const size_t ind, const bool has_next) {
        const uint16_t* r0 = &sm.raw[stage][0][t][0];
        const float* dw = &sm.decay[buf][t][0];
        const float4 w0 = *reinterpret_cast<const float4*>(dw + 4 * l);
        const float4 w1 = *reinterpret_cast<const float4*>(dw + 32 + 4 * l);
        u64 k2[4], b2v[4], q2[4], an[4];
        unpack8(*reinterpret_cast<const uint4*>(r0 + 2 * TS + 8 * l), k2);
        unpack8(*reinterpret_cast<const uint4*>(r0 + 5 * TS + 8 * l), b2v);
        unpack8(*reinterpret_cast<const uint4*>(r0 + 1 * TS + 8 * l), q2);
        if (has_next) unpack8(*reinterpret_cast<const uint4*>(r0 + 4 * TS + N + 8 * l), an);  // a of row t+1
        float vf[R];
        if constexpr (R == 4) {
            const uint2 vu = *reinterpret_cast<const uint2*>(r0 + 3 * TS + i0);
            vf[0] = __uint_as_float(vu.x << 16); vf[1] = __uint_as_float(vu.x & 0xffff0000u);
            vf[2] = __uint_as_float(vu.y << 16); vf[3] = __uint_as_float(vu.y & 0xffff0000u);
        } else {
            const uint32_t vu = *reinterpret_cast<const uint32_t*>(r0 + 3 * TS + i0);
            vf[0] = __uint_as_float(vu << 16); vf[1] = __uint_as_float(vu & 0xffff0000u);
        }
        const u64 w2[4] = {pk2(w0.x, w0.y), pk2(w0.z, w0.w), pk2(w1.x, w1.y), pk2(w1.z, w1.w)};
        float x[2][R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const u64 v2 = pk2(vf[r], vf[r]), sa2 = pk2(sa[r], sa[r]);
            u64 ya = 0ull, yb = 0ull, na = 0ull, nb = 0ull;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                u64 s = ffma2(S[r][c], w2[c], fmul2(k2[c], v2));
                s = ffma2(sa2, b2v[c], s);
                S[r][c] = s;
                if (c & 1) yb = ffma2(s, q2[c], yb); else ya = ffma2(s, q2[c], ya);
                if (has_next) {
                    if (c & 1) nb = ffma2(s, an[c], nb); else na = ffma2(s, an[c], na);
                }
            }
            x[0][r] = hsum2(fadd2(ya, yb));
            x[1][r] = hsum2(fadd2(na, nb));
        }
        if (has_next) {
            float z[2];
            L.template reduce<2>(x, z);
            st_pred_b16(p.y + ind + myrow, f32_to_bf16_bits(z[0]), st_y);
            st_pred_f32(p.sa + ind + (size_t)H * N + myrow, z[1], st_sa);
            L.allgather(z[1], sa);
        } else {
            float x1[1][R], z[1];
#pragma unroll
            for (int r = 0; r < R; r++) x1[0][r] = x[0][r];
            L.template reduce<1>(x1, z);
            st_pred_b16(p.y + ind + myrow, f32_to_bf16_bits(z[0]), st_y);
        }
    };

    for (int c = 0; c < nchunks; c++) {
        const int stage = c % NSTAGE, buf = c & 1;
        mbar_wait(&sm.full_raw[stage], (c / NSTAGE) & 1);
        mbar_wait(&sm.full_w[buf], (c >> 1) & 1);
        __syncwarp();
        const size_t ind0 = (((size_t)bb * T + (size_t)c * TC) * H + hh) * N;
        const int nsteps = min(TC, T - c * TC);
        {   // sa of the first step of the chunk: a plain dot + reduce (once per 16 steps)
            u64 a2[4];
            unpack8(*reinterpret_cast<const uint4*>(&sm.raw[stage][4][0][0] + 8 * l), a2);
            float x[1][R], z[1];
#pragma unroll
            for (int r = 0; r < R; r++) {
                // same association as the fused next-sa dot inside step() (pairs 0,2 and 1,3): a sequence split at
                // any point reproduces the one-shot run bit for bit
                const u64 e = ffma2(S[r][2], a2[2], fmul2(S[r][0], a2[0]));
                const u64 o = ffma2(S[r][3], a2[3], fmul2(S[r][1], a2[1]));
                x[0][r] = hsum2(fadd2(e, o));
            }
            L.template reduce<1>(x, z);
            st_pred_f32(p.sa + ind0 + myrow, z[0], st_sa);
            L.allgather(z[0], sa);
        }
        if (nsteps == TC) {
#pragma unroll UNROLL
            for (int t = 0; t < TC - 1; t++) step(stage, buf, t, ind0 + (size_t)t * H * N, true);
            step(stage, buf, TC - 1, ind0 + (size_t)(TC - 1) * H * N, false);
            if (p.s) {  // transposed checkpoint: s[b,h,c,j,i] = S_ij  (wkv7_cuda.cu:44-50)
                float* dst = p.s + (((size_t)bb * H + hh) * (T / TC) + c) * N * N + i0;
#pragma unroll
                for (int cp = 0; cp < 4; cp++) {
                    float lo[R], hi[R];
#pragma unroll
                    for (int r = 0; r < R; r++) upk2(S[r][cp], lo[r], hi[r]);
                    const int j = 8 * l + 2 * cp;
                    if constexpr (R == 4) {
                        *reinterpret_cast<float4*>(dst + (size_t)j * N) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                        *reinterpret_cast<float4*>(dst + (size_t)(j + 1) * N) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    } else {
                        *reinterpret_cast<float2*>(dst + (size_t)j * N) = make_float2(lo[0], lo[1]);
                        *reinterpret_cast<float2*>(dst + (size_t)(j + 1) * N) = make_float2(hi[0], hi[1]);
                    }
                }
            }
        } else {
            for (int t = 0; t < nsteps; t++) step(stage, buf, t, ind0 + (size_t)t * H * N, t + 1 < nsteps);
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&sm.empty_w[buf]);
            mbar_arrive(&sm.empty_raw[stage]);
        }
    }
    if (p.state_out) {
        float* dst = p.state_out + ((size_t)bb * H + hh) * N * N;
#pragma unroll
        for (int r = 0; r < R; r++) {
            float4 x0, x1;
            upk2(S[r][0], x0.x, x0.y); upk2(S[r][1], x0.z, x0.w);
            upk2(S[r][2], x1.x, x1.y); upk2(S[r][3], x1.z, x1.w);
            *reinterpret_cast<float4*>(dst + (i0 + r) * N + 8 * l) = x0;
            *reinterpret_cast<float4*>(dst + (i0 + r) * N + 8 * l + 4) = x1;
        }
    }
}

}  // namespace vrwkv
