// wkv7_chunk_bwd.cuh — chunk-local WKV7 backward on the tcgen05 tensor cores: one CTA per (batch, head, 64-step
// chunk), all chunks concurrent.  Inputs besides the seven bf16 streams: the forward's `sa` (rows of U) and 16-step
// state checkpoints `s` (S at the start / end of the chunk), and dL/dS at the end of the chunk from
// wkv7_chunk_dstate.cuh.  Replaces the reference's serial reverse-time walk (VisualRWKV-v7/v7.00/cuda/wkv7_cuda.cu:
// 54-130) for callers that promise bounded decay; the algebra is the hand-derived reverse pass of the chunk form
// (the tests' chunk_backward restatement, checked against the fp64 adjoint of the step-by-step oracle):
//     dZ = dS_L e^{G_L}                      [dU;dV] = [Bt;Kt] dZ^T + [A_qb;A_qk]^T dY        dR = (I - A_ab)^-T dU
//     dA = mask([dR;dY] [U;V]^T)             dV += A_ak^T dR
//     [dAt;dQt] = [dR;dY] S_0 + dA [Bt;Kt]   [dBt;dKt] = [U;V] dZ + dA^T [At;Qt]
//     da = dAt e^{G_{t-1}}, dq = dQt e^{G_t}, dk = dKt e^{-G_t}, db = dBt e^{-G_t}
//     dG_t = dq q - dk k - db b + (da a)_{t+1}  (+ sum_i dS_L S_L at t = 63);  dw = (reverse cumsum of dG) * (-e^w)
// TF32 operands, fp32 accumulation in TMEM; same operand layouts and helpers as wkv7_chunk_fwd.cuh.  Buffers are reused
// phase by phase (the comments on Wkv7ChunkBwdSmem give the lifetimes); operands with 64 real rows are issued as
// M = 128 products whose upper accumulator rows are never read.
#pragma once
#include "wkv7_chunk_common.cuh"

namespace vrwkv {

struct Wkv7ChunkBwdArgs {
    int B, T, H;
    const uint16_t *w, *q, *k, *v, *a, *b, *dy;  // for the direct (non-TMA) re-reads
    const float* sa;                             // [B,T,H,64]
    const float* s;                              // transposed state checkpoints, `ck_per_chunk` per 64-step chunk (4: the
                                                 // reference's [B,H,T/16,64,64]; 1: [B,H,T/64,64,64] chunk-end states only)
    int ck_per_chunk;
    const float* ds;                             // [B,H,T/64,64,64] dL/dS at the end of each chunk (last: not read)
    float* gws;                                  // [B,H,T/64,64,64] scratch: G_t of every chunk
    uint16_t *dw, *dq, *dk, *dv, *da, *db;
};

struct alignas(1024) Wkv7ChunkBwdSmem {
    uint8_t aq[32768];   // [At;Qt]: K-major for the scores, re-swizzled in place to MN-major for dA^T [At;Qt]
    uint8_t bk[32768];   // [Bt;Kt]: K-major for the scores and [Bt;Kt] dZ^T, then MN-major for dA [Bt;Kt]
    uint8_t dz[16384];   // dZ [i][j]: K-major for [Bt;Kt] dZ^T, then MN-major for [U;V] dZ; then (with x16) dA (q rows) MN
    uint8_t x16[16384];  // dY (MN) -> dU (MN) -> dR (MN)
    uint8_t sa_[32768];  // a-row scores [t][A_ab | A_ak], MN-major (M = column); then dA (q rows) K-major
    uint8_t z1[32768];   // TMA tiles 0-3 | q-row scores [t][A_qb | A_qk] MN | [dR;dY] K-major | dA (a rows) K-major | epilogue
    uint8_t z2[32768];   // TMA tiles 4-6 | A_ab fp32 (16 KB) + Tinv MN (16 KB) | [U;V] K-major | dA (a rows) MN | epilogue
    uint8_t s0[16384];   // S_0 as stored by the forward ([j][i]) = K-major B operand of [dR;dY] S_0; Tinv's M=128 tail
    float esc[32 * 32];
    float part[8][WKV_N];
    float el[WKV_N];
    float gl[WKV_N];
    uint64_t bar_in, bar_mma;  // bar_mma: one arrival per issuing warp (warps 0-3)
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(CK_THREADS, 1)
wkv7_chunk_bwd_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_q,
                      const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                      const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                      const __grid_constant__ CUtensorMap tm_dy, const Wkv7ChunkBwdArgs p) {
    constexpr int N = WKV_N, L = CK_L;
    extern __shared__ __align__(1024) uint8_t chunk_bwd_smem_bytes[];
    Wkv7ChunkBwdSmem& sm = *reinterpret_cast<Wkv7ChunkBwdSmem*>(chunk_bwd_smem_bytes);
    uint8_t* const tiles = sm.z1;         // 7 x 8 KB (runs on into z2)
    uint8_t* const sq = sm.z1;            // q-row scores MN: 4 column blocks x 64 k-lines
    uint8_t* const ry = sm.z1;            // [dR;dY] K-major: 2 k-atoms x 128 rows
    uint8_t* const aab = sm.z2;           // A_ab fp32
    uint8_t* const tinv = sm.z2 + 16384;  // Tinv MN: 2 column blocks x 64 k-lines (M = 128 reads on into s0)
    uint8_t* const uv = sm.z2;            // [U;V] K-major: 2 k-atoms x 128 rows

    const int hh = blockIdx.x, bb = blockIdx.y, c = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int qd = warp & 3, cs = warp >> 2;
    const int r = 32 * qd + lane;
    const int T = p.T, H = p.H;
    const int nch = T / L;
    const size_t row0 = ((size_t)bb * T + (size_t)c * L) * H * N + (size_t)hh * N;  // element offset of (t = 0, channel 0)
    const size_t rstride = (size_t)H * N;
    const size_t chunk_id = ((size_t)bb * H + hh) * nch + c;

    if (tid == 0) {
        mbar_init(&sm.bar_in, 1);
        mbar_init(&sm.bar_mma, 4);
        fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tmem_alloc<512>(&sm.tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    const uint32_t tm_row = tmem + ((uint32_t)(32 * qd) << 16);
    // TMEM columns: scores 0-127 (then dA^T), [dU;dV] 128-191, dR 192-255, dA 256-383, [dBt;dKt] 384-447, [dAt;dQt] 448-511
    constexpr uint32_t C_SC = 0, C_ACC1 = 128, C_R = 192, C_DA = 256, C_DAT = 0, C_BK = 384, C_AQ = 448;

    if (tid == 0) {
        mbar_arrive_expect_tx(&sm.bar_in, 7 * L * N * 2);
        const int x0 = hh * N, y0 = bb * T + c * L;
        tma_load_2d(tiles + 0 * 8192, &tm_w, x0, y0, &sm.bar_in);
        tma_load_2d(tiles + 1 * 8192, &tm_q, x0, y0, &sm.bar_in);
        tma_load_2d(tiles + 2 * 8192, &tm_k, x0, y0, &sm.bar_in);
        tma_load_2d(tiles + 3 * 8192, &tm_v, x0, y0, &sm.bar_in);
        tma_load_2d(tiles + 4 * 8192, &tm_a, x0, y0, &sm.bar_in);
        tma_load_2d(tiles + 5 * 8192, &tm_b, x0, y0, &sm.bar_in);
        tma_load_2d(tiles + 6 * 8192, &tm_dy, x0, y0, &sm.bar_in);
    }
    __syncwarp();

#ifdef VRWKV_PHASE_STAMPS   // development builds only (VRWKV_PHASE_STAMPS=1): per-phase clocks of chunk 1 of CTA (0,0)
    float* const dbg = (hh == 0 && bb == 0 && c == 1) ? g_chunk_dbg : nullptr;
    const long long tstamp0 = clock64();
    int tsi = 0;
    auto stamp = [&]() {
        if (dbg && tid == 0) dbg[3072 + tsi++] = (float)(clock64() - tstamp0);
    };
#else
    auto stamp = [] {};
#endif
    {   // the delta phase reads U (= sa rows) and S_0 from global memory: start pulling them into L2 now
        const int t = tid >> 3, i0 = 8 * (tid & 7);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.sa + row0 + (size_t)t * rstride + i0));
        if (c > 0)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.s + ((((size_t)bb * H + hh) * nch * p.ck_per_chunk + (size_t)c * p.ck_per_chunk - 1) * N + t) * N + i0));
    }
    uint32_t mph = 0;
    auto mma_wait = [&]() {
        stamp();
        mbar_wait(&sm.bar_mma, mph & 1);
        mph++;
        tc_fence_after();
        __syncwarp();
        stamp();
    };
    auto operands_ready = [&]() {
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
    };
    const uint32_t b4 = smem_u32(&sm) >> 4;
    const uint32_t O_AQ = (uint32_t)(sm.aq - (uint8_t*)&sm), O_BK = (uint32_t)(sm.bk - (uint8_t*)&sm), O_DZ = (uint32_t)(sm.dz - (uint8_t*)&sm),
                   O_X16 = (uint32_t)(sm.x16 - (uint8_t*)&sm), O_SA = (uint32_t)(sm.sa_ - (uint8_t*)&sm), O_Z1 = (uint32_t)(sm.z1 - (uint8_t*)&sm),
                   O_Z2 = (uint32_t)(sm.z2 - (uint8_t*)&sm), O_S0 = (uint32_t)(sm.s0 - (uint8_t*)&sm);
    const bool issuer = lane == 0 && warp < 4;  // the four threads that issue tensor-core work (independent accumulators)
    constexpr uint32_t ID_KK_128 = umma_idesc_tf32(128, 128), ID_KK = umma_idesc_tf32(128, 64), ID_KM = umma_idesc_tf32(128, 64, 0, 1),
                       ID_MM = umma_idesc_tf32(128, 64, 1, 1);
    auto f4 = [](const uint32_t* v) { return make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3])); };

    stamp();
    mbar_wait(&sm.bar_in, 0);
    stamp();
    // ================= P1: decay prefix sums, scaled operands, dY operand, G to the workspace =================
    {
        const int hf = warp & 1, rg = warp >> 1;
        const int j = 32 * hf + lane;
        const uint16_t* in16 = reinterpret_cast<const uint16_t*>(tiles);
        float g[8];
        float loc = 0.f;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            g[k] = -__expf(bf16lo_to_f32(in16[0 * 4096 + (8 * rg + k) * N + j]));
            loc += g[k];
        }
        sm.part[rg][j] = loc;
        __syncthreads();
        float G = 0.f;
#pragma unroll
        for (int k = 0; k < 7; k++) G += (k < rg) ? sm.part[k][j] : 0.f;
        float Eprev = __expf(G);
        float* gout = p.gws + chunk_id * (L * N) + j;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int t = 8 * rg + k;
            G += g[k];
            const float E = __expf(G), F = __expf(-G);
            const float qv = bf16lo_to_f32(in16[1 * 4096 + t * N + j]), kv = bf16lo_to_f32(in16[2 * 4096 + t * N + j]);
            const float av_ = bf16lo_to_f32(in16[4 * 4096 + t * N + j]), bv = bf16lo_to_f32(in16[5 * 4096 + t * N + j]);
            const float dyv = bf16lo_to_f32(in16[6 * 4096 + t * N + j]);
            *reinterpret_cast<float*>(sm.aq + hf * 16384 + sw128_off(t, lane)) = rt32(av_ * Eprev);
            *reinterpret_cast<float*>(sm.aq + hf * 16384 + sw128_off(64 + t, lane)) = rt32(qv * E);
            *reinterpret_cast<float*>(sm.bk + hf * 16384 + sw128_off(t, lane)) = rt32(bv * F);
            *reinterpret_cast<float*>(sm.bk + hf * 16384 + sw128_off(64 + t, lane)) = rt32(kv * F);
            *reinterpret_cast<float*>(sm.x16 + hf * 8192 + sw32_off(t, lane)) = dyv;
            gout[(size_t)t * N] = G;
            if (t == L - 1) {
                sm.el[j] = E;
                if (G < -80.f) g_chunk_domain_err = 1;
            }
            Eprev = E;
        }
    }
    __syncthreads();
    // dZ = dS_L diag(e^{G_L}) as the K-major B operand [i][j] (zero at the end of the sequence)
    {
        const int i = tid >> 3, j0 = 8 * (tid & 7);
        float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
        if (c + 1 < nch) {
            const float4* src = reinterpret_cast<const float4*>(p.ds + chunk_id * (N * N) + i * N + j0);
            x0 = __ldg(src);
            x1 = __ldg(src + 1);
        }
        const float4 e0 = *reinterpret_cast<const float4*>(&sm.el[j0]), e1 = *reinterpret_cast<const float4*>(&sm.el[j0 + 4]);
        x0.x *= e0.x; x0.y *= e0.y; x0.z *= e0.z; x0.w *= e0.w;
        x1.x *= e1.x; x1.y *= e1.y; x1.z *= e1.z; x1.w *= e1.w;
        uint8_t* dst = sm.dz + (j0 >> 5) * 8192 + i * 128;
        const int ch = (j0 & 31) >> 2;
        *reinterpret_cast<float4*>(dst + ((ch ^ (i & 7)) << 4)) = rt32(x0);
        *reinterpret_cast<float4*>(dst + (((ch + 1) ^ (i & 7)) << 4)) = rt32(x1);
    }
    operands_ready();
    // ================= alpha: scores = [At;Qt][Bt;Kt]^T ; ACC1 = [Bt;Kt] dZ^T =================
    if (issuer) {
        tc_fence_after();
        if (warp == 0) {
#pragma unroll
            for (int k = 0; k < 8; k++)
                umma_tf32(tmem + C_SC, desc_km(b4, O_AQ + (k >> 2) * 16384 + (k & 3) * 32), desc_km(b4, O_BK + (k >> 2) * 16384 + (k & 3) * 32),
                          ID_KK_128, k > 0);
        } else if (warp == 1) {
#pragma unroll
            for (int k = 0; k < 8; k++)
                umma_tf32(tmem + C_ACC1, desc_km(b4, O_BK + (k >> 2) * 16384 + (k & 3) * 32), desc_km(b4, O_DZ + (k >> 2) * 8192 + (k & 3) * 32),
                          ID_KK, k > 0);
        }
        umma_commit(&sm.bar_mma);
    }
    mma_wait();
    // ================= P2: masked scores as MN-major operands; A_ab in fp32; zero Tinv =================
    {
        uint32_t v[32];
        tmem_ld32(tm_row + C_SC + 32 * cs, v);
        const int t = r & 63;
        const bool incl = r >= 64;
        float o[32];
#pragma unroll
        for (int e = 0; e < 32; e++) {
            const int s = 32 * (cs & 1) + e;
            const bool keep = incl ? (s <= t) : (s < t);
            o[e] = keep ? __uint_as_float(v[e]) : 0.f;
        }
        uint8_t* dst = (incl ? sq : sm.sa_) + cs * 8192;
#pragma unroll
        for (int c4 = 0; c4 < 8; c4++)
            *reinterpret_cast<float4*>(dst + sw32_off(t, 4 * c4)) = rt32(make_float4(o[4 * c4], o[4 * c4 + 1], o[4 * c4 + 2], o[4 * c4 + 3]));
        if (!incl && cs < 2) {
#pragma unroll
            for (int c4 = 0; c4 < 8; c4++)
                *reinterpret_cast<float4*>(aab + t * 256 + (((8 * cs + c4) ^ (t & 7)) << 4)) =
                    make_float4(o[4 * c4], o[4 * c4 + 1], o[4 * c4 + 2], o[4 * c4 + 3]);
        }
        *reinterpret_cast<float4*>(tinv + tid * 32) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(tinv + tid * 32 + 16) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // [At;Qt], [Bt;Kt], dZ: K-major (SWIZZLE_128B) -> MN-major (SWIZZLE_128B_BASE32B), in place (same 128-byte lines)
    {
        float4 xa[4], xb[4], xz[2];
#pragma unroll
        for (int n = 0; n < 4; n++) {
            const int qq = tid + 512 * n, atom = qq >> 10, row = (qq >> 3) & 127, c16 = qq & 7;
            xa[n] = *reinterpret_cast<const float4*>(sm.aq + atom * 16384 + row * 128 + ((c16 ^ (row & 7)) << 4));
            xb[n] = *reinterpret_cast<const float4*>(sm.bk + atom * 16384 + row * 128 + ((c16 ^ (row & 7)) << 4));
        }
#pragma unroll
        for (int n = 0; n < 2; n++) {
            const int qq = tid + 512 * n, atom = qq >> 9, row = (qq >> 3) & 63, c16 = qq & 7;
            xz[n] = *reinterpret_cast<const float4*>(sm.dz + atom * 8192 + row * 128 + ((c16 ^ (row & 7)) << 4));
        }
        __syncthreads();
#pragma unroll
        for (int n = 0; n < 4; n++) {
            const int qq = tid + 512 * n, atom = qq >> 10, row = (qq >> 3) & 127, c16 = qq & 7;
            const int pos = ((((c16 >> 1) ^ (row & 3)) << 1) + (c16 & 1)) << 4;
            *reinterpret_cast<float4*>(sm.aq + atom * 16384 + row * 128 + pos) = xa[n];
            *reinterpret_cast<float4*>(sm.bk + atom * 16384 + row * 128 + pos) = xb[n];
        }
#pragma unroll
        for (int n = 0; n < 2; n++) {
            const int qq = tid + 512 * n, atom = qq >> 9, row = (qq >> 3) & 63, c16 = qq & 7;
            const int pos = ((((c16 >> 1) ^ (row & 3)) << 1) + (c16 & 1)) << 4;
            *reinterpret_cast<float4*>(sm.dz + atom * 8192 + row * 128 + pos) = xz[n];
        }
    }
    operands_ready();
    // ================= beta: ACC1 += [A_qb|A_qk]^T dY   (inverse meanwhile) =================
    if (issuer) {
        tc_fence_after();
        if (warp == 0) {
#pragma unroll
            for (int k = 0; k < 8; k++)
                umma_tf32(tmem + C_ACC1, desc_mn(b4, O_Z1 + k * 1024, 8192), desc_mn(b4, O_X16 + k * 1024, 8192), ID_MM, 1);
        }
        umma_commit(&sm.bar_mma);
    }
    stamp();
    chunk_tri_inverse(aab, sm.esc, tid, [&](int t, int s) { return tinv + (s >> 5) * 8192 + sw32_off(t, s & 31); });
    mma_wait();
    if (r < 64) {  // dU -> operand
        uint32_t v[16];
        tmem_ld16(tm_row + C_ACC1 + 16 * cs, v);
#pragma unroll
        for (int c4 = 0; c4 < 4; c4++)
            *reinterpret_cast<float4*>(sm.x16 + (cs >> 1) * 8192 + sw32_off(r, 16 * (cs & 1) + 4 * c4)) = rt32(f4(v + 4 * c4));
    }
    operands_ready();
    // ================= gamma: dR = Tinv^T dU =================
    if (issuer) {
        tc_fence_after();
        if (warp == 0) {
#pragma unroll
            for (int k = 0; k < 8; k++)
                umma_tf32(tmem + C_R, desc_mn(b4, O_Z2 + 16384 + k * 1024, 8192), desc_mn(b4, O_X16 + k * 1024, 8192), ID_MM, k > 0);
        }
        umma_commit(&sm.bar_mma);
    }
    // operands of the delta phase that come from global memory: issue the loads now, behind the gamma product
    const int dt = tid >> 3, di0 = 8 * (tid & 7);
    const uint4 g_dy = __ldg(reinterpret_cast<const uint4*>(p.dy + row0 + (size_t)dt * rstride + di0));
    const uint4 g_v = __ldg(reinterpret_cast<const uint4*>(p.v + row0 + (size_t)dt * rstride + di0));
    const float4 g_u0 = __ldg(reinterpret_cast<const float4*>(p.sa + row0 + (size_t)dt * rstride + di0));
    const float4 g_u1 = __ldg(reinterpret_cast<const float4*>(p.sa + row0 + (size_t)dt * rstride + di0) + 1);
    float4 g_s0 = make_float4(0.f, 0.f, 0.f, 0.f), g_s1 = g_s0;
    if (c > 0) {  // S_0: checkpoint memory [j][i] holds S_ij (wkv7_cuda.cu:44-50) = K-major (N = j, K = i)
        const float4* ss = reinterpret_cast<const float4*>(p.s + ((((size_t)bb * H + hh) * nch * p.ck_per_chunk + (size_t)c * p.ck_per_chunk - 1) * N + dt) * N + di0);
        g_s0 = __ldg(ss);
        g_s1 = __ldg(ss + 1);
    }
    mma_wait();
    // ================= delta operands: [dR;dY] (K-major), dR (MN), [U;V] (K-major), S_0 =================
    if (r < 64) {
        uint32_t v[16];
        tmem_ld16(tm_row + C_R + 16 * cs, v);
#pragma unroll
        for (int c4 = 0; c4 < 4; c4++) {
            const float4 x = rt32(f4(v + 4 * c4));
            const int ch = 4 * (cs & 1) + c4;
            *reinterpret_cast<float4*>(ry + (cs >> 1) * 16384 + r * 128 + ((ch ^ (r & 7)) << 4)) = x;
            *reinterpret_cast<float4*>(sm.x16 + (cs >> 1) * 8192 + sw32_off(r, 16 * (cs & 1) + 4 * c4)) = x;
        }
    }
    {
        const int t = dt, i0 = di0;
        const int ch = (i0 & 31) >> 2, at = i0 >> 5;
        auto put8 = [&](uint8_t* base, int row, float4 lo, float4 hi) {
            *reinterpret_cast<float4*>(base + at * 16384 + row * 128 + ((ch ^ (row & 7)) << 4)) = lo;
            *reinterpret_cast<float4*>(base + at * 16384 + row * 128 + (((ch + 1) ^ (row & 7)) << 4)) = hi;
        };
        auto bf8 = [&](const uint4 u, float4& lo, float4& hi) {
            lo = make_float4(bf16lo_to_f32(u.x), bf16hi_to_f32(u.x), bf16lo_to_f32(u.y), bf16hi_to_f32(u.y));
            hi = make_float4(bf16lo_to_f32(u.z), bf16hi_to_f32(u.z), bf16lo_to_f32(u.w), bf16hi_to_f32(u.w));
        };
        float4 lo, hi;
        bf8(g_dy, lo, hi);
        put8(ry, 64 + t, lo, hi);
        bf8(g_v, lo, hi);
        put8(uv, 64 + t, lo, hi);
        put8(uv, t, rt32(g_u0), rt32(g_u1));
        *reinterpret_cast<float4*>(sm.s0 + at * 8192 + t * 128 + ((ch ^ (t & 7)) << 4)) = rt32(g_s0);
        *reinterpret_cast<float4*>(sm.s0 + at * 8192 + t * 128 + (((ch + 1) ^ (t & 7)) << 4)) = rt32(g_s1);
    }
    operands_ready();
    // ================= delta: dA = [dR;dY][U;V]^T ; dAt1 / dQt1 = [dR;dY] S_0 ; [dBt1;dKt1] = [U;V] dZ ; dV += A_ak^T dR =====
    if (issuer) {
        tc_fence_after();
        if (warp == 0) {
#pragma unroll
            for (int k = 0; k < 8; k++)
                umma_tf32(tmem + C_DA, desc_km(b4, O_Z1 + (k >> 2) * 16384 + (k & 3) * 32), desc_km(b4, O_Z2 + (k >> 2) * 16384 + (k & 3) * 32),
                          ID_KK_128, k > 0);
        } else if (warp == 1) {  // dA^T = [U;V] [dR;dY]^T (the transposed operand of dA^T [At;Qt] comes from a product, not a transpose)
#pragma unroll
            for (int k = 0; k < 8; k++)
                umma_tf32(tmem + C_DAT, desc_km(b4, O_Z2 + (k >> 2) * 16384 + (k & 3) * 32), desc_km(b4, O_Z1 + (k >> 2) * 16384 + (k & 3) * 32),
                          ID_KK_128, k > 0);
        } else if (warp == 2) {
#pragma unroll
            for (int k = 0; k < 8; k++)  // [dAt;dQt] = [dR;dY] S_0
                umma_tf32(tmem + C_AQ, desc_km(b4, O_Z1 + (k >> 2) * 16384 + (k & 3) * 32), desc_km(b4, O_S0 + (k >> 2) * 8192 + (k & 3) * 32), ID_KK, k > 0);
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++)  // dV += A_ak^T dR
                umma_tf32(tmem + C_ACC1, desc_mn(b4, O_SA + k * 1024, 8192), desc_mn(b4, O_X16 + k * 1024, 8192), ID_MM, 1);
#pragma unroll
            for (int k = 0; k < 8; k++)  // [dBt;dKt] = [U;V] dZ
                umma_tf32(tmem + C_BK, desc_km(b4, O_Z2 + (k >> 2) * 16384 + (k & 3) * 32), desc_mn(b4, O_DZ + k * 1024, 8192), ID_KM, k > 0);
        }
        umma_commit(&sm.bar_mma);
    }
    mma_wait();
    // ================= epsilon: dA and dA^T masked in place in TMEM, then fed back as A operands =================
    {
        uint32_t v[32];
        const int t = r & 63;
        const bool qrow = r >= 64;
        tmem_ld32(tm_row + C_DA + 32 * cs, v);  // dA[row r][s'], s' = 32cs + e: columns 0-63 vs U (A_ab / A_qb), 64-127 vs V
#pragma unroll
        for (int e = 0; e < 32; e++) {
            const int s_ = 32 * (cs & 1) + e;
            const bool keep = qrow ? (s_ <= t) : (s_ < t);
            v[e] = keep ? __float_as_uint(rt32(__uint_as_float(v[e]))) : 0u;
        }
        tmem_st32(tm_row + C_DA + 32 * cs, v);
        tmem_ld32(tm_row + C_DAT + 32 * cs, v);  // dA^T[row r = s'][column = a row t (0-63) or q row t (64-127)], s = r & 63
#pragma unroll
        for (int e = 0; e < 32; e++) {
            const int tc = 32 * (cs & 1) + e;
            const bool keep = (cs >= 2) ? (t <= tc) : (t < tc);
            v[e] = keep ? __float_as_uint(rt32(__uint_as_float(v[e]))) : 0u;
        }
        tmem_st32(tm_row + C_DAT + 32 * cs, v);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    if (issuer) {
        tc_fence_after();
        // (one issuing thread per accumulator: splitting a K range over two threads that accumulate into the same TMEM
        //  columns passed a micro-benchmark but produced wrong gradients here)
        if (warp == 0) {  // [dAt;dQt] += dA [Bt;Kt]
#pragma unroll
            for (int k = 0; k < 16; k++) umma_tf32_ts(tmem + C_AQ, tmem + C_DA + 8 * k, desc_mn(b4, O_BK + k * 1024, 16384), ID_KM, 1);
        } else if (warp == 1) {  // [dBt;dKt] += dA^T [At;Qt]
#pragma unroll
            for (int k = 0; k < 16; k++) umma_tf32_ts(tmem + C_BK, tmem + C_DAT + 8 * k, desc_mn(b4, O_AQ + k * 1024, 16384), ID_KM, 1);
        }
        umma_commit(&sm.bar_mma);
    }
    stamp();
    // inputs of the element-wise epilogue: loads issued now, consumed after the products above have finished
    const int et = r & 63, ej0 = 16 * cs;
    const size_t ego = row0 + (size_t)et * rstride + ej0;
    const float* egrow = p.gws + chunk_id * (L * N) + et * N + ej0;
    float4 eG[4], eGm[4];
    uint4 ein[2][2];
#pragma unroll
    for (int c4 = 0; c4 < 4; c4++) {
        eG[c4] = *reinterpret_cast<const float4*>(egrow + 4 * c4);
        eGm[c4] = (r < 64 && et > 0) ? *reinterpret_cast<const float4*>(egrow - N + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    {   // rows 0-63 finish da, db (need a, b); rows 64-127 finish dq, dk, dv (need q, k)
        const uint16_t* src0 = (r >= 64) ? p.q : p.a;
        const uint16_t* src1 = (r >= 64) ? p.k : p.b;
        ein[0][0] = __ldg(reinterpret_cast<const uint4*>(src0 + ego));
        ein[0][1] = __ldg(reinterpret_cast<const uint4*>(src0 + ego) + 1);
        ein[1][0] = __ldg(reinterpret_cast<const uint4*>(src1 + ego));
        ein[1][1] = __ldg(reinterpret_cast<const uint4*>(src1 + ego) + 1);
    }
    mma_wait();
    // ================= epilogue =================
    stamp();
    // three [64][64] fp32 scratch arrays with a row pitch of 65 floats: written row-wise by lanes that differ in t
    // (pitch 64 would put a whole warp on one bank), read column-wise by lanes that differ in j
    constexpr int EP = 65;
    float* const kk_s = reinterpret_cast<float*>(sm.z1);              // (db b)[t][j]
    float* const p1_s = reinterpret_cast<float*>(sm.z1) + 64 * EP;    // (dq q - dk k)[t][j]
    float* const p2_s = reinterpret_cast<float*>(sm.z1) + 128 * EP;   // (da a)[t][j]   (runs on into z2)
    {
        const int t = et, j0 = ej0;
        const size_t go = ego;
        float G[16], Gm[16];
#pragma unroll
        for (int c4 = 0; c4 < 4; c4++) {
            G[4 * c4] = eG[c4].x; G[4 * c4 + 1] = eG[c4].y; G[4 * c4 + 2] = eG[c4].z; G[4 * c4 + 3] = eG[c4].w;
            Gm[4 * c4] = eGm[c4].x; Gm[4 * c4 + 1] = eGm[c4].y; Gm[4 * c4 + 2] = eGm[c4].z; Gm[4 * c4 + 3] = eGm[c4].w;
        }
        auto un16 = [&](const uint4 (&u)[2], float (&o)[16]) {
            const uint32_t w[8] = {u[0].x, u[0].y, u[0].z, u[0].w, u[1].x, u[1].y, u[1].z, u[1].w};
#pragma unroll
            for (int e = 0; e < 8; e++) {
                o[2 * e] = bf16lo_to_f32(w[e]);
                o[2 * e + 1] = bf16hi_to_f32(w[e]);
            }
        };
        auto st16 = [&](uint16_t* ptr, const float (&o)[16]) {
            uint4 u0, u1;
            u0.x = pack_bf16x2(o[0], o[1]); u0.y = pack_bf16x2(o[2], o[3]); u0.z = pack_bf16x2(o[4], o[5]); u0.w = pack_bf16x2(o[6], o[7]);
            u1.x = pack_bf16x2(o[8], o[9]); u1.y = pack_bf16x2(o[10], o[11]); u1.z = pack_bf16x2(o[12], o[13]); u1.w = pack_bf16x2(o[14], o[15]);
            *reinterpret_cast<uint4*>(ptr) = u0;
            *(reinterpret_cast<uint4*>(ptr) + 1) = u1;
        };
        if (r >= 64) {  // dq, dk, dv
            uint32_t vq[16], vk[16], vv[16];
            tmem_ld16_nowait(tm_row + C_AQ + j0, vq);
            tmem_ld16_nowait(tm_row + C_BK + j0, vk);
            tmem_ld16_nowait(tm_row + C_ACC1 + j0, vv);
            tmem_ld_wait();
            float qin[16], kin[16], dq[16], dk[16], dv[16];
            un16(ein[0], qin);
            un16(ein[1], kin);
#pragma unroll
            for (int e = 0; e < 16; e++) {
                dq[e] = __uint_as_float(vq[e]) * __expf(G[e]);
                dk[e] = __uint_as_float(vk[e]) * __expf(-G[e]);
                dv[e] = __uint_as_float(vv[e]);
                p1_s[t * EP + j0 + e] = dq[e] * qin[e] - dk[e] * kin[e];
            }
            st16(p.dq + go, dq);
            st16(p.dk + go, dk);
            st16(p.dv + go, dv);
        } else {  // da, db
            uint32_t va[16], vb[16];
            tmem_ld16_nowait(tm_row + C_AQ + j0, va);
            tmem_ld16_nowait(tm_row + C_BK + j0, vb);
            tmem_ld_wait();
            float ain[16], bin[16], da[16], db[16];
            un16(ein[0], ain);
            un16(ein[1], bin);
#pragma unroll
            for (int e = 0; e < 16; e++) {
                da[e] = __uint_as_float(va[e]) * __expf(Gm[e]);
                db[e] = __uint_as_float(vb[e]) * __expf(-G[e]);
                kk_s[t * EP + j0 + e] = db[e] * bin[e];
                p2_s[t * EP + j0 + e] = da[e] * ain[e];
            }
            st16(p.da + go, da);
            st16(p.db + go, db);
        }
    }
    stamp();
    // d/dG_L through S_L = Z diag(e^{G_L}): sum_i dS_L[i][j] S_L[i][j]   (S_L: checkpoint after step 63 of this chunk)
    {
        const int j = tid & 63, ig = tid >> 6;
        float acc = 0.f;
        if (c + 1 < nch) {
            const float* ck = p.s + ((((size_t)bb * H + hh) * nch * p.ck_per_chunk + (size_t)(c + 1) * p.ck_per_chunk - 1) * N + j) * N + 8 * ig;
            const float4 s0 = __ldg(reinterpret_cast<const float4*>(ck)), s1 = __ldg(reinterpret_cast<const float4*>(ck) + 1);
            const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
            const float* dsl = p.ds + chunk_id * (N * N) + (size_t)(8 * ig) * N + j;
#pragma unroll
            for (int e = 0; e < 8; e++) acc = fmaf(__ldg(dsl + (size_t)e * N), sv[e], acc);
        }
        sm.part[ig][j] = acc;
    }
    __syncthreads();
    if (tid < N) {
        float x = 0.f;
#pragma unroll
        for (int k = 0; k < 8; k++) x += sm.part[k][tid];
        sm.gl[tid] = x;
    }
    __syncthreads();
    stamp();
    // dG_t = p1_t - kk_t + p2_{t+1} (+ gl at t = 63); dg = suffix sum over t; dw = dg * (-e^w)
    {
        const int j = tid & 63, rg = tid >> 6;
        float dG[8];
        float run = 0.f;
#pragma unroll
        for (int k = 7; k >= 0; k--) {
            const int t = 8 * rg + k;
            const float nxt = (t == L - 1) ? sm.gl[j] : p2_s[(t + 1) * EP + j];
            run += p1_s[t * EP + j] - kk_s[t * EP + j] + nxt;
            dG[k] = run;  // suffix sum inside the row group
        }
        __syncthreads();  // everyone has read gl / part before part is reused
        sm.part[rg][j] = run;
        __syncthreads();
        float off = 0.f;
#pragma unroll
        for (int k = 1; k < 8; k++) off += (k > rg) ? sm.part[k][j] : 0.f;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int t = 8 * rg + k;
            const size_t go = row0 + (size_t)t * rstride + j;
            const float g = -__expf(bf16lo_to_f32((uint32_t)__ldg(p.w + go)));
            p.dw[go] = f32_to_bf16_bits((dG[k] + off) * g);
        }
    }
    stamp();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

}  // namespace vrwkv
