// wkv7_chunk_fwd.cuh — WKV7 forward evaluated chunk by chunk on the tcgen05 tensor cores (TF32 operands, fp32 TMEM
// accumulators).  Same operator contract as the step-by-step kernels (wkv7_fwd2.cuh) and the reference forward_kernel
// (VisualRWKV-v7/v7.00/cuda/wkv7_cuda.cu:10-52): y (bf16), sa (fp32, every step), s (fp32, transposed state every 16
// steps); the chunk-wise restatement kept with the tests states the algebra and is checked against the step-by-step oracle.
//
// One 512-thread CTA per (batch, head) walks the sequence in chunks of L = 64 steps.  With G_t the running sum of
// -exp(w) inside the chunk and  At = a exp(G_{t-1}), Qt = q exp(G_t), Kt = k exp(-G_t), Bt = b exp(-G_t):
//   scores  = [At;Qt] [Bt;Kt]^T      (128x128x64)  -> A_ab, A_ak (strictly lower), A_qb, A_qk (lower triangular)
//   ACC     = [A_ak;A_qk] V          (128x64x64)   rows 0-63: AV
//   Tinv    = (I - A_ab)^-1          fp32 on the CUDA cores (chunk_tri_inverse: 16x16 blocks + two coupling levels)
//   W       = [A_ab;A_qb] Tinv       (128x64x64)   stays in TMEM and is the A operand of the next two products
//   CORR    = W At ;  ACC += W AV     (= [A_ab;A_qb][Ah | Uh])   => [At;Qt] + CORR = [Ah;Qp],  ACC = [Uh; Y_intra]
//   ACC    += [Ah;Qp] S_0^T                                    => ACC = [U; Y]  (rows of U are the sa_t)
//   D_g     = U_g^T Bt_g + V_g^T Kt_g for the four 16-step groups g; S_{16(g+1)} = (S_0 + D_0 + .. + D_g) diag(exp(G))
// Operand layouts: K-major operands use SWIZZLE_128B rows of 32 tf32; [Bt;Kt], [U;V] and [At|AV] -> [Ah|Uh] are
// stored row-major [step][channel] in the SWIZZLE_128B_BASE32B layout and consumed MN-major, so nothing is transposed
// by hand.  Each product batch is issued by lane 0 of warps 0 and 1 (two accumulators in flight, see the comment at the
// issue sites) and followed by commits that all threads wait for: the phases are sequential, the tensor-core time per
// chunk is a few thousand cycles, the rest is operand preparation on the CUDA cores (DESIGN.md 2.2b has the breakdown).
// Domain: exp(+-G) must stay finite, i.e. sum over a chunk of exp(w) < ~85 — guaranteed by RWKV-7's
// w = -softplus(.) - 0.5 (exp(w) <= 0.607, model.py:176); the host dispatcher keeps the step-by-step kernel for
// callers that cannot promise that.
#pragma once
#include "wkv7_chunk_common.cuh"

namespace vrwkv {

struct alignas(1024) Wkv7ChunkSmem {
    uint8_t in[6 * CK_L * WKV_N * 2];  // TMA tiles w,q,k,v,a,b [64][64] bf16; later RMN (32 KB) and TINV (16 KB)
    uint8_t aq[32768];                 // A operand [At;Qt] -> [Ah;Qp]: 2 k-atoms x 128 rows x 128 B (also TINV's tail)
    uint8_t bk[32768];                 // B operand [Bt;Kt] of the score product (K-major); then `sc`
    uint8_t bk2[32768];                // [Bt;Kt], MN-major (rows = step: Bt 0-63, Kt 64-127; 2 channel blocks)
    uint8_t uv[32768];                 // [U;V], MN-major (rows = step: U 0-63, V 64-127; 2 channel blocks)
    uint8_t sb[16384];                 // B operand S_0 [i][j] (tf32): 2 k-atoms x 64 rows
    uint8_t aab[16384];                // A_ab fp32 [t][s], 16-byte chunks XOR-swizzled by (t & 7)
    float esc[32 * 32];                // coupling-block scratch of the inverse
    float part[8][WKV_N];              // per row-group decay sums
    float echk[4][WKV_N];              // exp(G_t) at t = 15, 31, 47, 63
    float eck[4][WKV_N];               // the same for the previous chunk (its checkpoints are written one chunk late)
    uint64_t bar_in, bar_mma, bar_w;
    uint32_t tmem_base;
};

// CHUNK_CK: s holds one (transposed) state per 64-step chunk, s[b][h][c] = S after step 64c+63 — all the chunk-local
// backward reads — instead of the reference's four 16-step checkpoints per chunk (3/4 of the forward's DRAM writes).
template <bool CHUNK_CK>
__global__ void __launch_bounds__(CK_THREADS, 1)
wkv7_chunk_fwd_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_q,
                      const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                      const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                      const Wkv7FwdArgs p) {
    constexpr int N = WKV_N, L = CK_L;
    extern __shared__ __align__(1024) uint8_t chunk_smem_bytes[];
    Wkv7ChunkSmem& sm = *reinterpret_cast<Wkv7ChunkSmem*>(chunk_smem_bytes);
    uint8_t* const rmn = sm.in;           // [At|AV] -> [Ah|Uh], MN-major: 4 channel blocks x 64 k-lines x 128 B
    uint8_t* const tinv = sm.in + 32768;  // Tinv [t][s] row-major, MN-major B operand: 2 column blocks x 64 k-lines
    uint8_t* const sc = sm.bk;            // A operand [A_ak;A_qk], then [A_ab;A_qb]

    const int hh = blockIdx.x, bb = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int qd = warp & 3, cs = warp >> 2;  // TMEM lane quadrant / column slice of this warp
    const int r = 32 * qd + lane;             // accumulator row (TMEM lane) of this thread
    const int t_r = r & 63;
    const int T = p.T, H = p.H;
    const int nch = T / L;

    if (tid == 0) {
        mbar_init(&sm.bar_in, 1);
        mbar_init(&sm.bar_mma, 2);  // two issuing threads (lane 0 of warps 0 and 1) commit every batch
        mbar_init(&sm.bar_w, 1);    // W = [A_ab;A_qb] Tinv complete (waited for by the two issuers only)
        fence_mbar_init();
        tma_prefetch_desc(&tm_w); tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k);
        tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_a); tma_prefetch_desc(&tm_b);
    }
    __syncwarp();
    if (warp == 0) tmem_alloc<512>(&sm.tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    const uint32_t tm_row = tmem + ((uint32_t)(32 * qd) << 16);
    constexpr uint32_t C_SC = 0, C_UY = 128, C_CORR = 192, C_TX = 320;
    constexpr uint32_t C_SC2 = 384;  // second K-half of the scores (free until the state products at the end of the chunk)
    // partial state sums of the four 16-step groups reuse the score / TX columns (both are dead by then)
    constexpr uint32_t C_R0 = 0, C_R1 = 64, C_R2 = 320, C_R3 = 384;

    auto issue_loads = [&](int c) {
        mbar_arrive_expect_tx(&sm.bar_in, 6 * L * N * 2);
        const int x0 = hh * N, y0 = bb * T + c * L;
        tma_load_2d(sm.in + 0 * 8192, &tm_w, x0, y0, &sm.bar_in);
        tma_load_2d(sm.in + 1 * 8192, &tm_q, x0, y0, &sm.bar_in);
        tma_load_2d(sm.in + 2 * 8192, &tm_k, x0, y0, &sm.bar_in);
        tma_load_2d(sm.in + 3 * 8192, &tm_v, x0, y0, &sm.bar_in);
        tma_load_2d(sm.in + 4 * 8192, &tm_a, x0, y0, &sm.bar_in);
        tma_load_2d(sm.in + 5 * 8192, &tm_b, x0, y0, &sm.bar_in);
    };
    if (tid == 0) issue_loads(0);
    __syncwarp();
#ifdef VRWKV_PHASE_STAMPS   // development builds only (VRWKV_PHASE_STAMPS=1)
    float* const dbg = (hh == 0 && bb == 0) ? g_chunk_dbg : nullptr;
#endif

    // ---- initial state: fp32 in registers (thread r < 64 holds S[r][16cs .. 16cs+15]) + tf32 image as B operand ----
    float Sprev[16];
#pragma unroll
    for (int e = 0; e < 16; e++) Sprev[e] = 0.f;
    if (r < N) {
        if (p.state_in) {
            const float4* src = reinterpret_cast<const float4*>(p.state_in + (((size_t)bb * H + hh) * N + r) * N + 16 * cs);
#pragma unroll
            for (int c4 = 0; c4 < 4; c4++) {
                const float4 x = __ldg(src + c4);
                Sprev[4 * c4] = x.x; Sprev[4 * c4 + 1] = x.y; Sprev[4 * c4 + 2] = x.z; Sprev[4 * c4 + 3] = x.w;
            }
        }
#pragma unroll
        for (int c4 = 0; c4 < 4; c4++) {
            const int ch = 4 * (cs & 1) + c4;
            *reinterpret_cast<float4*>(sm.sb + (cs >> 1) * 8192 + r * 128 + ((ch ^ (r & 7)) << 4)) =
                rt32(make_float4(Sprev[4 * c4], Sprev[4 * c4 + 1], Sprev[4 * c4 + 2], Sprev[4 * c4 + 3]));
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

#ifdef VRWKV_PHASE_STAMPS
    long long tstamp0 = 0;
    int tsi = 0;
    auto stamp = [&](int c) {
        if (dbg && c == 1 && tid == 0) {
            const long long now = clock64();
            if (tsi == 0) tstamp0 = now;
            dbg[2048 + tsi++] = (float)(now - tstamp0);
        }
    };
#else
    auto stamp = [](int) {};
#endif
    uint32_t mph = 0;  // phase of bar_mma (every commit is waited for before the next one is issued)
    auto mma_wait = [&](int c) {
        stamp(c);
        mbar_wait(&sm.bar_mma, mph & 1);
        mph++;
        tc_fence_after();
        __syncwarp();
        stamp(c);
    };
    auto operands_ready = [&]() {  // generic-proxy writes -> visible to the tensor core, then CTA barrier
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
    };
    // A single thread issues one tcgen05.mma per ~120 cycles whatever its size (scripts/ubench_mma_issue.cu); two threads
    // feeding different accumulators reach one per ~62 cycles.  Every product batch is therefore split over two issuers,
    // by accumulator where there are two, else by K range into two accumulators that the following phase adds.
    const uint32_t b4 = smem_u32(&sm) >> 4;
    const uint32_t O_IN = (uint32_t)(sm.in - (uint8_t*)&sm), O_AQ = (uint32_t)(sm.aq - (uint8_t*)&sm), O_BK = (uint32_t)(sm.bk - (uint8_t*)&sm),
                   O_BK2 = (uint32_t)(sm.bk2 - (uint8_t*)&sm), O_UV = (uint32_t)(sm.uv - (uint8_t*)&sm), O_SB = (uint32_t)(sm.sb - (uint8_t*)&sm);
    const uint32_t O_RMN = O_IN, O_TINV = O_IN + 32768, O_SC = O_BK;
    const bool issuer = lane == 0 && warp < 2;
    const int iw = warp;  // issuer index (valid when `issuer`)
    constexpr uint32_t ID_128x128 = umma_idesc_tf32(128, 128), ID_128x64 = umma_idesc_tf32(128, 64),
                       ID_B_MN_64 = umma_idesc_tf32(128, 64, 0, 1),
                       ID_ST = umma_idesc_tf32(128, 64, 1, 1);

    // checkpoints of chunk c-1: written at the start of chunk c so that the stores drain behind P1 instead of in
    // front of the end-of-chunk fence; the four partial-state regions stay valid until the next score product
    float Sck[16];  // state at the start of the chunk whose checkpoints are still to be written
#pragma unroll
    for (int e = 0; e < 16; e++) Sck[e] = 0.f;
    auto write_checkpoints = [&](int c) {
        if (p.s == nullptr || r >= N) return;
        uint32_t v0[16], v1[16], v2[16], v3[16];
        tmem_ld16_nowait(tm_row + C_R0 + 16 * cs, v0);
        tmem_ld16_nowait(tm_row + C_R1 + 16 * cs, v1);
        tmem_ld16_nowait(tm_row + C_R2 + 16 * cs, v2);
        tmem_ld16_nowait(tm_row + C_R3 + 16 * cs, v3);
        tmem_ld_wait();
        float* ck = p.s + ((((size_t)bb * H + hh) * (T / WKV_TC) + (size_t)c * 4) * N + 16 * cs) * N + r;
#pragma unroll
        for (int e = 0; e < 16; e++) {
            float acc = Sck[e] + __uint_as_float(v0[e]);
            ck[(size_t)e * N] = acc * sm.eck[0][16 * cs + e];  // transposed checkpoint [j][i] (wkv7_cuda.cu:44-50)
            acc += __uint_as_float(v1[e]);
            ck[(size_t)(N + e) * N] = acc * sm.eck[1][16 * cs + e];
            acc += __uint_as_float(v2[e]);
            ck[(size_t)(2 * N + e) * N] = acc * sm.eck[2][16 * cs + e];
            acc += __uint_as_float(v3[e]);
            ck[(size_t)(3 * N + e) * N] = acc * sm.eck[3][16 * cs + e];
        }
    };

    for (int c = 0; c < nch; c++) {
        stamp(c);
        if (!CHUNK_CK && c > 0) write_checkpoints(c - 1);
        mbar_wait(&sm.bar_in, c & 1);
        stamp(c);
        // ================= P1: decay prefix sums and scaled operands (8 rows x 1 column per thread) =================
        {
            const int hf = warp & 1, rg = warp >> 1;
            const int j = 32 * hf + lane;
            const uint16_t* in16 = reinterpret_cast<const uint16_t*>(sm.in);
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
            uint8_t* const aq_a = sm.aq + hf * 16384;
            uint8_t* const bk_a = sm.bk + hf * 16384;
            uint8_t* const uv_a = sm.uv + hf * 16384;
            uint8_t* const bk2_a = sm.bk2 + hf * 16384;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int t = 8 * rg + k;
                G += g[k];
                const float E = __expf(G), F = __expf(-G);
                const float qv = bf16lo_to_f32(in16[1 * 4096 + t * N + j]), kv = bf16lo_to_f32(in16[2 * 4096 + t * N + j]);
                const float vv = bf16lo_to_f32(in16[3 * 4096 + t * N + j]), av_ = bf16lo_to_f32(in16[4 * 4096 + t * N + j]);
                const float bv = bf16lo_to_f32(in16[5 * 4096 + t * N + j]);
                const float bt = rt32(bv * F), kt = rt32(kv * F);
                *reinterpret_cast<float*>(aq_a + sw128_off(t, lane)) = rt32(av_ * Eprev);
                *reinterpret_cast<float*>(aq_a + sw128_off(64 + t, lane)) = rt32(qv * E);
                *reinterpret_cast<float*>(bk_a + sw128_off(t, lane)) = bt;
                *reinterpret_cast<float*>(bk_a + sw128_off(64 + t, lane)) = kt;
                *reinterpret_cast<float*>(bk2_a + sw32_off(t, lane)) = bt;
                *reinterpret_cast<float*>(bk2_a + sw32_off(64 + t, lane)) = kt;
                *reinterpret_cast<float*>(uv_a + sw32_off(64 + t, lane)) = vv;  // bf16 values are exact in tf32
                if ((t & 15) == 15) sm.echk[t >> 4][j] = E;
                if (t == L - 1 && G < -80.f) g_chunk_domain_err = 1;
                Eprev = E;
            }
        }
        operands_ready();
        // ================= scores =================
        if (issuer) {  // issuer 0: channels 0-31 -> C_SC, issuer 1: channels 32-63 -> C_SC2 (added in P2)
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; k++)
                umma_tf32(tmem + (iw ? C_SC2 : C_SC), desc_km(b4, O_AQ + iw * 16384 + k * 32), desc_km(b4, O_BK + iw * 16384 + k * 32), ID_128x128, k > 0);
            umma_commit(&sm.bar_mma);
        }
        mma_wait(c);
        // ================= P2: masks; [A_ak;A_qk] operand; A_ab in fp32 for the inverse =================
        *reinterpret_cast<float4*>(tinv + tid * 32) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(tinv + tid * 32 + 16) = make_float4(0.f, 0.f, 0.f, 0.f);
        float sc0[32];  // warps with cs < 2 keep their slice of [A_ab;A_qb] until the operand buffer is free again
        {
            uint32_t v[32];
            const bool incl = r >= 64;  // q rows keep the diagonal
            uint32_t v2[32];
            tmem_ld32_nowait(tm_row + C_SC + 32 * cs, v);
            tmem_ld32_nowait(tm_row + C_SC2 + 32 * cs, v2);
            tmem_ld_wait();
            const int sbase = 32 * (cs & 1);
#pragma unroll
            for (int e = 0; e < 32; e++) {
                const int s = sbase + e;
                const bool keep = incl ? (s <= t_r) : (s < t_r);
                sc0[e] = keep ? __uint_as_float(v[e]) + __uint_as_float(v2[e]) : 0.f;
            }
            if (cs < 2) {
                if (r < 64) {
#pragma unroll
                    for (int cc = 0; cc < 8; cc++)
                        *reinterpret_cast<float4*>(sm.aab + r * 256 + (((8 * cs + cc) ^ (r & 7)) << 4)) =
                            make_float4(sc0[4 * cc], sc0[4 * cc + 1], sc0[4 * cc + 2], sc0[4 * cc + 3]);
                }
            } else {
#pragma unroll
                for (int cc = 0; cc < 8; cc++)
                    *reinterpret_cast<float4*>(sc + (cs & 1) * 16384 + r * 128 + ((cc ^ (r & 7)) << 4)) =
                        rt32(make_float4(sc0[4 * cc], sc0[4 * cc + 1], sc0[4 * cc + 2], sc0[4 * cc + 3]));
            }
        }
        operands_ready();
        // ================= ACC = [A_ak;A_qk] V (runs while the inverse is being computed) =================
        if (issuer) {  // (hidden behind the inverse: one issuer is enough)
            tc_fence_after();
            if (iw == 0) {
#pragma unroll
                for (int k = 0; k < 8; k++)
                    umma_tf32(tmem + C_UY, desc_km(b4, O_SC + (k >> 2) * 16384 + (k & 3) * 32), desc_mn(b4, O_UV + 64 * 128 + k * 1024, 16384), ID_B_MN_64, k > 0);
            }
            umma_commit(&sm.bar_mma);
        }
        // ================= At (K-major rows 0-63 of aq) -> rmn blocks 0,1 (MN-major, k-line = step) =================
        {
            const int t = tid >> 3, q8 = tid & 7;  // two 16-byte chunks per thread
#pragma unroll
            for (int x = 0; x < 2; x++) {
                const int ch = 2 * q8 + x;  // chunk 0..15 of the 64-float row
                const float4 val = *reinterpret_cast<const float4*>(sm.aq + (ch >> 3) * 16384 + t * 128 + (((ch & 7) ^ (t & 7)) << 4));
                *reinterpret_cast<float4*>(rmn + (ch >> 3) * 8192 + sw32_off(t, 4 * (ch & 7))) = val;
            }
        }
        // ================= Tinv = (I - A_ab)^-1 =================
        chunk_tri_inverse(sm.aab, sm.esc, tid, [&](int t, int s) { return tinv + (s >> 5) * 8192 + sw32_off(t, s & 31); });
        mma_wait(c);
        // ================= P3: AV -> rmn blocks 2,3; [A_ab;A_qb] operand =================
        if (r < 64) {
            uint32_t v[16];
            tmem_ld16(tm_row + C_UY + 16 * cs, v);
#pragma unroll
            for (int c4 = 0; c4 < 4; c4++)
                *reinterpret_cast<float4*>(rmn + (2 + (cs >> 1)) * 8192 + sw32_off(r, 16 * (cs & 1) + 4 * c4)) =
                    rt32(make_float4(__uint_as_float(v[4 * c4]), __uint_as_float(v[4 * c4 + 1]), __uint_as_float(v[4 * c4 + 2]),
                                     __uint_as_float(v[4 * c4 + 3])));
        }
        if (cs < 2) {
#pragma unroll
            for (int cc = 0; cc < 8; cc++)
                *reinterpret_cast<float4*>(sc + cs * 16384 + r * 128 + ((cc ^ (r & 7)) << 4)) =
                    rt32(make_float4(sc0[4 * cc], sc0[4 * cc + 1], sc0[4 * cc + 2], sc0[4 * cc + 3]));
        }
        operands_ready();
        // ================= W = [A_ab;A_qb] Tinv ;  CORR = W At ;  ACC += W AV   (W never leaves TMEM) =================
        // (= [A_ab;A_qb] [Ah | Uh] with [Ah | Uh] = Tinv [At | AV]; the second product reads its A operand from the
        //  accumulator columns of the first, scripts/ubench_mma_modes.cu)
        if (issuer) {
            tc_fence_after();
            if (iw == 0) {
#pragma unroll
                for (int k = 0; k < 8; k++)
                    umma_tf32(tmem + C_TX, desc_km(b4, O_SC + (k >> 2) * 16384 + (k & 3) * 32), desc_mn(b4, O_TINV + k * 1024, 8192), ID_B_MN_64, k > 0);
                umma_commit(&sm.bar_w);
            }
            mbar_wait(&sm.bar_w, c & 1);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 8; k++)
                umma_tf32_ts(tmem + (iw ? C_UY : C_CORR), tmem + C_TX + 8 * k, desc_mn(b4, O_RMN + iw * 2 * 8192 + k * 1024, 8192), ID_B_MN_64,
                             iw ? 1 : (k > 0));
            umma_commit(&sm.bar_mma);
        }
        mma_wait(c);
        if (tid == 0 && c + 1 < nch) issue_loads(c + 1);  // rmn / tinv are dead: the input buffer is free again
        __syncwarp();
        // ================= P5: [At;Qt] += CORR  ->  [Ah;Qp] =================
        {
            uint32_t v[16];
            tmem_ld16(tm_row + C_CORR + 16 * cs, v);
#pragma unroll
            for (int c4 = 0; c4 < 4; c4++) {
                const int ch = 4 * (cs & 1) + c4;
                float4* ptr = reinterpret_cast<float4*>(sm.aq + (cs >> 1) * 16384 + r * 128 + ((ch ^ (r & 7)) << 4));
                float4 o = *ptr;
                o.x = rt32(o.x + __uint_as_float(v[4 * c4]));
                o.y = rt32(o.y + __uint_as_float(v[4 * c4 + 1]));
                o.z = rt32(o.z + __uint_as_float(v[4 * c4 + 2]));
                o.w = rt32(o.w + __uint_as_float(v[4 * c4 + 3]));
                *ptr = o;
            }
        }
        operands_ready();
        // ================= ACC += [Ah;Qp] S_0^T  ->  [U;Y] =================
        if (issuer) {  // issuer 0: j 0-31 -> ACC, issuer 1: j 32-63 -> the CORR columns (consumed by P5); added in P6
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; k++)
                umma_tf32(tmem + (iw ? C_CORR : C_UY), desc_km(b4, O_AQ + iw * 16384 + k * 32), desc_km(b4, O_SB + iw * 8192 + k * 32), ID_128x64,
                          iw ? (k > 0) : 1);
            umma_commit(&sm.bar_mma);
        }
        mma_wait(c);
        // ================= P6: U operand; then D_g = U_g^T Bt_g + V_g^T Kt_g; sa / y stores behind the barrier ====
        uint32_t uy[16];
        {
            uint32_t u2[16];
            tmem_ld16_nowait(tm_row + C_UY + 16 * cs, uy);
            tmem_ld16_nowait(tm_row + C_CORR + 16 * cs, u2);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; e++) uy[e] = __float_as_uint(__uint_as_float(uy[e]) + __uint_as_float(u2[e]));
        }
        if (r < 64) {
#pragma unroll
            for (int c4 = 0; c4 < 4; c4++)
                *reinterpret_cast<float4*>(sm.uv + (cs >> 1) * 16384 + sw32_off(r, 16 * (cs & 1) + 4 * c4)) =
                    rt32(make_float4(__uint_as_float(uy[4 * c4]), __uint_as_float(uy[4 * c4 + 1]), __uint_as_float(uy[4 * c4 + 2]),
                                     __uint_as_float(uy[4 * c4 + 3])));
        }
        operands_ready();
        if (issuer) {  // issuer 0: groups 0,1; issuer 1: groups 2,3
            tc_fence_after();
            constexpr uint32_t creg[4] = {C_R0, C_R1, C_R2, C_R3};
#pragma unroll
            for (int gg = 0; gg < 2; gg++)
#pragma unroll
                for (int part = 0; part < 2; part++)
#pragma unroll
                    for (int x = 0; x < 2; x++) {
                        const int g = 2 * iw + gg;
                        const uint32_t off = (uint32_t)(part * 64 + 16 * g + 8 * x) * 128;
                        umma_tf32(tmem + (iw ? creg[2 + gg] : creg[gg]), desc_mn(b4, O_UV + off, 16384), desc_mn(b4, O_BK2 + off, 16384), ID_ST, (part | x) != 0);
                    }
            umma_commit(&sm.bar_mma);
        }
        {
            const size_t row = (((size_t)bb * T + (size_t)c * L + t_r) * H + hh) * N + 16 * cs;
            if (r < 64) {
                if (p.sa) {
#pragma unroll
                    for (int c4 = 0; c4 < 4; c4++)
                        *reinterpret_cast<float4*>(p.sa + row + 4 * c4) =
                            make_float4(__uint_as_float(uy[4 * c4]), __uint_as_float(uy[4 * c4 + 1]), __uint_as_float(uy[4 * c4 + 2]),
                                        __uint_as_float(uy[4 * c4 + 3]));
                }
            } else {
                uint4 o0, o1;
                o0.x = pack_bf16x2(__uint_as_float(uy[0]), __uint_as_float(uy[1]));
                o0.y = pack_bf16x2(__uint_as_float(uy[2]), __uint_as_float(uy[3]));
                o0.z = pack_bf16x2(__uint_as_float(uy[4]), __uint_as_float(uy[5]));
                o0.w = pack_bf16x2(__uint_as_float(uy[6]), __uint_as_float(uy[7]));
                o1.x = pack_bf16x2(__uint_as_float(uy[8]), __uint_as_float(uy[9]));
                o1.y = pack_bf16x2(__uint_as_float(uy[10]), __uint_as_float(uy[11]));
                o1.z = pack_bf16x2(__uint_as_float(uy[12]), __uint_as_float(uy[13]));
                o1.w = pack_bf16x2(__uint_as_float(uy[14]), __uint_as_float(uy[15]));
                *reinterpret_cast<uint4*>(p.y + row) = o0;
                *reinterpret_cast<uint4*>(p.y + row + 8) = o1;
            }
        }
        mma_wait(c);
        // ================= new state: S_L = (S_0 + D_0 + D_1 + D_2 + D_3) diag(exp(G_L)) =================
        if (r < N) {
            uint32_t v0[16], v1[16], v2[16], v3[16];
            tmem_ld16_nowait(tm_row + C_R0 + 16 * cs, v0);
            tmem_ld16_nowait(tm_row + C_R1 + 16 * cs, v1);
            tmem_ld16_nowait(tm_row + C_R2 + 16 * cs, v2);
            tmem_ld16_nowait(tm_row + C_R3 + 16 * cs, v3);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; e++) {
                Sck[e] = Sprev[e];
                const float acc = ((Sprev[e] + __uint_as_float(v0[e])) + __uint_as_float(v1[e])) + __uint_as_float(v2[e]) +
                                  __uint_as_float(v3[e]);
                Sprev[e] = acc * sm.echk[3][16 * cs + e];
            }
#pragma unroll
            for (int c4 = 0; c4 < 4; c4++) {
                const int ch = 4 * (cs & 1) + c4;
                *reinterpret_cast<float4*>(sm.sb + (cs >> 1) * 8192 + r * 128 + ((ch ^ (r & 7)) << 4)) =
                    rt32(make_float4(Sprev[4 * c4], Sprev[4 * c4 + 1], Sprev[4 * c4 + 2], Sprev[4 * c4 + 3]));
            }
            if (CHUNK_CK && p.s) {
                float* ck = p.s + ((((size_t)bb * H + hh) * nch + (size_t)c) * N + 16 * cs) * N + r;
#pragma unroll
                for (int e = 0; e < 16; e++) ck[(size_t)e * N] = Sprev[e];  // transposed [j][i], like the 16-step checkpoints
            }
        }
        if (!CHUNK_CK && tid < 4 * N) sm.eck[tid >> 6][tid & 63] = sm.echk[tid >> 6][tid & 63];
        operands_ready();
    }
    if (!CHUNK_CK) write_checkpoints(nch - 1);
    if (p.state_out && r < N) {
        float4* dst = reinterpret_cast<float4*>(p.state_out + (((size_t)bb * H + hh) * N + r) * N + 16 * cs);
#pragma unroll
        for (int c4 = 0; c4 < 4; c4++) dst[c4] = make_float4(Sprev[4 * c4], Sprev[4 * c4 + 1], Sprev[4 * c4 + 2], Sprev[4 * c4 + 3]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

}  // namespace vrwkv
