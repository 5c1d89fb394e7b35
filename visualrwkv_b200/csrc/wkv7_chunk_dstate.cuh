// wkv7_chunk_dstate.cuh — reverse-time scan of dL/dS over 64-step chunks on the tcgen05 tensor cores.
//
// The WKV7 backward (reference backward_kernel, VisualRWKV-v7/v7.00/cuda/wkv7_cuda.cu:54-130) is one serial walk over
// T steps per (batch, head).  The state gradient obeys a linear recurrence that does not involve the state itself,
//     dS_{t-1} = dS_t diag(d_t) + (dS_t b_t) a_t^T + dy_t q_t^T        (:96, :105, :125)
// so its values at chunk boundaries can be produced ahead of time by a much shorter serial chain of matrix products;
// the step-by-step kernel (wkv7_bwd2.cuh) then runs on all 64-step segments of all heads concurrently, each started
// from its boundary value.  With the chunk quantities of wkv7_chunk_fwd.cuh (G, At, Qt, Bt, A_ab, A_qb) and dS_L the
// gradient at the end of a chunk:
//     dZ  = dS_L diag(exp(G_L))
//     dU  = Bt dZ^T + A_qb^T dY
//     dR  = (I - A_ab)^-T dU
//     dS_0 = dZ + dR^T At + dY^T Qt
// (the tests' chunk_backward restatement, checked against the fp64 adjoint of the step-by-step oracle).
// One 512-thread CTA per (batch, head); TF32 operands, fp32 accumulation in TMEM; the triangular inverse is computed in
// fp32 on the CUDA cores exactly as in the forward kernel.  Output: ds[b][h][c][i][j] = dL/dS at the end of chunk c
// for c = 0 .. T/64 - 2 (the last chunk ends the sequence: zero, not stored).
#pragma once
#include "wkv7_chunk_common.cuh"

namespace vrwkv {

struct Wkv7DstateArgs {
    int B, T, H;
    float* ds;  // [B][H][T/64][64][64]
};

struct alignas(1024) Wkv7DstateSmem {
    uint8_t in[5 * CK_L * WKV_N * 2];  // TMA tiles w,q,a,b,dy; later SQB (16 KB) and DU (16 KB)
    uint8_t aq[32768];                 // A operand [At;Qt] (K-major)
    uint8_t bt[16384];                 // Bt (K-major): B operand of the scores, A operand of Bt dZ^T (rows 64-127: what follows)
    uint8_t aq2[32768];                // [At;Qt] MN-major (k-line = step), 2 channel blocks
    uint8_t ry[32768];                 // [dR;dY] MN-major (k-line = step), 2 channel blocks
    uint8_t tinv[16384];               // Tinv [t][s] MN-major (k-line = t), 2 blocks of 32 s
    uint8_t dz[16384];                 // dZ [i][j] K-major
    uint8_t aab[16384];                // A_ab fp32 [t][s], chunk-swizzled
    // (M = 128 operands built from 64 real rows read on into the buffer that follows them: bt -> aq2, ry -> tinv/dz,
    //  tinv -> dz, sqb -> du; those accumulator rows are never used)
    float esc[32 * 32];
    float part[8][WKV_N];
    float el[WKV_N];
    uint64_t bar_in, bar_mma;
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(CK_THREADS, 1)
wkv7_chunk_dstate_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_q,
                         const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                         const __grid_constant__ CUtensorMap tm_dy, const Wkv7DstateArgs p) {
    constexpr int N = WKV_N, L = CK_L;
    extern __shared__ __align__(1024) uint8_t dstate_smem_bytes[];
    Wkv7DstateSmem& sm = *reinterpret_cast<Wkv7DstateSmem*>(dstate_smem_bytes);
    uint8_t* const sqb = sm.in;          // A_qb [t][s] MN-major: 2 blocks x 64 k-lines
    uint8_t* const du = sm.in + 16384;   // dU [t][i] MN-major: 2 blocks x 64 k-lines

    const int hh = blockIdx.x, bb = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int qd = warp & 3, cs = warp >> 2;
    const int r = 32 * qd + lane;
    const int T = p.T, H = p.H;
    const int nch = T / L;

    if (tid == 0) {
        mbar_init(&sm.bar_in, 1);
        mbar_init(&sm.bar_mma, 1);
        fence_mbar_init();
        tma_prefetch_desc(&tm_w); tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_a);
        tma_prefetch_desc(&tm_b); tma_prefetch_desc(&tm_dy);
    }
    __syncwarp();
    if (warp == 0) tmem_alloc<256>(&sm.tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    const uint32_t tm_row = tmem + ((uint32_t)(32 * qd) << 16);
    constexpr uint32_t C_SC = 0, C_U = 64, C_DR = 128, C_DS = 192;

    auto issue_loads = [&](int c) {
        mbar_arrive_expect_tx(&sm.bar_in, 5 * L * N * 2);
        const int x0 = hh * N, y0 = bb * T + c * L;
        tma_load_2d(sm.in + 0 * 8192, &tm_w, x0, y0, &sm.bar_in);
        tma_load_2d(sm.in + 1 * 8192, &tm_q, x0, y0, &sm.bar_in);
        tma_load_2d(sm.in + 2 * 8192, &tm_a, x0, y0, &sm.bar_in);
        tma_load_2d(sm.in + 3 * 8192, &tm_b, x0, y0, &sm.bar_in);
        tma_load_2d(sm.in + 4 * 8192, &tm_dy, x0, y0, &sm.bar_in);
    };
    if (nch < 2) {  // nothing to produce
        __syncthreads();
        if (warp == 0) tmem_dealloc<256>(tmem);
        return;
    }
    if (tid == 0) issue_loads(nch - 1);
    __syncwarp();
    if (r < N) {  // dL/dS at the end of the sequence is zero
        uint32_t v[16];
#pragma unroll
        for (int e = 0; e < 16; e++) v[e] = 0u;
        tmem_st16(tm_row + C_DS + 16 * cs, v);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    uint32_t mph = 0, iph = 0;
    auto mma_wait = [&]() {
        mbar_wait(&sm.bar_mma, mph & 1);
        mph++;
        tc_fence_after();
        __syncwarp();
    };
    auto operands_ready = [&]() {
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
    };
    constexpr uint32_t ID_KK = umma_idesc_tf32(128, 64), ID_MM = umma_idesc_tf32(128, 64, 1, 1);

    for (int c = nch - 1; c >= 1; c--) {
        mbar_wait(&sm.bar_in, iph & 1);
        iph++;
        // ================= P1: decay prefix sums and scaled operands =================
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
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int t = 8 * rg + k;
                G += g[k];
                const float E = __expf(G), F = __expf(-G);
                const float qv = bf16lo_to_f32(in16[1 * 4096 + t * N + j]), av_ = bf16lo_to_f32(in16[2 * 4096 + t * N + j]);
                const float bv = bf16lo_to_f32(in16[3 * 4096 + t * N + j]), dyv = bf16lo_to_f32(in16[4 * 4096 + t * N + j]);
                const float at = rt32(av_ * Eprev), qt = rt32(qv * E);
                *reinterpret_cast<float*>(sm.aq + hf * 16384 + sw128_off(t, lane)) = at;
                *reinterpret_cast<float*>(sm.aq + hf * 16384 + sw128_off(64 + t, lane)) = qt;
                *reinterpret_cast<float*>(sm.aq2 + hf * 16384 + sw32_off(t, lane)) = at;
                *reinterpret_cast<float*>(sm.aq2 + hf * 16384 + sw32_off(64 + t, lane)) = qt;
                *reinterpret_cast<float*>(sm.bt + hf * 8192 + sw128_off(t, lane)) = rt32(bv * F);
                *reinterpret_cast<float*>(sm.ry + hf * 16384 + sw32_off(64 + t, lane)) = dyv;
                if (t == L - 1) {
                    sm.el[j] = E;
                    if (G < -80.f) g_chunk_domain_err = 1;
                }
                Eprev = E;
            }
        }
        __syncthreads();
        // ================= dZ = dS_L diag(exp(G_L)): operand + accumulator of dS_0 =================
        if (r < N) {
            uint32_t v[16];
            tmem_ld16(tm_row + C_DS + 16 * cs, v);
#pragma unroll
            for (int e = 0; e < 16; e++) v[e] = __float_as_uint(__uint_as_float(v[e]) * sm.el[16 * cs + e]);
            tmem_st16(tm_row + C_DS + 16 * cs, v);
#pragma unroll
            for (int c4 = 0; c4 < 4; c4++) {
                const int ch = 4 * (cs & 1) + c4;
                *reinterpret_cast<float4*>(sm.dz + (cs >> 1) * 8192 + r * 128 + ((ch ^ (r & 7)) << 4)) =
                    rt32(make_float4(__uint_as_float(v[4 * c4]), __uint_as_float(v[4 * c4 + 1]), __uint_as_float(v[4 * c4 + 2]),
                                     __uint_as_float(v[4 * c4 + 3])));
            }
            tmem_st_wait();
        }
        operands_ready();
        // ================= SC = [At;Qt] Bt^T ;  U = Bt dZ^T =================
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint64_t da = umma_desc_advance(umma_desc_sw128(sm.aq + (k >> 2) * 16384), (k & 3) * 32);
                const uint64_t db = umma_desc_advance(umma_desc_sw128(sm.bt + (k >> 2) * 8192), (k & 3) * 32);
                umma_tf32(tmem + C_SC, da, db, ID_KK, k > 0);
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint64_t da = umma_desc_advance(umma_desc_sw128(sm.bt + (k >> 2) * 8192), (k & 3) * 32);
                const uint64_t db = umma_desc_advance(umma_desc_sw128(sm.dz + (k >> 2) * 8192), (k & 3) * 32);
                umma_tf32(tmem + C_U, da, db, ID_KK, k > 0);
            }
            umma_commit(&sm.bar_mma);
        }
        mma_wait();
        // ================= P2: A_ab (fp32, for the inverse) and A_qb (operand) =================
        *reinterpret_cast<float4*>(sm.tinv + tid * 32) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(sm.tinv + tid * 32 + 16) = make_float4(0.f, 0.f, 0.f, 0.f);
        {
            uint32_t v[16];
            tmem_ld16(tm_row + C_SC + 16 * cs, v);
            if (r < 64) {
#pragma unroll
                for (int c4 = 0; c4 < 4; c4++) {
                    float o[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) o[e] = (16 * cs + 4 * c4 + e < r) ? __uint_as_float(v[4 * c4 + e]) : 0.f;
                    *reinterpret_cast<float4*>(sm.aab + r * 256 + (((4 * cs + c4) ^ (r & 7)) << 4)) = make_float4(o[0], o[1], o[2], o[3]);
                }
            } else {
                const int t = r - 64;
#pragma unroll
                for (int c4 = 0; c4 < 4; c4++) {
                    float o[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) o[e] = (16 * cs + 4 * c4 + e <= t) ? rt32(__uint_as_float(v[4 * c4 + e])) : 0.f;
                    *reinterpret_cast<float4*>(sqb + (cs >> 1) * 8192 + sw32_off(t, 16 * (cs & 1) + 4 * c4)) =
                        make_float4(o[0], o[1], o[2], o[3]);
                }
            }
        }
        operands_ready();
        // ================= U += A_qb^T dY (runs while the inverse is being computed) =================
        if (tid == 0) {
            tc_fence_after();
            const uint64_t da = umma_desc_mn_tf32(sqb, 8192, 512);
            const uint64_t db = umma_desc_mn_tf32(sm.ry + 64 * 128, 16384, 512);
#pragma unroll
            for (int k = 0; k < 8; k++) umma_tf32(tmem + C_U, umma_desc_advance(da, k * 1024), umma_desc_advance(db, k * 1024), ID_MM, 1);
            umma_commit(&sm.bar_mma);
        }
        // ---- Tinv = (I - A_ab)^-1, written row-major [t][s] in the MN-major operand layout ----
        chunk_tri_inverse(sm.aab, sm.esc, tid, [&](int t, int s) { return sm.tinv + (s >> 5) * 8192 + sw32_off(t, s & 31); });
        mma_wait();
        // ================= dU -> operand =================
        if (r < 64) {
            uint32_t v[16];
            tmem_ld16(tm_row + C_U + 16 * cs, v);
#pragma unroll
            for (int c4 = 0; c4 < 4; c4++)
                *reinterpret_cast<float4*>(du + (cs >> 1) * 8192 + sw32_off(r, 16 * (cs & 1) + 4 * c4)) =
                    rt32(make_float4(__uint_as_float(v[4 * c4]), __uint_as_float(v[4 * c4 + 1]), __uint_as_float(v[4 * c4 + 2]),
                                     __uint_as_float(v[4 * c4 + 3])));
        }
        operands_ready();
        // ================= dR = Tinv^T dU =================
        if (tid == 0) {
            tc_fence_after();
            const uint64_t da = umma_desc_mn_tf32(sm.tinv, 8192, 512);
            const uint64_t db = umma_desc_mn_tf32(du, 8192, 512);
#pragma unroll
            for (int k = 0; k < 8; k++) umma_tf32(tmem + C_DR, umma_desc_advance(da, k * 1024), umma_desc_advance(db, k * 1024), ID_MM, k > 0);
            umma_commit(&sm.bar_mma);
        }
        mma_wait();
        if (tid == 0 && c - 1 >= 1) issue_loads(c - 1);  // SQB / DU are dead: the input buffer is free again
        __syncwarp();
        if (r < 64) {
            uint32_t v[16];
            tmem_ld16(tm_row + C_DR + 16 * cs, v);
#pragma unroll
            for (int c4 = 0; c4 < 4; c4++)
                *reinterpret_cast<float4*>(sm.ry + (cs >> 1) * 16384 + sw32_off(r, 16 * (cs & 1) + 4 * c4)) =
                    rt32(make_float4(__uint_as_float(v[4 * c4]), __uint_as_float(v[4 * c4 + 1]), __uint_as_float(v[4 * c4 + 2]),
                                     __uint_as_float(v[4 * c4 + 3])));
        }
        operands_ready();
        // ================= dS_0 = dZ + [dR;dY]^T [At;Qt] =================
        if (tid == 0) {
            tc_fence_after();
            const uint64_t da = umma_desc_mn_tf32(sm.ry, 16384, 512);
            const uint64_t db = umma_desc_mn_tf32(sm.aq2, 16384, 512);
#pragma unroll
            for (int k = 0; k < 16; k++) umma_tf32(tmem + C_DS, umma_desc_advance(da, k * 1024), umma_desc_advance(db, k * 1024), ID_MM, 1);
            umma_commit(&sm.bar_mma);
        }
        mma_wait();
        if (r < N) {
            uint32_t v[16];
            tmem_ld16(tm_row + C_DS + 16 * cs, v);
            float4* dst = reinterpret_cast<float4*>(p.ds + ((((size_t)bb * H + hh) * nch + (c - 1)) * N + r) * N + 16 * cs);
#pragma unroll
            for (int c4 = 0; c4 < 4; c4++)
                dst[c4] = make_float4(__uint_as_float(v[4 * c4]), __uint_as_float(v[4 * c4 + 1]), __uint_as_float(v[4 * c4 + 2]),
                                      __uint_as_float(v[4 * c4 + 3]));
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tmem);
}

}  // namespace vrwkv
