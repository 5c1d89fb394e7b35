// fused_tmix.cu — the element-wise / per-head parts of RWKV_Tmix_x070 and RWKV_CMix_x070 as four fused kernels
// (forward + backward each), replacing ~35 bf16 eager kernels per layer of the reference
// (VisualRWKV-v7/v7.00/src/model.py:176-193, 225):
//
//   tmix_mid  : w = -softplus(-(w0+ww)) - 0.5 ; a = sigmoid(a0+aa) ; v' = v + (v_first - v) sigmoid(v0+vv) ;
//               kk = normalize_head(k*k_k) ; k' = k (1 + (a-1) k_a) ; emits w, k', v', -kk, kk*a   (:176-190)
//   tmix_post : GroupNorm_H(y, eps) + (sum_head r k' r_k) v' , times g                               (:191-194)
//   relu_sq   : relu(x)^2                                                                            (:225)
//
// Every bf16 rounding point of the eager graph is reproduced in the forward kernels; backward kernels treat the
// roundings as identities (what autograd does).  Mapping as in rowops.cuh: thread = 8 consecutive channels, a
// 64-channel head = 8 adjacent lanes (per-head sums are three xor-shuffles), CTA = run of rows; per-channel
// parameter gradients leave as one fp32 partial row per CTA.
#include "host_util.h"
#include "rowops.cuh"

namespace vrwkv {

constexpr int TM_RUN = 16;

__device__ __forceinline__ float head_sum(float x) {  // over the 8 lanes (64 channels) of a head
    x += __shfl_xor_sync(0xffffffffu, x, 1);
    x += __shfl_xor_sync(0xffffffffu, x, 2);
    x += __shfl_xor_sync(0xffffffffu, x, 4);
    return x;
}
// MUFU-based fast forms: their ~2 ulp fp32 error is far below the bf16 rounding applied right after
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float softplusf_(float x) { return x > 20.f ? x : __logf(1.f + __expf(x)); }

// ------------------------------------------------------------------------------------------------------------
// tmix_mid
// ------------------------------------------------------------------------------------------------------------
struct TmixMidArgs {
    int rows, C, has_vres;
    const uint16_t *k, *v, *vfirst, *ww, *aa, *vv;       // inputs [rows, C]
    const uint16_t *w0, *a0, *v0, *k_k, *k_a;            // params [C]
    uint16_t *w, *k2, *v2, *nkk, *kka;                   // forward outputs
    // backward
    const uint16_t *dw, *dk2, *dv2, *dnkk, *dkka;
    const uint16_t *dk2b, *dv2b;  // optional second contributions to dk2 / dv2 (summed here)
    uint16_t *dk, *dv, *dvfirst, *dww, *daa, *dvv;
    float* partial;  // [grid][5][C]: dw0, da0, dv0, dk_k, dk_a
};

__device__ __forceinline__ uint32_t& u4(uint4& v, int i) { return reinterpret_cast<uint32_t*>(&v)[i]; }
__device__ __forceinline__ uint32_t u4(const uint4& v, int i) { return reinterpret_cast<const uint32_t*>(&v)[i]; }

// The chain of bf16 element-wise ops of model.py:176-190 runs in packed bf16x2 arithmetic (rowops.cuh: one rounding per
// op, as the eager graph), two channels per instruction; only softplus / sigmoid / the head norm go through fp32.  The
// fp32 version of this kernel spent ~120 issue slots per element on convert-and-shift round trips (ncu: 42 % issue at
// 18 % occupancy, 2.9 TB/s).
__global__ void __launch_bounds__(256) tmix_mid_fwd_kernel(const TmixMidArgs a) {
    const int c0 = threadIdx.x * 8;
    const bool active = c0 < a.C;
    const uint4 w0 = ldraw(active, a.w0 + c0), a0 = ldraw(active, a.a0 + c0), kk_ = ldraw(active, a.k_k + c0), ka = ldraw(active, a.k_a + c0);
    const uint4 v0 = ldraw(active && a.has_vres, a.v0 + c0);
    const int row0 = blockIdx.x * TM_RUN, row1 = min(row0 + TM_RUN, a.rows);
    constexpr uint32_t ONE2 = 0x3F803F80u, HALF2 = 0x3F003F00u, SIGN2 = 0x80008000u;
    // software pipeline: the loads of row+1 are issued before row is computed and stored
    struct In { uint4 k, v, ww, aa, vf, vv; };
    auto load = [&](int row) {
        In r;
        const size_t o = (size_t)row * a.C + c0;
        const bool ok_ = active && row < row1;
        r.k = ldraw(ok_, a.k + o); r.v = ldraw(ok_, a.v + o); r.ww = ldraw(ok_, a.ww + o); r.aa = ldraw(ok_, a.aa + o);
        r.vf = ldraw(ok_ && a.has_vres, a.vfirst + o); r.vv = ldraw(ok_ && a.has_vres, a.vv + o);
        return r;
    };
    In nxt = load(row0);
    for (int row = row0; row < row1; row++) {
        const size_t o = (size_t)row * a.C + c0;
        const In cur = nxt;
        nxt = load(row + 1);
        uint4 ow, ok, ov, onkk, okka, u, av;
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            // w = -softplus(-(w0 + ww)) - 0.5
            const uint32_t z = bf2_add(u4(w0, i), u4(cur.ww, i));
            const uint32_t sp = pack_bf16x2(softplusf_(-bf16lo_to_f32(z)), softplusf_(-bf16hi_to_f32(z)));
            u4(ow, i) = bf2_sub(sp ^ SIGN2, HALF2);
            // a = sigmoid(a0 + aa)
            const uint32_t za = bf2_add(u4(a0, i), u4(cur.aa, i));
            const uint32_t avp = pack_bf16x2(sigmoidf_(bf16lo_to_f32(za)), sigmoidf_(bf16hi_to_f32(za)));
            u4(av, i) = avp;
            // u = k * k_k (normalised below);  k' = k * (1 + (a - 1) * k_a)
            const uint32_t up = bf2_mul(u4(cur.k, i), u4(kk_, i));
            u4(u, i) = up;
            const float u0 = bf16lo_to_f32(up), u1 = bf16hi_to_f32(up);
            ss = fmaf(u0, u0, ss);
            ss = fmaf(u1, u1, ss);
            u4(ok, i) = bf2_mul(u4(cur.k, i), bf2_add(ONE2, bf2_mul(bf2_sub(avp, ONE2), u4(ka, i))));
            // v' = v + (v_first - v) * sigmoid(v0 + vv)      (layers > 0)
            if (a.has_vres) {
                const uint32_t zv = bf2_add(u4(v0, i), u4(cur.vv, i));
                const uint32_t vg = pack_bf16x2(sigmoidf_(bf16lo_to_f32(zv)), sigmoidf_(bf16hi_to_f32(zv)));
                u4(ov, i) = bf2_add(u4(cur.v, i), bf2_mul(bf2_sub(u4(cur.vf, i), u4(cur.v, i)), vg));
            } else {
                u4(ov, i) = u4(cur.v, i);
            }
        }
        const float nrm = fmaxf(rb(sqrtf(head_sum(ss))), 1e-12f);  // F.normalize(p=2, eps=1e-12)
        const float inrm = 1.f / nrm;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t kk = pack_bf16x2(bf16lo_to_f32(u4(u, i)) * inrm, bf16hi_to_f32(u4(u, i)) * inrm);
            u4(onkk, i) = kk ^ SIGN2;
            u4(okka, i) = bf2_mul(kk, u4(av, i));
        }
        if (active) {
            *reinterpret_cast<uint4*>(a.w + o) = ow; *reinterpret_cast<uint4*>(a.k2 + o) = ok; *reinterpret_cast<uint4*>(a.v2 + o) = ov;
            *reinterpret_cast<uint4*>(a.nkk + o) = onkk; *reinterpret_cast<uint4*>(a.kka + o) = okka;
        }
    }
}

__global__ void __launch_bounds__(256) tmix_mid_bwd_kernel(const TmixMidArgs a) {
    const int c0 = threadIdx.x * 8;
    const bool active = c0 < a.C;
    const F8 w0 = ldz(active, a.w0 + c0), a0 = ldz(active, a.a0 + c0), kk_ = ldz(active, a.k_k + c0), ka = ldz(active, a.k_a + c0);
    F8 v0 = zero8();
    if (a.has_vres) v0 = ldz(active, a.v0 + c0);
    F8 gw0 = zero8(), ga0 = zero8(), gv0 = zero8(), gkk = zero8(), gka = zero8();
    const int row0 = blockIdx.x * TM_RUN, row1 = min(row0 + TM_RUN, a.rows);
    struct In { uint4 k, ww, aa, dw, dk2, dv2, dnkk, dkka, v, vf, vv, dk2b, dv2b; };
    auto load = [&](int row) {
        In r;
        const size_t o = (size_t)row * a.C + c0;
        const bool ok_ = active && row < row1;
        r.k = ldraw(ok_, a.k + o); r.ww = ldraw(ok_, a.ww + o); r.aa = ldraw(ok_, a.aa + o);
        r.dw = ldraw(ok_, a.dw + o); r.dk2 = ldraw(ok_, a.dk2 + o); r.dv2 = ldraw(ok_, a.dv2 + o);
        r.dnkk = ldraw(ok_, a.dnkk + o); r.dkka = ldraw(ok_, a.dkka + o);
        r.dk2b = ldraw(ok_ && a.dk2b != nullptr, a.dk2b + o); r.dv2b = ldraw(ok_ && a.dv2b != nullptr, a.dv2b + o);
        const bool ov_ = ok_ && a.has_vres;
        r.v = ldraw(ov_, a.v + o); r.vf = ldraw(ov_, a.vfirst + o); r.vv = ldraw(ov_, a.vv + o);
        return r;
    };
    In nxt = load(row0);
    for (int row = row0; row < row1; row++) {
        const size_t o = (size_t)row * a.C + c0;
        const In cur = nxt;
        nxt = load(row + 1);
        const F8 k = f8(cur.k), ww = f8(cur.ww), aa = f8(cur.aa), dw = f8(cur.dw), dnkk = f8(cur.dnkk), dkka = f8(cur.dkka);
        F8 dk2 = f8(cur.dk2), dv2 = f8(cur.dv2);
        {
            const F8 kb = f8(cur.dk2b), vb = f8(cur.dv2b);  // zeros when absent
#pragma unroll
            for (int e = 0; e < 8; e++) {
                dk2.v[e] += kb.v[e];
                dv2.v[e] += vb.v[e];
            }
        }
        F8 u, av, kk, dkk, odk, odww, odaa;
        float ss = 0.f;
#pragma unroll
        for (int e = 0; e < 8; e++) {
            av.v[e] = sigmoidf_(a0.v[e] + aa.v[e]);
            u.v[e] = k.v[e] * kk_.v[e];
            ss += u.v[e] * u.v[e];
        }
        const float n2 = head_sum(ss);
        const float nrm = fmaxf(sqrtf(n2), 1e-12f), inrm = 1.f / nrm;
        float dot = 0.f;
#pragma unroll
        for (int e = 0; e < 8; e++) {
            kk.v[e] = u.v[e] * inrm;
            dkk.v[e] = -dnkk.v[e] + dkka.v[e] * av.v[e];
            dot += dkk.v[e] * kk.v[e];
        }
        dot = head_sum(dot);
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const float du = (dkk.v[e] - kk.v[e] * dot) * inrm;
            const float da = dkka.v[e] * kk.v[e] + dk2.v[e] * k.v[e] * ka.v[e];
            odk.v[e] = dk2.v[e] * (1.f + (av.v[e] - 1.f) * ka.v[e]) + du * kk_.v[e];
            gkk.v[e] += du * k.v[e];
            gka.v[e] += dk2.v[e] * k.v[e] * (av.v[e] - 1.f);
            const float dpa = da * av.v[e] * (1.f - av.v[e]);
            odaa.v[e] = dpa;
            ga0.v[e] += dpa;
            const float z = w0.v[e] + ww.v[e];
            const float dz = dw.v[e] * sigmoidf_(-z);  // d/dz [-softplus(-z)] = sigmoid(-z)
            odww.v[e] = dz;
            gw0.v[e] += dz;
        }
        stz(active, a.dk + o, odk); stz(active, a.dww + o, odww); stz(active, a.daa + o, odaa);
        if (a.has_vres) {
            const F8 v = f8(cur.v), vf = f8(cur.vf), vv = f8(cur.vv);
            F8 odv, odvf, odvv;
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const float vg = sigmoidf_(v0.v[e] + vv.v[e]);
                odv.v[e] = dv2.v[e] * (1.f - vg);
                odvf.v[e] = dv2.v[e] * vg;
                const float dp = dv2.v[e] * (vf.v[e] - v.v[e]) * vg * (1.f - vg);
                odvv.v[e] = dp;
                gv0.v[e] += dp;
            }
            stz(active, a.dv + o, odv); stz(active, a.dvfirst + o, odvf); stz(active, a.dvv + o, odvv);
        } else {
            stz(active, a.dv + o, dv2);
        }
    }
    if (!active) return;
    float* dst = a.partial + (size_t)blockIdx.x * 5 * a.C + c0;
#pragma unroll
    for (int e = 0; e < 8; e++) {
        dst[e] = gw0.v[e]; dst[a.C + e] = ga0.v[e]; dst[2 * a.C + e] = gv0.v[e];
        dst[3 * a.C + e] = gkk.v[e]; dst[4 * a.C + e] = gka.v[e];
    }
}

// ------------------------------------------------------------------------------------------------------------
// tmix_post
// ------------------------------------------------------------------------------------------------------------
struct TmixPostArgs {
    int rows, C;
    float eps;
    const uint16_t *y, *r, *k2, *v2, *g;     // inputs [rows, C]
    const uint16_t *gamma, *beta, *r_k;      // params [C]
    uint16_t* z;                             // forward output
    const uint16_t* dz;                      // backward
    uint16_t *dy, *dr, *dk2, *dv2, *dg;
    float* partial;  // [grid][3][C]: dgamma, dbeta, dr_k
};

__global__ void __launch_bounds__(256) tmix_post_fwd_kernel(const TmixPostArgs a) {
    const int c0 = threadIdx.x * 8;
    const bool active = c0 < a.C;
    const F8 gm = ldz(active, a.gamma + c0), bt = ldz(active, a.beta + c0), rk = ldz(active, a.r_k + c0);
    const int row0 = blockIdx.x * TM_RUN, row1 = min(row0 + TM_RUN, a.rows);
    struct In { uint4 y, r, k, v, g; };
    auto load = [&](int row) {
        In q;
        const size_t o = (size_t)row * a.C + c0;
        const bool ok_ = active && row < row1;
        q.y = ldraw(ok_, a.y + o); q.r = ldraw(ok_, a.r + o); q.k = ldraw(ok_, a.k2 + o); q.v = ldraw(ok_, a.v2 + o); q.g = ldraw(ok_, a.g + o);
        return q;
    };
    In nxt = load(row0);
    for (int row = row0; row < row1; row++) {
        const size_t o = (size_t)row * a.C + c0;
        const In cur = nxt;
        nxt = load(row + 1);
        const F8 y = f8(cur.y), r = f8(cur.r), k = f8(cur.k), v = f8(cur.v), g = f8(cur.g);
        float s1 = 0.f, sb = 0.f;
#pragma unroll
        for (int e = 0; e < 8; e++) {
            s1 += y.v[e];
            sb += rb(rb(r.v[e] * k.v[e]) * rk.v[e]);
        }
        const float mean = head_sum(s1) * (1.f / 64.f);
        float s2 = 0.f;
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const float d = y.v[e] - mean;
            s2 += d * d;
        }
        const float rstd = rsqrtf(head_sum(s2) * (1.f / 64.f) + a.eps);
        const float bonus = rb(head_sum(sb));
        F8 z;
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const float yn = rb((y.v[e] - mean) * rstd * gm.v[e] + bt.v[e]);
            z.v[e] = rb(yn + rb(bonus * v.v[e])) * g.v[e];
        }
        stz(active, a.z + o, z);
    }
}

__global__ void __launch_bounds__(256) tmix_post_bwd_kernel(const TmixPostArgs a) {
    const int c0 = threadIdx.x * 8;
    const bool active = c0 < a.C;
    const F8 gm = ldz(active, a.gamma + c0), bt = ldz(active, a.beta + c0), rk = ldz(active, a.r_k + c0);
    F8 ggm = zero8(), gbt = zero8(), grk = zero8();
    const int row0 = blockIdx.x * TM_RUN, row1 = min(row0 + TM_RUN, a.rows);
    struct In { uint4 y, r, k, v, g, dz; };
    auto load = [&](int row) {
        In q;
        const size_t o = (size_t)row * a.C + c0;
        const bool ok_ = active && row < row1;
        q.y = ldraw(ok_, a.y + o); q.r = ldraw(ok_, a.r + o); q.k = ldraw(ok_, a.k2 + o); q.v = ldraw(ok_, a.v2 + o); q.g = ldraw(ok_, a.g + o);
        q.dz = ldraw(ok_, a.dz + o);
        return q;
    };
    In nxt = load(row0);
    for (int row = row0; row < row1; row++) {
        const size_t o = (size_t)row * a.C + c0;
        const In cur = nxt;
        nxt = load(row + 1);
        const F8 y = f8(cur.y), r = f8(cur.r), k = f8(cur.k), v = f8(cur.v), g = f8(cur.g), dz = f8(cur.dz);
        float s1 = 0.f, sb = 0.f;
#pragma unroll
        for (int e = 0; e < 8; e++) {
            s1 += y.v[e];
            sb += r.v[e] * k.v[e] * rk.v[e];
        }
        const float mean = head_sum(s1) * (1.f / 64.f);
        float s2 = 0.f;
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const float d = y.v[e] - mean;
            s2 += d * d;
        }
        const float rstd = rsqrtf(head_sum(s2) * (1.f / 64.f) + a.eps);
        const float bonus = head_sum(sb);
        F8 xh, du, odg, odv;
        float ds = 0.f, m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int e = 0; e < 8; e++) {
            xh.v[e] = (y.v[e] - mean) * rstd;
            const float u = xh.v[e] * gm.v[e] + bt.v[e] + bonus * v.v[e];
            du.v[e] = dz.v[e] * g.v[e];
            odg.v[e] = dz.v[e] * u;
            ds += du.v[e] * v.v[e];
            odv.v[e] = du.v[e] * bonus;
            const float d = du.v[e] * gm.v[e];
            m1 += d;
            m2 += d * xh.v[e];
            ggm.v[e] += du.v[e] * xh.v[e];
            gbt.v[e] += du.v[e];
        }
        ds = head_sum(ds);
        m1 = head_sum(m1) * (1.f / 64.f);
        m2 = head_sum(m2) * (1.f / 64.f);
        F8 ody, odr, odk;
#pragma unroll
        for (int e = 0; e < 8; e++) {
            ody.v[e] = rstd * (du.v[e] * gm.v[e] - m1 - xh.v[e] * m2);
            odr.v[e] = ds * k.v[e] * rk.v[e];
            odk.v[e] = ds * r.v[e] * rk.v[e];
            grk.v[e] += ds * r.v[e] * k.v[e];
        }
        stz(active, a.dy + o, ody); stz(active, a.dr + o, odr); stz(active, a.dk2 + o, odk);
        stz(active, a.dv2 + o, odv); stz(active, a.dg + o, odg);
    }
    if (!active) return;
    float* dst = a.partial + (size_t)blockIdx.x * 3 * a.C + c0;
#pragma unroll
    for (int e = 0; e < 8; e++) {
        dst[e] = ggm.v[e]; dst[a.C + e] = gbt.v[e]; dst[2 * a.C + e] = grk.v[e];
    }
}

// ------------------------------------------------------------------------------------------------------------
// relu^2 (channel-mix, model.py:225) — flat element-wise, 8 elements per thread, grid-stride
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) relu_sq_fwd_kernel(const uint16_t* x, uint16_t* y, size_t n8) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        F8 v = ld_bf16x8(x + i * 8);
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const float r = fmaxf(v.v[e], 0.f);
            v.v[e] = rb(r) * rb(r);
        }
        st_bf16x8(y + i * 8, v);
    }
}
__global__ void __launch_bounds__(256) relu_sq_bwd_kernel(const uint16_t* x, const uint16_t* dy, uint16_t* dx, size_t n8) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        const F8 v = ld_bf16x8(x + i * 8), d = ld_bf16x8(dy + i * 8);
        F8 o;
#pragma unroll
        for (int e = 0; e < 8; e++) o.v[e] = 2.f * fmaxf(v.v[e], 0.f) * d.v[e];
        st_bf16x8(dx + i * 8, o);
    }
}

// backward of relu(x)^2 from the ACTIVATION y = relu(x)^2 alone (the tcgen05 GEMM epilogue emits y and never
// materialises x): d/dx = 2 relu(x) = 2 sqrt(y)
__global__ void __launch_bounds__(256) relu_sq_bwd_from_act_kernel(const uint16_t* y, const uint16_t* dy, uint16_t* dx, size_t n8) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        const F8 v = ld_bf16x8(y + i * 8), d = ld_bf16x8(dy + i * 8);
        F8 o;
#pragma unroll
        for (int e = 0; e < 8; e++) o.v[e] = 2.f * sqrtf(v.v[e]) * d.v[e];
        st_bf16x8(dx + i * 8, o);
    }
}

// second stage of the per-channel parameter gradients: out[i] (bf16) = sum over blocks of partial[blk][i], i in [0, n*C).
// CTA = 8 column groups (4 consecutive columns each: one 128-byte line per block row) x 32 slices of the block range,
// 8 independent 16-byte loads in flight per thread; the slices are combined through shared memory.  A [512, 6144]
// array of partials becomes 192 CTAs with ~6 MB in flight instead of 24 CTAs walking 128 rows each.
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* partial, uint16_t* out, int nb, int nc) {
    __shared__ float4 red[32][8];
    const int cg = threadIdx.x & 7, slice = threadIdx.x >> 3;
    const int col = (blockIdx.x * 8 + cg) * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col < nc) {
        for (int b0 = slice; b0 < nb; b0 += 32 * 8) {
            float4 x[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int b = b0 + 32 * u;
                x[u] = (b < nb) ? __ldg(reinterpret_cast<const float4*>(partial + (size_t)b * nc + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                s.x += x[u].x; s.y += x[u].y; s.z += x[u].z; s.w += x[u].w;
            }
        }
    }
    red[slice][cg] = s;
    __syncthreads();
    if (slice < 4) {  // 32 -> 4 -> 1
        float4 t = red[slice][cg];
#pragma unroll
        for (int k = 1; k < 8; k++) {
            const float4 o = red[slice + 4 * k][cg];
            t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
        }
        red[slice][cg] = t;
    }
    __syncthreads();
    if (slice == 0 && col < nc) {
        float4 t = red[0][cg];
#pragma unroll
        for (int k = 1; k < 4; k++) {
            t.x += red[k][cg].x; t.y += red[k][cg].y; t.z += red[k][cg].z; t.w += red[k][cg].w;
        }
        uint2 o;
        o.x = pack_bf16x2(t.x, t.y);
        o.y = pack_bf16x2(t.z, t.w);
        *reinterpret_cast<uint2*>(out + col) = o;
    }
}

}  // namespace vrwkv

using namespace vrwkv;

static int tm_check(int rows, int C, const char* what) {
    if (rows <= 0 || C <= 0) return vrwkv_fail(VRWKV_EINVAL, "%s: bad shape (%d,%d)", what, rows, C);
    if (C % 64 || C / 8 > 256) return vrwkv_fail(VRWKV_EUNSUP, "%s: C=%d must be a multiple of 64 and <= 2048", what, C);
    return VRWKV_OK;
}

extern "C" int vrwkv_tmix_blocks(int rows) { return (rows + TM_RUN - 1) / TM_RUN; }

extern "C" int vrwkv_tmix_mid_forward(int rows, int C, const uint16_t* k, const uint16_t* v, const uint16_t* vfirst,
                                      const uint16_t* ww, const uint16_t* aa, const uint16_t* vv, const uint16_t* w0,
                                      const uint16_t* a0, const uint16_t* v0, const uint16_t* k_k, const uint16_t* k_a,
                                      uint16_t* w, uint16_t* k2, uint16_t* v2, uint16_t* nkk, uint16_t* kka, void* stream) {
    int rc = tm_check(rows, C, "tmix_mid_forward");
    if (rc) return rc;
    TmixMidArgs a{};
    a.rows = rows; a.C = C; a.has_vres = vfirst != nullptr;
    a.k = k; a.v = v; a.vfirst = vfirst; a.ww = ww; a.aa = aa; a.vv = vv;
    a.w0 = w0; a.a0 = a0; a.v0 = v0; a.k_k = k_k; a.k_a = k_a;
    a.w = w; a.k2 = k2; a.v2 = v2; a.nkk = nkk; a.kka = kka;
    if (!k || !v || !ww || !aa || !w0 || !a0 || !k_k || !k_a || !w || !k2 || !v2 || !nkk || !kka || (a.has_vres && (!vv || !v0)))
        return vrwkv_fail(VRWKV_EINVAL, "tmix_mid_forward: null pointer");
    tmix_mid_fwd_kernel<<<vrwkv_tmix_blocks(rows), row_threads(C), 0, (cudaStream_t)stream>>>(a);
    VRWKV_CUDA(cudaGetLastError());
    vrwkv_count_launch(1);
    return VRWKV_OK;
}

extern "C" int vrwkv_tmix_mid_backward(int rows, int C, const uint16_t* k, const uint16_t* v, const uint16_t* vfirst,
                                       const uint16_t* ww, const uint16_t* aa, const uint16_t* vv, const uint16_t* w0,
                                       const uint16_t* a0, const uint16_t* v0, const uint16_t* k_k, const uint16_t* k_a,
                                       const uint16_t* dw, const uint16_t* dk2, const uint16_t* dv2, const uint16_t* dnkk,
                                       const uint16_t* dkka, const uint16_t* dk2b, const uint16_t* dv2b, uint16_t* dk, uint16_t* dv,
                                       uint16_t* dvfirst, uint16_t* dww, uint16_t* daa, uint16_t* dvv, float* partial, void* stream) {
    int rc = tm_check(rows, C, "tmix_mid_backward");
    if (rc) return rc;
    TmixMidArgs a{};
    a.rows = rows; a.C = C; a.has_vres = vfirst != nullptr;
    a.k = k; a.v = v; a.vfirst = vfirst; a.ww = ww; a.aa = aa; a.vv = vv;
    a.w0 = w0; a.a0 = a0; a.v0 = v0; a.k_k = k_k; a.k_a = k_a;
    a.dw = dw; a.dk2 = dk2; a.dv2 = dv2; a.dnkk = dnkk; a.dkka = dkka; a.dk2b = dk2b; a.dv2b = dv2b;
    a.dk = dk; a.dv = dv; a.dvfirst = dvfirst; a.dww = dww; a.daa = daa; a.dvv = dvv; a.partial = partial;
    if (!k || !v || !ww || !aa || !w0 || !a0 || !k_k || !k_a || !dw || !dk2 || !dv2 || !dnkk || !dkka || !dk || !dv || !dww ||
        !daa || !partial || (a.has_vres && (!vv || !v0 || !dvfirst || !dvv)))
        return vrwkv_fail(VRWKV_EINVAL, "tmix_mid_backward: null pointer");
    tmix_mid_bwd_kernel<<<vrwkv_tmix_blocks(rows), row_threads(C), 0, (cudaStream_t)stream>>>(a);
    VRWKV_CUDA(cudaGetLastError());
    vrwkv_count_launch(1);
    return VRWKV_OK;
}

extern "C" int vrwkv_tmix_post_forward(int rows, int C, float eps, const uint16_t* y, const uint16_t* r, const uint16_t* k2,
                                       const uint16_t* v2, const uint16_t* g, const uint16_t* gamma, const uint16_t* beta,
                                       const uint16_t* r_k, uint16_t* z, void* stream) {
    int rc = tm_check(rows, C, "tmix_post_forward");
    if (rc) return rc;
    if (!y || !r || !k2 || !v2 || !g || !gamma || !beta || !r_k || !z) return vrwkv_fail(VRWKV_EINVAL, "tmix_post_forward: null pointer");
    TmixPostArgs a{};
    a.rows = rows; a.C = C; a.eps = eps; a.y = y; a.r = r; a.k2 = k2; a.v2 = v2; a.g = g;
    a.gamma = gamma; a.beta = beta; a.r_k = r_k; a.z = z;
    tmix_post_fwd_kernel<<<vrwkv_tmix_blocks(rows), row_threads(C), 0, (cudaStream_t)stream>>>(a);
    VRWKV_CUDA(cudaGetLastError());
    vrwkv_count_launch(1);
    return VRWKV_OK;
}

extern "C" int vrwkv_tmix_post_backward(int rows, int C, float eps, const uint16_t* y, const uint16_t* r, const uint16_t* k2,
                                        const uint16_t* v2, const uint16_t* g, const uint16_t* gamma, const uint16_t* beta,
                                        const uint16_t* r_k, const uint16_t* dz, uint16_t* dy, uint16_t* dr, uint16_t* dk2,
                                        uint16_t* dv2, uint16_t* dg, float* partial, void* stream) {
    int rc = tm_check(rows, C, "tmix_post_backward");
    if (rc) return rc;
    if (!y || !r || !k2 || !v2 || !g || !gamma || !beta || !r_k || !dz || !dy || !dr || !dk2 || !dv2 || !dg || !partial)
        return vrwkv_fail(VRWKV_EINVAL, "tmix_post_backward: null pointer");
    TmixPostArgs a{};
    a.rows = rows; a.C = C; a.eps = eps; a.y = y; a.r = r; a.k2 = k2; a.v2 = v2; a.g = g;
    a.gamma = gamma; a.beta = beta; a.r_k = r_k; a.dz = dz; a.dy = dy; a.dr = dr; a.dk2 = dk2; a.dv2 = dv2; a.dg = dg;
    a.partial = partial;
    tmix_post_bwd_kernel<<<vrwkv_tmix_blocks(rows), row_threads(C), 0, (cudaStream_t)stream>>>(a);
    VRWKV_CUDA(cudaGetLastError());
    vrwkv_count_launch(1);
    return VRWKV_OK;
}

extern "C" int vrwkv_relu_sq_forward(size_t n, const uint16_t* x, uint16_t* y, void* stream) {
    if (!x || !y || (n % 8)) return vrwkv_fail(VRWKV_EINVAL, "relu_sq_forward: null pointer or n %% 8 != 0");
    relu_sq_fwd_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(x, y, n / 8);
    VRWKV_CUDA(cudaGetLastError());
    vrwkv_count_launch(1);
    return VRWKV_OK;
}
extern "C" int vrwkv_relu_sq_backward(size_t n, const uint16_t* x, const uint16_t* dy, uint16_t* dx, void* stream) {
    if (!x || !dy || !dx || (n % 8)) return vrwkv_fail(VRWKV_EINVAL, "relu_sq_backward: null pointer or n %% 8 != 0");
    relu_sq_bwd_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(x, dy, dx, n / 8);
    VRWKV_CUDA(cudaGetLastError());
    vrwkv_count_launch(1);
    return VRWKV_OK;
}

extern "C" int vrwkv_relu_sq_backward_from_act(size_t n, const uint16_t* y, const uint16_t* dy, uint16_t* dx, void* stream) {
    if (!y || !dy || !dx || (n % 8)) return vrwkv_fail(VRWKV_EINVAL, "relu_sq_backward_from_act: null pointer or n %% 8 != 0");
    relu_sq_bwd_from_act_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(y, dy, dx, n / 8);
    VRWKV_CUDA(cudaGetLastError());
    vrwkv_count_launch(1);
    return VRWKV_OK;
}

extern "C" int vrwkv_reduce_partials(int nblocks, int n_times_c, const float* partial, uint16_t* out, void* stream) {
    if (nblocks <= 0 || n_times_c <= 0 || !partial || !out) return vrwkv_fail(VRWKV_EINVAL, "reduce_partials: bad arguments");
    if (n_times_c % 4) return vrwkv_fail(VRWKV_EINVAL, "reduce_partials: n*C must be a multiple of 4");
    reduce_partials_kernel<<<(n_times_c / 4 + 7) / 8, 256, 0, (cudaStream_t)stream>>>(partial, out, nblocks, n_times_c);
    VRWKV_CUDA(cudaGetLastError());
    vrwkv_count_launch(1);
    return VRWKV_OK;
}
