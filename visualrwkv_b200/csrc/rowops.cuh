// rowops.cuh — helpers for the row-wise fused kernels (LayerNorm, token-shift mixes, per-head ops).
//
// Mapping used by every kernel here: a CTA owns a run of consecutive rows (tokens) and has C/8 threads;
// thread `tid` owns the 8 consecutive channels [8*tid, 8*tid+8) of every row it visits (one 128-bit load /
// store per tensor per row, fully coalesced).  Per-row statistics are block-reduced through shared memory;
// per-channel parameter gradients accumulate in the owning thread's registers over the CTA's rows and leave
// as one fp32 partial row per CTA (deterministic second-stage sum on the host side).
#pragma once
#include "common.cuh"

namespace vrwkv {

struct F8 {
    float v[8];
};

__device__ __forceinline__ F8 ld_bf16x8(const uint16_t* p) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    F8 r;
    r.v[0] = bf16lo_to_f32(u.x); r.v[1] = bf16hi_to_f32(u.x);
    r.v[2] = bf16lo_to_f32(u.y); r.v[3] = bf16hi_to_f32(u.y);
    r.v[4] = bf16lo_to_f32(u.z); r.v[5] = bf16hi_to_f32(u.z);
    r.v[6] = bf16lo_to_f32(u.w); r.v[7] = bf16hi_to_f32(u.w);
    return r;
}
__device__ __forceinline__ void st_bf16x8(uint16_t* p, const F8& r) {
    uint4 u;
    u.x = pack_bf16x2(r.v[0], r.v[1]); u.y = pack_bf16x2(r.v[2], r.v[3]);
    u.z = pack_bf16x2(r.v[4], r.v[5]); u.w = pack_bf16x2(r.v[6], r.v[7]);
    *reinterpret_cast<uint4*>(p) = u;
}
// guarded variants: CTAs are launched with a multiple of 32 threads; threads past C/8 are inactive
// (they load zeros, store nothing) but still take part in shuffles and barriers
__device__ __forceinline__ F8 ldz(bool active, const uint16_t* p);
__device__ __forceinline__ void stz(bool active, uint16_t* p, const F8& r);
__device__ __forceinline__ F8 zero8() {
    F8 r;
#pragma unroll
    for (int e = 0; e < 8; e++) r.v[e] = 0.f;
    return r;
}
__device__ __forceinline__ F8 ldz(bool active, const uint16_t* p) { return active ? ld_bf16x8(p) : zero8(); }
__device__ __forceinline__ void stz(bool active, uint16_t* p, const F8& r) {
    if (active) st_bf16x8(p, r);
}
// raw (still packed bf16) variants used by the software-pipelined row loops: half the registers of an F8
__device__ __forceinline__ uint4 ldraw(bool active, const uint16_t* p) {
    return active ? *reinterpret_cast<const uint4*>(p) : make_uint4(0u, 0u, 0u, 0u);
}
__device__ __forceinline__ F8 f8(const uint4 u) {
    F8 r;
    r.v[0] = bf16lo_to_f32(u.x); r.v[1] = bf16hi_to_f32(u.x);
    r.v[2] = bf16lo_to_f32(u.y); r.v[3] = bf16hi_to_f32(u.y);
    r.v[4] = bf16lo_to_f32(u.z); r.v[5] = bf16hi_to_f32(u.z);
    r.v[6] = bf16lo_to_f32(u.w); r.v[7] = bf16hi_to_f32(u.w);
    return r;
}
__host__ __device__ inline int row_threads(int C) { return ((C / 8 + 31) / 32) * 32; }
// round-trip through bf16: what a bf16 eager op leaves in memory
__device__ __forceinline__ float rb(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// Packed bf16 arithmetic with one rounding per operation — exactly what an eager bf16 op does to two bf16 operands (the fp32
// sum / product of two bf16 values is exact or rounds the same way), two elements per instruction.
__device__ __forceinline__ uint32_t bf2_sub(uint32_t a, uint32_t b) { uint32_t r; asm("sub.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t bf2_mul(uint32_t a, uint32_t b) { uint32_t r; asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t bf2_add(uint32_t a, uint32_t b) { uint32_t r; asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// Block-wide sum of NV (<= 4) values; `red` is [2][4][32] floats of shared memory, `phase` toggles per call so
// a single __syncthreads per reduction is enough.
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* red, int& phase, int nwarps) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    static_assert(NV <= 4, "block_sum: at most 4 values");
    float* buf = red + phase * 4 * 32;
#pragma unroll
    for (int k = 0; k < NV; k++) {
        const float s = warp_sum(v[k]);
        if (lane == 0) buf[k * 32 + warp] = s;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; k++) {
        float s = 0.f;
        for (int w = 0; w < nwarps; w++) s += buf[k * 32 + w];
        v[k] = s;
    }
    phase ^= 1;
}

}  // namespace vrwkv
