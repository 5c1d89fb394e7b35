// wkv7_chunk_common.cuh — pieces shared by the chunked WKV7 kernels (wkv7_chunk_fwd.cuh, wkv7_chunk_dstate.cuh).
#pragma once
#include "common.cuh"
#include "umma.cuh"
#include "wkv7_fwd.cuh"

namespace vrwkv {

constexpr int CK_L = 64;         // steps per chunk
constexpr int CK_THREADS = 512;  // 16 warps: TMEM lane quadrant = warp & 3, column slice = warp >> 2

__device__ int g_chunk_domain_err = 0;    // set when a chunk's accumulated decay leaves the fp32-safe range
__device__ float* g_chunk_dbg = nullptr;  // development aid: CTA (0,0) records per-phase clocks of chunk 1

// Shared-memory matrix descriptors built from the operand's byte offset inside the (1024-byte aligned) shared-memory
// struct and b4 = (struct base address) >> 4: with compile-time offsets each descriptor is one integer add, which
// matters because a single thread issues every tcgen05.mma of a phase (~100 per chunk in the backward).
__device__ __forceinline__ uint64_t desc_km(uint32_t b4, uint32_t off_bytes) {  // K-major SWIZZLE_128B, 8-row groups 1024 B apart
    return ((uint64_t)(64u | (1u << 14) | (2u << 29)) << 32) | (uint64_t)(b4 + (off_bytes >> 4) + (1u << 16));
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t b4, uint32_t off_bytes, uint32_t lbo, uint32_t sbo = 512) {  // MN-major BASE32B
    return ((uint64_t)((sbo >> 4) | (1u << 14) | (1u << 29)) << 32) | (uint64_t)(b4 + (off_bytes >> 4) + ((lbo >> 4) << 16));
}

// address of A[t][s] (fp32) in the chunk-swizzled 64x64 array `aab`: rows of 256 B, 16-byte chunks XOR (t & 7)
__device__ __forceinline__ const float4* aab_chunk(const uint8_t* aab, int t, int chunk) {
    return reinterpret_cast<const float4*>(aab + t * 256 + ((chunk ^ (t & 7)) << 4));
}

// Tinv = (I - A)^-1 for a strictly lower triangular 64x64 A (fp32, in `aab`), computed in fp32 by all 512 threads:
// the four 16x16 diagonal blocks by substitution (one column per thread), then the coupling blocks of the two 32x32
// blocks and of the 64x64 matrix as X_lower = X_b (A_lower X_a).  The result is written tf32-rounded straight into
// the caller's operand layout through at(t, s) (byte address of element (t, s); vector loads of 2 / 4 consecutive s
// must be legal there).  The buffer must have been zeroed (the upper-right blocks are never written).  `esc` is a
// 4 KB scratch.  Contains __syncthreads(): call from all threads.
template <class At>
__device__ __forceinline__ void chunk_tri_inverse(const uint8_t* aab, float* esc, const int tid, At at) {
    // ---- diagonal blocks ----
    if (tid < 64) {
        const int n = tid >> 4, cc = tid & 15;
        float x[16];
#pragma unroll
        for (int t = 0; t < 16; t++) {
            float a0 = (t == cc) ? 1.f : 0.f, a1 = 0.f;
#pragma unroll
            for (int c4 = 0; c4 < (t + 3) / 4; c4++) {
                const float4 m = *aab_chunk(aab, 16 * n + t, 4 * n + c4);
                const float mm[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int s = 4 * c4 + e;
                    if (s < t) {
                        if (s & 1) a1 = fmaf(mm[e], x[s], a1);
                        else a0 = fmaf(mm[e], x[s], a0);
                    }
                }
            }
            x[t] = a0 + a1;
            *reinterpret_cast<float*>(at(16 * n + t, 16 * n + cc)) = rt32(x[t]);
        }
    }
    __syncthreads();
    // ---- level 1: the (1,0) block of each 32x32 diagonal block: X10 = D1 (A10 D0) ----
    {
        const int m = tid >> 8, t = (tid >> 4) & 15, c = tid & 15, o = 32 * m;
        float e0 = 0.f, e1 = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < 4; c4++) {
            const float4 a = *aab_chunk(aab, o + 16 + t, 8 * m + c4);
            e0 = fmaf(a.x, *reinterpret_cast<const float*>(at(o + 4 * c4 + 0, o + c)), e0);
            e1 = fmaf(a.y, *reinterpret_cast<const float*>(at(o + 4 * c4 + 1, o + c)), e1);
            e0 = fmaf(a.z, *reinterpret_cast<const float*>(at(o + 4 * c4 + 2, o + c)), e0);
            e1 = fmaf(a.w, *reinterpret_cast<const float*>(at(o + 4 * c4 + 3, o + c)), e1);
        }
        esc[m * 256 + t * 16 + c] = e0 + e1;
    }
    __syncthreads();
    {
        const int m = tid >> 8, t = (tid >> 4) & 15, c = tid & 15, o = 32 * m;
        float e0 = 0.f, e1 = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < 4; c4++) {
            const float4 d = *reinterpret_cast<const float4*>(at(o + 16 + t, o + 16 + 4 * c4));
            e0 = fmaf(d.x, esc[m * 256 + (4 * c4 + 0) * 16 + c], e0);
            e1 = fmaf(d.y, esc[m * 256 + (4 * c4 + 1) * 16 + c], e1);
            e0 = fmaf(d.z, esc[m * 256 + (4 * c4 + 2) * 16 + c], e0);
            e1 = fmaf(d.w, esc[m * 256 + (4 * c4 + 3) * 16 + c], e1);
        }
        *reinterpret_cast<float*>(at(o + 16 + t, o + c)) = rt32(e0 + e1);
    }
    __syncthreads();
    // ---- level 2: the lower-left 32x32 block: Xc = Xb (Ac Xa) ----
    {
        const int t = tid >> 4, c0 = 2 * (tid & 15);
        float e0 = 0.f, e1 = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < 8; c4++) {
            const float4 m = *aab_chunk(aab, 32 + t, c4);
            const float mm[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const float2 xa = *reinterpret_cast<const float2*>(at(4 * c4 + e, c0));
                e0 = fmaf(mm[e], xa.x, e0);
                e1 = fmaf(mm[e], xa.y, e1);
            }
        }
        *reinterpret_cast<float2*>(&esc[t * 32 + c0]) = make_float2(e0, e1);
    }
    __syncthreads();
    {
        const int t = tid >> 4, c0 = 2 * (tid & 15);
        float e0 = 0.f, e1 = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < 8; c4++) {
            const float4 m = *reinterpret_cast<const float4*>(at(32 + t, 32 + 4 * c4));
            const float mm[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const float2 ev = *reinterpret_cast<const float2*>(&esc[(4 * c4 + e) * 32 + c0]);
                e0 = fmaf(mm[e], ev.x, e0);
                e1 = fmaf(mm[e], ev.y, e1);
            }
        }
        *reinterpret_cast<float2*>(at(32 + t, c0)) = make_float2(rt32(e0), rt32(e1));
    }
}

}  // namespace vrwkv
