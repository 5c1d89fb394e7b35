// umma.cuh — tcgen05 (5th-generation tensor core) helpers shared by the GEMM and the chunked WKV7 kernels:
// shared-memory matrix descriptors for the canonical SWIZZLE_128B layouts, instruction descriptors, MMA issue for
// kind::f16 (bf16) and kind::tf32, commit, TMEM allocation and TMEM <-> register moves.
#pragma once
#include "common.cuh"

namespace vrwkv {

// ---- shared-memory matrix descriptors (SWIZZLE_128B, descriptor version 1) ----
// K-major: rows of 128 B (64 bf16 / 32 tf32 along K), 8-row groups 1024 B apart (SBO); LBO unused (=1).
__device__ __forceinline__ uint64_t umma_desc_sw128(const void* smem_tile) {
    const uint64_t addr = (uint64_t)((smem_u32(smem_tile) & 0x3FFFFu) >> 4);
    return addr | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major operands of 32-bit types must use the SWIZZLE_128B_BASE32B layout (descriptor layout type 1): lines of 128 B
// (32 tf32 contiguous along M/N), 4 K-lines per 512-B swizzle atom, 32-byte granules XOR-ed with (K-line & 3).
// `lbo_bytes` is the distance between consecutive 32-wide M/N blocks, `sbo_bytes` between consecutive groups of 4 K-lines.
__device__ __forceinline__ uint64_t umma_desc_mn_tf32(const void* smem_tile, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    const uint64_t addr = (uint64_t)((smem_u32(smem_tile) & 0x3FFFFu) >> 4);
    return addr | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (1ull << 61);
}
// byte offset of element (k_line, mn) inside one 32-wide M/N block of that layout
__device__ __forceinline__ uint32_t sw32_off(int k_line, int mn) {
    return (uint32_t)k_line * 128u + ((((uint32_t)mn >> 3) ^ ((uint32_t)k_line & 3u)) << 5) + ((uint32_t)mn & 7u) * 4u;
}
// advance a descriptor's start address by `bytes` (must keep the 16-byte granularity)
__device__ __forceinline__ uint64_t umma_desc_advance(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }

// ---- instruction descriptors: fp32 accumulate, dense; a_major / b_major: 0 = K-major, 1 = MN-major ----
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N, int a_mn = 0, int b_mn = 0) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_c),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_c),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A operand read from TMEM (lane = row, 32-bit column = k): `tmem_a` addresses the first of the 8 k-columns
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_c, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_c),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }

// ---- TMEM <-> registers: the calling warp's 32 lanes x 32 consecutive fp32 columns ----
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
    __syncwarp();  // .sync.aligned: the whole warp must arrive converged
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    tmem_ld32_nowait(taddr, r);
    tmem_ld_wait();
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    __syncwarp();
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// TMEM allocation: warp-wide, NCOLS a power of two in [32, 512]; the base address lands in *smem_slot
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}

// round-to-nearest fp32 -> tf32 (10-bit mantissa), returned as fp32 bits
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// byte offset of element (row, k) inside a SWIZZLE_128B tile whose rows hold 32 fp32 (one 128-byte line per row)
__device__ __forceinline__ uint32_t sw128_off(int row, int k) {
    return (uint32_t)row * 128u + ((((uint32_t)k >> 2) ^ ((uint32_t)row & 7u)) << 4) + ((uint32_t)k & 3u) * 4u;
}

}  // namespace vrwkv

namespace vrwkv {
// 16-column variants of the TMEM <-> register moves
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
    __syncwarp();
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    tmem_ld16_nowait(taddr, r);
    tmem_ld_wait();
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    __syncwarp();
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
// round-to-nearest for tensor-core operands: the tf32 datapath ignores the 13 low mantissa bits, so adding half a
// tf32 ulp to the fp32 bit pattern makes that truncation a rounding (one integer add instead of cvt.rna's sequence)
__device__ __forceinline__ float rt32(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }
__device__ __forceinline__ float4 rt32(float4 v) { return make_float4(rt32(v.x), rt32(v.y), rt32(v.z), rt32(v.w)); }
}  // namespace vrwkv

// ---- bf16 operand tiles (round 2): 64 rows x 64 bf16 = 8 KB, SWIZZLE_128B.  The same bytes serve as a K-major operand
// (row = M/N index, the 64 columns = K) and as an MN-major operand (row = K line, the 64 columns = M/N): only the
// instruction descriptor's major bit and the descriptor's K-step differ.  TMA boxes of [64 rows][64 bf16] loaded with
// CU_TENSOR_MAP_SWIZZLE_128B land in exactly this layout.
namespace vrwkv {
constexpr uint32_t BT_BYTES = 8192;  // one bf16 tile
__device__ __forceinline__ uint32_t bt_off(int row, int col) {  // byte offset of element (row, col)
    return (uint32_t)row * 128u + ((((uint32_t)col >> 3) ^ ((uint32_t)row & 7u)) << 4) + ((uint32_t)col & 7u) * 2u;
}
// descriptors from b4 = (1024-aligned struct base) >> 4 and a byte offset (cheap: one add per descriptor)
__device__ __forceinline__ uint64_t bdesc_k(uint32_t b4, uint32_t off_bytes) {  // K-major; K-step of 16 elements = +32 bytes
    return ((uint64_t)(64u | (1u << 14) | (2u << 29)) << 32) | (uint64_t)(b4 + (off_bytes >> 4) + (1u << 16));
}
// MN-major; K-step of 16 K-lines = +2048 bytes; `lbo` = byte distance between consecutive 64-wide M/N blocks
__device__ __forceinline__ uint64_t bdesc_mn(uint32_t b4, uint32_t off_bytes, uint32_t lbo = BT_BYTES) {
    return ((uint64_t)(64u | (1u << 14) | (2u << 29)) << 32) | (uint64_t)(b4 + (off_bytes >> 4) + ((lbo >> 4) << 16));
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mj(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
// A operand from TMEM (bf16 pairs packed in 32-bit columns: column c holds K elements 2c, 2c+1), B from shared memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_c, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_c),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 3-way bf16 split of an fp32 value: x = s0 + s1 + s2 to within 2^-25 |x| (each part round-to-nearest)
__device__ __forceinline__ void split3(float x, uint16_t& s0, uint16_t& s1, uint16_t& s2) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(h0);
    const __nv_bfloat16 h1 = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(h1);
    s0 = __bfloat16_as_ushort(h0); s1 = __bfloat16_as_ushort(h1); s2 = __bfloat16_as_ushort(__float2bfloat16_rn(r2));
}
// the same for a pair (packed cvt): p0/p1/p2 hold {lo: x, hi: y}
__device__ __forceinline__ void split3x2(float x, float y, uint32_t& p0, uint32_t& p1, uint32_t& p2) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p0) : "f"(y), "f"(x));
    const float rx = x - __uint_as_float(p0 << 16), ry = y - __uint_as_float(p0 & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p1) : "f"(ry), "f"(rx));
    const float qx = rx - __uint_as_float(p1 << 16), qy = ry - __uint_as_float(p1 & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p2) : "f"(qy), "f"(qx));
}
}  // namespace vrwkv
