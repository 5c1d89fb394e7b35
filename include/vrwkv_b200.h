/*
 * vrwkv_b200.h — C ABI of the B200-native VisualRWKV hot path (libvrwkv_b200.so).
 *
 * Plain pointers and sizes only; no torch types.  All pointers are DEVICE pointers on the
 * current CUDA device unless stated otherwise.  `stream` is a cudaStream_t passed as void*
 * (NULL = legacy default stream, which is what the reference launches on,
 * VisualRWKV-v7/v7.00/cuda/wkv7_cuda.cu:133,137).
 *
 * Every entry point returns 0 on success or a negative VRWKV_E* code; the message of the last
 * failure on the calling thread is available from vrwkv_last_error().  Nothing here ever
 * falls back to a CPU path.
 *
 * "bf16" buffers are raw uint16_t storage of __nv_bfloat16.
 */
#ifndef VRWKV_B200_H
#define VRWKV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VRWKV_OK 0
#define VRWKV_EINVAL (-1)  /* bad shape / alignment / null pointer */
#define VRWKV_ECUDA (-2)   /* CUDA runtime or driver error */
#define VRWKV_EUNSUP (-3)  /* shape outside what the kernels support (e.g. head size != 64) */

#define VRWKV_HEAD_SIZE 64 /* v7.00/src/model.py:69 hard-wires 64 */
#define VRWKV_CHUNK_LEN 16 /* v7.00/src/model.py:41 CHUNK_LEN */

const char* vrwkv_last_error(void);
int vrwkv_version(void);
/* Number of CUDA kernels this library has launched in this process (all entry points). */
unsigned long long vrwkv_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * WKV7 recurrence — replaces cuda_forward / cuda_backward
 * (VisualRWKV-v7/v7.00/cuda/wkv7_cuda.cu:132-138; bound by wkv7_op.cpp:7-19 as
 * torch.ops.wind_backstepping.{forward,backward}).
 *
 * w,q,k,v,a,b,y,dy,d*: bf16 [B,T,H,64] contiguous.  Argument order is the KERNEL order
 * (schema names w,q,k,v,z,a == kernel names w,q,k,v,a,b == python (w, r, k, v, -kk, kk*a)).
 * s : f32 [B,H,T/16,64,64]  state checkpoints, TRANSPOSED: s[b,h,c,j,i] = S_ij after step
 *     16c+15 (wkv7_cuda.cu:44-50).
 * sa: f32 [B,T,H,64]        sa[b,t,h,i] = sum_j a[t,j] S_ij (state BEFORE step t) (:27-32).
 * Requirements: T % 16 == 0 (model.py:49), all pointers 16-byte aligned, same device.
 * Caller allocates every output (model.py:52-54,63); the op keeps no state between calls.
 * ------------------------------------------------------------------------------------------ */
int vrwkv_wkv7_forward(int B, int T, int H, const uint16_t* w, const uint16_t* q,
                       const uint16_t* k, const uint16_t* v, const uint16_t* a,
                       const uint16_t* b, uint16_t* y, float* s, float* sa, void* stream);

int vrwkv_wkv7_backward(int B, int T, int H, const uint16_t* w, const uint16_t* q,
                        const uint16_t* k, const uint16_t* v, const uint16_t* a,
                        const uint16_t* b, const uint16_t* dy, const float* s, const float* sa,
                        uint16_t* dw, uint16_t* dq, uint16_t* dk, uint16_t* dv, uint16_t* da,
                        uint16_t* db, void* stream);

/* Stateful forward (no reference counterpart for x070): as vrwkv_wkv7_forward
 * but the state starts from state_in (f32 [B,H,64,64], S_ij row-major; NULL = zeros) and the
 * final state is written to state_out (may alias state_in; NULL = not written).  s / sa may be
 * NULL (not written).  T may be any positive value when s == NULL. */
int vrwkv_wkv7_forward_state(int B, int T, int H, const uint16_t* w, const uint16_t* q,
                             const uint16_t* k, const uint16_t* v, const uint16_t* a,
                             const uint16_t* b, uint16_t* y, float* s, float* sa,
                             const float* state_in, float* state_out, void* stream);

/* Kernel-variant selection for benchmarking (0 = default heuristic). Thread-safe, process-wide. */
int vrwkv_wkv7_set_variant(int fwd_variant, int bwd_variant);

#ifdef __cplusplus
}
#endif
#endif /* VRWKV_B200_H */
