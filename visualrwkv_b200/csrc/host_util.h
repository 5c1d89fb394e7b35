// host_util.h — host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vrwkv_b200.h"

// Records a printf-style message for vrwkv_last_error() and returns `code`.
int vrwkv_fail(int code, const char* fmt, ...);

#define VRWKV_CUDA(expr)                                                                              \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess)                                                                        \
            return vrwkv_fail(VRWKV_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                              __FILE__, __LINE__);                                                    \
    } while (0)

// cuTensorMapEncodeTiled for a row-major 2-D matrix [rows, cols]; box = [box_rows, box_cols].
int vrwkv_encode_2d(CUtensorMap* m, const void* base, CUtensorMapDataType dt, int elem_bytes, uint64_t cols,
                    uint64_t rows, uint64_t row_stride_bytes, uint32_t box_cols, uint32_t box_rows,
                    CUtensorMapSwizzle swizzle);

// Every kernel launch made by this library is counted (bench.py reports it as gpu_launches).
void vrwkv_count_launch(int n);
