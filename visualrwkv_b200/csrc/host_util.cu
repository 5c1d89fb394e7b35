// host_util.cu — error reporting + TMA descriptor encoding (driver entry point fetched at run time,
// so the library links against cudart only).
#include "host_util.h"

#include <cudaTypedefs.h>

#include <cstdarg>
#include <cstdio>
#include <atomic>
#include <mutex>

static thread_local char g_err[512] = "";

int vrwkv_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char* vrwkv_last_error(void) { return g_err; }
extern "C" int vrwkv_version(void) { return 100; }

static std::atomic<unsigned long long> g_launches{0};
void vrwkv_count_launch(int n) { g_launches.fetch_add((unsigned long long)n); }
extern "C" unsigned long long vrwkv_launch_count(void) { return g_launches.load(); }

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    });
    return fn;
}

int vrwkv_encode_2d(CUtensorMap* m, const void* base, CUtensorMapDataType dt, int elem_bytes, uint64_t cols,
                    uint64_t rows, uint64_t row_stride_bytes, uint32_t box_cols, uint32_t box_rows,
                    CUtensorMapSwizzle swizzle) {
    // the driver call needs a current context on THIS thread (autograd worker threads may not have touched the
    // runtime yet): a no-op runtime call binds the primary context of the current device
    static thread_local bool ctx_ready = false;
    if (!ctx_ready) {
        cudaError_t e = cudaFree(nullptr);
        if (e != cudaSuccess) return vrwkv_fail(VRWKV_ECUDA, "CUDA context init failed: %s", cudaGetErrorString(e));
        ctx_ready = true;
    }
    auto enc = get_encode();
    if (!enc) return vrwkv_fail(VRWKV_ECUDA, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return vrwkv_fail(VRWKV_ECUDA, "cuTensorMapEncodeTiled failed (CUresult %d; cols=%llu rows=%llu stride=%llu box=%ux%u)",
                          (int)r, (unsigned long long)cols, (unsigned long long)rows,
                          (unsigned long long)row_stride_bytes, box_cols, box_rows);
    (void)elem_bytes;
    return VRWKV_OK;
}
