// wkv7_fwd.cuh — shared constants and argument structs of the WKV7 kernels (wkv7_fwd2.cuh / wkv7_bwd2.cuh).
#pragma once
#include "common.cuh"

namespace vrwkv {

constexpr int WKV_N = 64;   // head size (v7.00/src/model.py:69)
constexpr int WKV_TC = 16;  // steps per chunk == reference checkpoint interval (_CHUNK_LEN_, model.py:41)

struct Wkv7FwdArgs {
    int B, T, H;
    uint16_t* y;
    float* s;                // may be null
    float* sa;               // may be null
    const float* state_in;   // may be null (zeros)
    float* state_out;        // may be null
};

struct Wkv7BwdArgs {
    int B, T, H;
    const float* s;
    uint16_t *dw, *dq, *dk, *dv, *da, *db;
    // segment-parallel mode (wkv7_chunk_dstate.cuh provides the boundary values): blockIdx.z = segment of `span`
    // 16-step chunks, ds_in[b][h][segment][i][j] = dL/dS at the END of that segment (the last segment starts from 0)
    const float* ds_in = nullptr;
    int span = 0;  // 0: one CTA walks the whole sequence
};

}  // namespace vrwkv
