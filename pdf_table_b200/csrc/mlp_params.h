// Launch parameters of mlp_fused_tcgen05 (mlp_fused.cuh): x += W2 * GELU(W1 * h + b1) + b2 on one 128-token tile at a time.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace dv {

struct MlpParams {
    CUtensorMap tmA;   // h   [M, C]   fp16, box {64, 128}, SWIZZLE_128B (columns >= C are zero-filled by TMA)
    CUtensorMap tmW1;  // W1  [4C, C]  fp16, box {64, 64}
    CUtensorMap tmW2;  // W2  [C, 4C]  fp16, box {64, C}
    const float* b1;   // [4C]
    const float* b2;   // [C]
    float* x;          // [M, C] fp32 residual stream, updated in place
    int M, m_tiles;
    int dbg;           // tuning aid (DV_MLP_DEBUG): 1 = skip the GELU math (pipeline floor), results are then wrong
};

}  // namespace dv
