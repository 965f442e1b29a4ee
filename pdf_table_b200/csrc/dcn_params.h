// Launch parameters of dcn_fused_tcgen05 (dcn_fused.cuh): modulated deformable 3x3 convolution + bias + ReLU in one kernel.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace dv {

constexpr int kDcnTH = 8, kDcnTW = 16;  // one 128-row MMA tile = an 8 x 16 pixel patch (sampling locality in L1)

struct DcnParams {
    CUtensorMap tmB;    // W [cout, 9 * C] fp16 K-major (K = tap * C + c), box {64, cout}, SWIZZLE_128B
    const __half* in;   // NHWC fp16, pixel stride ldi
    const float* om;    // [N*H*W, 32] fp32: columns 2k, 2k+1 = (dy, dx) of tap k, 18 + k = mask logit
    const float* bias;  // [cout]
    __half* out;        // NHWC fp16, pixel stride ldo
    int N, H, W, C, ldi, ldo, cout;
    int tiles_x, tiles_y, n_tiles;
    int stages;         // depth of the A / B chunk rings
    int act;
    int dbg;            // tuning aid (DV_DCN_DEBUG): 1 = no gather loads (zeros), 2 = no blend math; results are then wrong
};

}  // namespace dv
