// Parameter block of conv_win_tcgen05 (kernel in win_conv.cuh, planner in igemm_host.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

namespace dv {

struct WinConvParams {
    CUtensorMap tmB;        // weights [Cout][KR*64] K-major, box {64, BLOCK_N}
    const __half* in;       // zero-bordered image [N][Hp][Wp][cpp] (cpp = 8 or 16 channels per pixel)
    int Hp, Wp, cpp, stride;
    int KR;                 // filter rows = k-blocks (K = 64 each: 128 B / (2*cpp) pixels x cpp channels)
    int Nimg, Ho, Wo, TH, TW, tiles_x, tiles_y, m_tiles;
    int BLOCK_N, Cout, num_stages, act;
    const float* bias;      // padded to a multiple of 256
    __half* out;            // [N][oHp][oWp][out_ld], written at (+opad, +opad)
    int oHp, oWp, opad, out_ld;
    // patch mode (stride 1): the (TH+KR-1) x 16-pixel input patch of a 16 x 8 tile is staged ONCE, as `planes` planes of
    // [row][pixel][8 channels]; the A operand of filter row r / pixel pair j is an UN-swizzled K-major descriptor into it with
    // LBO = 16 B (next pixel = next 8-element K chunk) and rows 16 B apart (next output pixel): overlapping windows, no copies
    int patch, planes, PR;
};

}  // namespace dv
