// Parameter block and mode enums of conv_igemm_tcgen05 (kernel in igemm.cuh, planner in igemm_host.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace dv {

enum AMode : int {
    A_FLAT = 0,      // A is a plain [M, K] matrix (1x1 stride-1 conv, linear layers)
    A_PATCH = 1,     // stride-1 KHxKW conv, tensor map dims {C, W, H, N, 1}
    A_PATCH_S2 = 2,  // stride-2 conv, parity-split view dims {2C, W/2, 2, H/2, N}
    A_STEM = 3,      // 7x7 conv on the zero-bordered image, overlapping-window view
    A_HALO = 4       // 3x3 stride-1 pad-1 conv: one (16+2)x(8+2) halo patch per channel block is staged ONCE in shared
                     // memory (TMA box {BK, 16, 18}) and the nine taps are shifted UMMA descriptors into it; only the
                     // per-tap weight tiles stream through the stage ring
};
enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_SIGMOID = 3, ACT_HSWISH = 4, ACT_SWISH = 5 };
enum ResMode : int { RES_NONE = 0, RES_SAME = 1, RES_UP2 = 2 };
enum OutMode : int { OUT_NHWC = 0, OUT_REPL = 1, OUT_SHUF2 = 2 };

struct IGemmParams {
    CUtensorMap tmA;
    CUtensorMap tmB;
    CUtensorMap tmD;       // tma_store: the fp16 NHWC output (slice) as {Cout, M} (A_FLAT) or {Cout, Wo, Ho, N}, box = one epilogue
                           // warp's 32 rows x 32 columns, SWIZZLE_64B
    int tma_store;         // 1: the epilogue stages its 32 x 32 fp16 tiles in shared memory and stores them with
                           // cp.async.bulk.tensor (full-line writes; per-thread row stores cost 32 L1 wavefronts per instruction)
    int stg_off;           // byte offset of the staging buffers (8 warps x 2 x 2 KB) from the aligned shared-memory base
    int st_bw;             // PATCH modes: the store box covers st_bw columns x 32 / st_bw rows of the tile's pixel patch
    const int4* kb_delta;  // [num_kb] per-k-block coordinate deltas (see producer)
    int mode;
    int num_kb;
    int BK;         // 64 / 32 / 16 fp16 per k-block (row_bytes = 2*BK = swizzle span)
    int num_stages;
    int halo_stages;  // A_HALO: depth of the halo-patch ring (each 18*16*2*BK bytes)
    int acc_stages;   // TMEM accumulator ring depth: 2 (n-tile > 128 columns), 4 (> 64) or 8
    int b_resident;   // A_HALO: all 9 * ncb weight tiles stay in shared memory for the whole kernel (one n-tile, small filter)
    int ncb;          // A_HALO: channel blocks (Cin_pad / BK); k-block index of (tap, cb) = tap * ncb + cb
    int desc_base_off;  // A_HALO: set the UMMA descriptor base-offset field from the start address
    int M;          // A_FLAT: number of rows
    int Nimg, Ho, Wo;
    int TH, TW, tiles_x, tiles_y;
    int m_tiles, n_tiles, BLOCK_N, Cout;
    // epilogue
    const float* bias;  // padded to n_tiles*BLOCK_N, or nullptr
    const void* res;  // fp16 (or fp32 when the kernel is instantiated with RES_F32)
    int res_mode, res_ld;
    int res_mod;      // > 0: residual row = output row % res_mod (broadcast table, e.g. position embeddings)
    int res_red;      // fp32 stream updated in place (res == out, same layout): the epilogue issues red.global.add.v4.f32
                      // instead of load / add / store, so no thread waits for the residual (same fp32 sum, formed at L2)
    int n_inner;      // tile order: a CTA walks all n-tiles of its m-tiles (required by the arg-max epilogue)
    int32_t* arg_out;  // ARGMAX epilogue: per-row arg-max over all Cout columns
    float* max_out;    // ARGMAX epilogue: per-row maximum (may be null)
    int act;
    int out_mode, out_ld, out_coff, rep, out_f32;
    void* out;
    // A_FLAT only: number of valid rows read from device memory at kernel start (<= M, the planned capacity);
    // lets data-dependent row counts (selected table cells) run without a host round trip.  nullptr = M.
    const int* m_dyn;
    // fp16 output only, > 0: also store lo = fp16(v - fp32(fp16(v))) at column + split_off, so the consumer
    // GEMM can run the 3-term split-fp16 product (plan_linear_split) at ~fp32 accuracy.
    int split_off;
    // fp16 residual stored as a split pair: > 0 = add the lo half found res_split_off columns after the hi half
    int res_split_off;
    // post_affine != 0: y = act(x) * post_scale + post_bias (the LearnableAffineBlock that follows the activation in
    // PPLCNetV3's rep layers; it cannot be folded into the next layer across a zero-padded depthwise conv or an SE gate)
    int post_affine;
    float post_scale, post_bias;
};

}  // namespace dv
