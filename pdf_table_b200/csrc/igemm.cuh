// conv_igemm_tcgen05 -- the engine's one tensor-core kernel.
//
// Implicit-GEMM convolution / linear layer for NHWC fp16 activations:
//     D[m, n] = sum_k A[m, k] * W[n, k]        m = output pixel, n = output channel,
//                                              k = (tap r,s ; input channel c)
// * A tiles (128 output pixels x BK input channels of one filter tap) are fetched by TMA straight from
//   the NHWC activation tensor: the 128 pixels are a TH x TW patch of one image, so the tile of tap
//   (r,s) is the same box shifted by (r-pad, s-pad); out-of-image taps are zero-filled by the TMA
//   unit, i.e. padding costs nothing and no im2col buffer ever exists in HBM.
// * W tiles (BLOCK_N x BK, K-major) are fetched by TMA from the packed weight matrix.
// * tcgen05.mma (kind::f16, M=128, N=BLOCK_N, K=16) accumulates in TMEM (fp32), issued by one thread.
// * The epilogue (8 warps, two per TMEM lane quadrant) reads TMEM with tcgen05.ld and fuses bias(+folded BN), residual add
//   (same resolution or nearest-2x-upsampled), activation, and the store pattern (plain NHWC slice,
//   s x s nearest replication into a concat slice, or 2x2 pixel-shuffle for ConvTranspose 2x2 s2).
// * Persistent CTAs (grid = min(tiles, #SM)), a num_stages-deep smem ring between TMA and MMA, and a
//   2-deep TMEM accumulator ring between MMA and epilogue so the epilogue of tile i overlaps the
//   main loop of tile i+1.
#pragma once
#include "igemm_params.h"
#include "ptx.cuh"

namespace dv {

static constexpr int kIGemmThreads = 320;  // warp 0: TMA, warp 1: MMA, warps 2..9: epilogue (two per TMEM lane quadrant)
static constexpr int kMaxKB = 224;  // k-blocks per tile whose coordinate deltas are staged in smem (3x3 x 512 channels x 3 split parts = 216)
static constexpr int kMaxHaloStages = 8;
static constexpr int kMaxAccStages = 8;  // TMEM accumulator ring: 512 columns / the n-tile's (power-of-two) width, 2..8 deep
static constexpr int kMaxStages = 32;  // TMA -> MMA ring depth.  Small-K layers (BK 16 / 32: 5-16 KB stages) are latency-bound on
                                       // bytes in flight: with the former cap of 8 a Cin=16 conv kept 40 KB per SM in flight
static constexpr int kBiasSmem = 2048;  // bias values staged in smem (layers with more padded columns read global)

template <int ACT>
__device__ __forceinline__ float apply_act(float x) {
    if constexpr (ACT == ACT_RELU) return fmaxf(x, 0.f);
    if constexpr (ACT == ACT_GELU) {
        // exact (erf) GELU = relu(x) - |x|/2 * erfc(|x|/sqrt2), erfc by Abramowitz-Stegun 7.1.26: one rcp, one ex2,
        // five FMAs and NO selects (erff costs ~28 issue slots per element, 9 of them FSELs, and this epilogue is
        // issue-bound).  |abs err| <= 3.4e-7 over [-12, 12] against float64, i.e. far below the fp16 rounding of the
        // value that is stored.
        const float ax = fabsf(x);
        float t, e;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.23164189f, ax, 1.f)));  // 0.3275911 / sqrt(2)
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"((x * -0.72134752f) * x));       // exp(-x^2 / 2)
        float p = fmaf(t, 1.061405429f, -1.453152027f);
        p = fmaf(p, t, 1.421413741f);
        p = fmaf(p, t, -0.284496736f);
        p = fmaf(p, t, 0.254829592f);
        return fmaf(-0.5f * ax, (p * t) * e, fmaxf(x, 0.f));
    }
    if constexpr (ACT == ACT_SIGMOID) return 1.f / (1.f + __expf(-x));
    if constexpr (ACT == ACT_HSWISH) return x * fminf(fmaxf(x + 3.f, 0.f), 6.f) * (1.f / 6.f);
    if constexpr (ACT == ACT_SWISH) return x / (1.f + __expf(-x));
    return x;
}

// The same GELU on two values at once with Blackwell's packed fp32 instructions (FFMA2 / FMUL2 / FADD2: one issue slot for
// two IEEE operations, so the result is bit-identical to apply_act<ACT_GELU> on each value): 8.5 instead of 16 issue
// slots per element in an epilogue that is issue-bound.  gelu_bias_x2 returns gelu(v0 + b0), gelu(v1 + b1).
namespace f32x2 {
__device__ __forceinline__ uint64_t pk(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t mul(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t add(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
}  // namespace f32x2
__device__ __forceinline__ void gelu_x2(float x0, float x1, float& y0, float& y1) {
    using namespace f32x2;
    const uint64_t X = pk(x0, x1), AX = pk(fabsf(x0), fabsf(x1));
    float u0, u1, t0, t1, s0, s1, e0, e1;
    upk(fma(pk(0.23164189f, 0.23164189f), AX, pk(1.f, 1.f)), u0, u1);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(u0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(u1));
    upk(mul(mul(X, pk(-0.72134752f, -0.72134752f)), X), s0, s1);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(s0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(s1));
    const uint64_t T = pk(t0, t1);
    uint64_t P = fma(T, pk(1.061405429f, 1.061405429f), pk(-1.453152027f, -1.453152027f));
    P = fma(P, T, pk(1.421413741f, 1.421413741f));
    P = fma(P, T, pk(-0.284496736f, -0.284496736f));
    P = fma(P, T, pk(0.254829592f, 0.254829592f));
    upk(fma(mul(AX, pk(-0.5f, -0.5f)), mul(mul(P, T), pk(e0, e1)), pk(fmaxf(x0, 0.f), fmaxf(x1, 0.f))), y0, y1);
}
__device__ __forceinline__ void gelu_bias_x2(float v0, float v1, float b0, float b1, float& y0, float& y1) {
    float x0, x1;
    f32x2::upk(f32x2::add(f32x2::pk(v0, v1), f32x2::pk(b0, b1)), x0, x1);
    gelu_x2(x0, x1, y0, y1);
}

// Shape contract enforced by the planner (igemm_host.cu): out_ld, out_coff, res_ld and Cout are multiples
// of 8 (fp16 out) / 4 (fp32 out), so the epilogue only ever issues 16-byte vector accesses, predicated
// per vector on the channel bound.  Keeping the epilogue this small matters: it is executed by only four
// warps and an earlier, fully generic version was instruction-fetch bound (profiles/r1_notes.md).
// Tile scheduler shared by the three roles.  Default: tiles are dealt round-robin (tile = m_tile * n_tiles +
// n_tile).  n_inner (classifier arg-max epilogue): a CTA owns whole m-tiles and walks all their n-tiles, so
// one epilogue thread sees every column of its row.
struct TileIter {
    int j = 0;
    int m_tiles;
    __device__ __forceinline__ explicit TileIter(int mt) : m_tiles(mt) {}
    __device__ __forceinline__ bool next(const IGemmParams& p, int& m_tile, int& n_tile) {
        if (p.n_inner) {
            const int q = j / p.n_tiles;
            m_tile = blockIdx.x + q * gridDim.x;
            n_tile = j - q * p.n_tiles;
        } else {
            const int tile = blockIdx.x + j * gridDim.x;
            m_tile = tile / p.n_tiles;
            n_tile = tile - m_tile * p.n_tiles;
        }
        ++j;
        return m_tile < m_tiles;
    }
};

// ACT: activation; OUT_F32: fp32 (else fp16) output; RES_F32: the residual operand is fp32 (the fp32
// residual stream of the ConvNeXt / ViT blocks, updated in place); ARGMAX: classifier epilogue that keeps a
// running (max, arg-max) per row over all n-tiles and writes ids instead of (optionally besides) logits.
template <int ACT, bool OUT_F32, bool RES_F32, bool ARGMAX>
__global__ void __launch_bounds__(kIGemmThreads, 1)
conv_igemm_tcgen05(const __grid_constant__ IGemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
    __shared__ __align__(8) uint64_t tfull_bar[kMaxAccStages];
    __shared__ __align__(8) uint64_t tempty_bar[kMaxAccStages];
    __shared__ __align__(8) uint64_t hfull_bar[kMaxHaloStages];
    __shared__ __align__(8) uint64_t hempty_bar[kMaxHaloStages];
    __shared__ __align__(8) uint64_t bres_bar;  // A_HALO with a resident filter: all weight tiles have landed
    __shared__ uint32_t tmem_base_smem;
    __shared__ int4 s_delta[kMaxKB];
    __shared__ __align__(16) float s_bias[kBiasSmem];  // the layer's bias, staged once (see the epilogue)
    __shared__ float s_argv[ARGMAX ? 128 : 1];         // arg-max epilogue: the second warp's candidate per row
    __shared__ int s_argi[ARGMAX ? 128 : 1];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t row_bytes = 2u * p.BK;
    const bool halo = p.mode == A_HALO;
    const uint32_t a_bytes = halo ? 0u : 128u * row_bytes;  // A_HALO: the ring holds weight tiles only
    const uint32_t b_bytes = static_cast<uint32_t>(p.BLOCK_N) * row_bytes;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    const uint32_t halo_bytes = 18u * 16u * row_bytes;
    const uint32_t ring_base = smem_base + (halo ? static_cast<uint32_t>(p.halo_stages) * halo_bytes : 0u);
    const int num_stages = p.num_stages;
    const int num_kb = p.num_kb;
    // accumulator ring: the epilogue of a tile (TMEM load, activation, staging, TMA store) is a ~2 us latency chain; with two
    // accumulators a small-K tile (one or two MMAs) waited for it -- narrow n-tiles get up to eight
    const int acc_stages = p.acc_stages;
    const uint32_t acc_stride = 512u / static_cast<uint32_t>(acc_stages);
    // data-dependent row count (A_FLAT): every role derives the same tile range from it
    const int M_rows = p.m_dyn != nullptr ? min(__ldg(p.m_dyn), p.M) : p.M;
    const int m_tiles_rt = p.m_dyn != nullptr ? (M_rows + 127) >> 7 : p.m_tiles;

    for (int i = threadIdx.x; i < num_kb; i += kIGemmThreads) s_delta[i] = __ldg(&p.kb_delta[i]);
    // ncu (profiles/r1g): the epilogue's first bias FADD of every chunk sat on the long scoreboard (an L1/L2 round trip
    // per 32 columns with only two warps per scheduler to hide it) -> stage the whole bias vector in shared memory.
    const bool bias_smem = p.bias != nullptr && p.n_tiles * p.BLOCK_N <= kBiasSmem;
    if (bias_smem)
        for (int i = threadIdx.x; i < p.n_tiles * p.BLOCK_N; i += kIGemmThreads) s_bias[i] = __ldg(p.bias + i);
    if (threadIdx.x == 0) {
        for (int i = 0; i < num_stages; ++i) {
            ptx::mbar_init(ptx::smem_u32(&full_bar[i]), 1);
            ptx::mbar_init(ptx::smem_u32(&empty_bar[i]), 1);
        }
        for (int i = 0; i < kMaxAccStages; ++i) {
            ptx::mbar_init(ptx::smem_u32(&tfull_bar[i]), 1);
            ptx::mbar_init(ptx::smem_u32(&tempty_bar[i]), 8);
        }
        for (int i = 0; i < kMaxHaloStages; ++i) {
            ptx::mbar_init(ptx::smem_u32(&hfull_bar[i]), 1);
            ptx::mbar_init(ptx::smem_u32(&hempty_bar[i]), 1);
        }
        ptx::mbar_init(ptx::smem_u32(&bres_bar), 1);
        ptx::fence_barrier_init();
        ptx::prefetch_tmap(&p.tmA);
        ptx::prefetch_tmap(&p.tmB);
    }
    if (warp == 1) {
        ptx::tmem_alloc(ptx::smem_u32(&tmem_base_smem), 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ===================== TMA producer (one thread) =====================
        // elect_one_sync (not `lane == 0`): the compiler then knows exactly one thread runs the role and emits the uniform-
        // datapath TMA / MMA instructions back to back instead of wrapping each in an ELECT / BRA.U.ANY loop
        if (ptx::elect_one_sync()) {
            int stage = 0, hstage = 0;
            uint32_t phase = 0, hphase = 0;
            const int tiles_per_img = p.tiles_x * p.tiles_y;
            const int mode = p.mode;
            const int BK = p.BK, BLOCK_N = p.BLOCK_N;
            if (halo && p.b_resident) {
                // the whole filter (9 taps x ncb channel blocks, one n-tile) is loaded once and stays: the per-tap weight stream
                // and its barrier round trips disappear, the ring budget goes to a deeper halo-patch ring
                const uint32_t bb = ptx::smem_u32(&bres_bar);
                ptx::mbar_expect_tx(bb, 9u * static_cast<uint32_t>(p.ncb) * b_bytes);
                for (int t = 0; t < 9 * p.ncb; ++t) ptx::tma_load_2d(ring_base + t * b_bytes, &p.tmB, bb, t * BK, 0);
            }
            TileIter it(m_tiles_rt);
            int m_tile, n_tile;
            while (it.next(p, m_tile, n_tile)) {
                int img = 0, y0 = 0, x0 = 0;
                if (mode != A_FLAT) {
                    img = m_tile / tiles_per_img;
                    const int t = m_tile - img * tiles_per_img;
                    const int ty = t / p.tiles_x;
                    y0 = ty * p.TH;
                    x0 = (t - ty * p.tiles_x) * p.TW;
                }
                if (halo) {
                    for (int cb = 0; cb < p.ncb; ++cb) {
                        ptx::mbar_wait(ptx::smem_u32(&hempty_bar[hstage]), hphase ^ 1u);
                        const uint32_t hb = ptx::smem_u32(&hfull_bar[hstage]);
                        ptx::mbar_expect_tx(hb, halo_bytes);
                        ptx::tma_load_5d(smem_base + hstage * halo_bytes, &p.tmA, hb, cb * BK, x0 - 1, y0 - 1, img, 0);
                        if (++hstage == p.halo_stages) { hstage = 0; hphase ^= 1u; }
                        if (p.b_resident) continue;
                        for (int tap = 0; tap < 9; ++tap) {
                            ptx::mbar_wait(ptx::smem_u32(&empty_bar[stage]), phase ^ 1u);
                            const uint32_t fb = ptx::smem_u32(&full_bar[stage]);
                            ptx::mbar_expect_tx(fb, b_bytes);
                            ptx::tma_load_2d(ring_base + stage * b_bytes, &p.tmB, fb, (tap * p.ncb + cb) * BK, n_tile * BLOCK_N);
                            if (++stage == num_stages) { stage = 0; phase ^= 1u; }
                        }
                    }
                    continue;
                }
                for (int kb = 0; kb < num_kb; ++kb) {
                    const int4 d = s_delta[kb];
                    ptx::mbar_wait(ptx::smem_u32(&empty_bar[stage]), phase ^ 1u);
                    const uint32_t fb = ptx::smem_u32(&full_bar[stage]);
                    ptx::mbar_expect_tx(fb, stage_bytes);
                    const uint32_t sa = smem_base + stage * stage_bytes;
                    const uint32_t sb = sa + a_bytes;
                    switch (mode) {
                        case A_FLAT: ptx::tma_load_5d(sa, &p.tmA, fb, d.x, m_tile * 128, 0, 0, 0); break;
                        case A_PATCH: ptx::tma_load_5d(sa, &p.tmA, fb, d.x, x0 + d.y, y0 + d.z, img, 0); break;
                        case A_PATCH_S2:
                            ptx::tma_load_5d(sa, &p.tmA, fb, d.x, x0 + d.y, d.z, y0 + d.w, img);
                            break;
                        default:  // A_STEM: dims {32, Wo, 7, Ho, N}; d.w = image offset of the lo copy (fp32x mode)
                            ptx::tma_load_5d(sa, &p.tmA, fb, 0, x0, d.z, y0, img + d.w);
                            break;
                    }
                    ptx::tma_load_2d(sb, &p.tmB, fb, kb * BK, n_tile * BLOCK_N);
                    if (++stage == num_stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (ptx::elect_one_sync()) {
            int stage = 0, hstage = 0;
            uint32_t phase = 0, hphase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            const uint32_t idesc = ptx::make_idesc_f16_m128(static_cast<uint32_t>(p.BLOCK_N));
            const int k_steps = p.BK >> 4;
            const bool b_res = halo && p.b_resident;
            const uint64_t b_res_desc = ptx::make_kmajor_desc(ring_base, row_bytes);
            if (b_res) {
                ptx::mbar_wait(ptx::smem_u32(&bres_bar), 0u);
                ptx::tc_fence_after();
            }
            TileIter it(m_tiles_rt);
            int m_tile, n_tile;
            while (it.next(p, m_tile, n_tile)) {
                ptx::mbar_wait(ptx::smem_u32(&tempty_bar[acc]), acc_phase ^ 1u);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc) * acc_stride;
                if (b_res) {
                    // Resident filter: no weight barriers, and every descriptor is the stage's base descriptor plus a constant --
                    // the issuing thread's own instruction stream is what bounds the small-N layers (ncu, profiles/r4y: 130
                    // scalar instructions per tap with run-time tap / 3, % 3 and descriptor packing made the thread, not the
                    // tensor pipe, TMA or the epilogue, the limit), so the tap loop is unrolled with compile-time patch offsets
                    // (shared-memory addresses are < 256 KB: the 14-bit start-address field never carries).
                    const uint32_t rb16 = row_bytes >> 4, bt16 = b_bytes >> 4;
                    // (two alternating accumulators were tried here: no change -- a chain of dependent N = 32 MMAs is not the limit)
                    uint32_t accum = 0u;
                    for (int cb = 0; cb < p.ncb; ++cb) {
                        ptx::mbar_wait(ptx::smem_u32(&hfull_bar[hstage]), hphase);
                        ptx::tc_fence_after();
                        const uint64_t a_base = ptx::make_kmajor_desc_sbo(smem_base + hstage * halo_bytes, row_bytes, 16u * row_bytes, 0);
                        uint64_t b_tap = b_res_desc + static_cast<uint64_t>(static_cast<uint32_t>(cb) * bt16);
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            const uint64_t a_tap = a_base + static_cast<uint64_t>(static_cast<uint32_t>((tap / 3) * 16 + tap % 3) * rb16);
                            for (int k = 0; k < k_steps; ++k) {
                                ptx::umma_f16_ss(d_tmem, a_tap + 2ull * k, b_tap + 2ull * k, idesc, accum);
                                accum = 1u;
                            }
                            b_tap += static_cast<uint64_t>(static_cast<uint32_t>(p.ncb) * bt16);
                        }
                        ptx::umma_commit(ptx::smem_u32(&hempty_bar[hstage]));  // patch free once its 9 x k_steps MMAs retire
                        if (++hstage == p.halo_stages) { hstage = 0; hphase ^= 1u; }
                    }
                    ptx::umma_commit(ptx::smem_u32(&tfull_bar[acc]));
                    if (++acc == acc_stages) { acc = 0; acc_phase ^= 1u; }
                    continue;
                }
                if (halo) {
                    for (int cb = 0; cb < p.ncb; ++cb) {
                        ptx::mbar_wait(ptx::smem_u32(&hfull_bar[hstage]), hphase);
                        ptx::tc_fence_after();
                        const uint32_t hbase = smem_base + hstage * halo_bytes;
                        for (int tap = 0; tap < 9; ++tap) {
                            ptx::mbar_wait(ptx::smem_u32(&full_bar[stage]), phase);
                            ptx::tc_fence_after();
                            // tap (r, s): the 16 x 8 output pixels read patch pixels (ly + r, lx + s); one patch row =
                            // 16 pixels = one 8-row group stride
                            const uint32_t a0 = hbase + static_cast<uint32_t>((tap / 3) * 16 + tap % 3) * row_bytes;
                            const uint64_t bdesc = ptx::make_kmajor_desc(ring_base + stage * b_bytes, row_bytes);
                            for (int k = 0; k < k_steps; ++k) {
                                const uint64_t adesc = ptx::make_kmajor_desc_sbo(a0 + 32u * k, row_bytes, 16u * row_bytes, p.desc_base_off);
                                ptx::umma_f16_ss(d_tmem, adesc, bdesc + 2ull * k, idesc, (cb | tap | k) != 0 ? 1u : 0u);
                            }
                            ptx::umma_commit(ptx::smem_u32(&empty_bar[stage]));
                            if (++stage == num_stages) { stage = 0; phase ^= 1u; }
                        }
                        ptx::umma_commit(ptx::smem_u32(&hempty_bar[hstage]));  // patch free once its 9 x k_steps MMAs retire
                        if (++hstage == p.halo_stages) { hstage = 0; hphase ^= 1u; }
                    }
                    ptx::umma_commit(ptx::smem_u32(&tfull_bar[acc]));
                    if (++acc == acc_stages) { acc = 0; acc_phase ^= 1u; }
                    continue;
                }
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(ptx::smem_u32(&full_bar[stage]), phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = smem_base + stage * stage_bytes;
                    const uint64_t adesc = ptx::make_kmajor_desc(sa, row_bytes);
                    const uint64_t bdesc = ptx::make_kmajor_desc(sa + a_bytes, row_bytes);
                    for (int k = 0; k < k_steps; ++k) {
                        // advance 16 fp16 = 32 B along K inside the swizzle span: +2 in the (addr>>4) field
                        ptx::umma_f16_ss(d_tmem, adesc + 2ull * k, bdesc + 2ull * k, idesc,
                                         (kb | k) != 0 ? 1u : 0u);
                    }
                    ptx::umma_commit(ptx::smem_u32(&empty_bar[stage]));  // frees the smem slot when MMAs retire
                    if (++stage == num_stages) { stage = 0; phase ^= 1u; }
                }
                ptx::umma_commit(ptx::smem_u32(&tfull_bar[acc]));  // accumulator complete -> epilogue
                if (++acc == acc_stages) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ===================== epilogue (8 warps, TMEM lane quadrant = warp % 4) =====================
        // Eight epilogue warps: warp w may only touch TMEM lanes 32*(w%4).., so the two warps of a quadrant split the
        // accumulator's 32-column chunks between them (even / odd chunk).  The epilogue, not the MMA, bounds every
        // GEMM of the recognisers (K <= 512: a GELU costs more issue slots than the 2*K/4096 MMA cycles it follows).
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        const int tiles_per_img = p.tiles_x * p.tiles_y;
        const int mode = p.mode, BLOCK_N = p.BLOCK_N, Cout = p.Cout;
        const int Ho = p.Ho, Wo = p.Wo, TW = p.TW, TH = p.TH, tiles_x = p.tiles_x;
        const int out_mode = p.out_mode, out_ld = p.out_ld, out_coff = p.out_coff, rep = p.rep;
        const int res_mode = p.res_mode, res_ld = p.res_ld;
        const float* __restrict__ bias = p.bias;
        const __half* __restrict__ res = reinterpret_cast<const __half*>(p.res);
        TileIter it(m_tiles_rt);
        int m_tile, n_tile;
        float best_v = -INFINITY;
        int best_i = 0;
        // TMA-store epilogue (fp16 NHWC output): this warp's two 2 KB staging tiles, used alternately
        const bool tma_st = !OUT_F32 && !ARGMAX && p.tma_store != 0;
        const uint32_t stg_base = smem_base + static_cast<uint32_t>(p.stg_off) + static_cast<uint32_t>(warp - 2) * 4096u;
        uint32_t n_st = 0;
        if (tma_st && lane == 0) ptx::prefetch_tmap(&p.tmD);
        while (it.next(p, m_tile, n_tile)) {
            // ---- which output pixel does this thread own?
            int img, y, x;
            bool valid;
            if (mode == A_FLAT) {
                const int m = m_tile * 128 + row;
                valid = m < M_rows;
                const int hw = Ho * Wo;
                img = m / hw;
                const int r = m - img * hw;
                y = r / Wo;
                x = r - y * Wo;
            } else {
                img = m_tile / tiles_per_img;
                const int t = m_tile - img * tiles_per_img;
                const int ty = t / tiles_x;
                const int tx = t - ty * tiles_x;
                const int ly = row / TW;
                y = ty * TH + ly;
                x = tx * TW + (row - ly * TW);
                valid = (y < Ho) && (x < Wo);
            }
            const long long pix = (static_cast<long long>(img) * Ho + y) * Wo + x;
            long long res_pix = pix;
            if (res_mode == RES_UP2)
                res_pix = (static_cast<long long>(img) * (Ho >> 1) + (y >> 1)) * (Wo >> 1) + (x >> 1);
            if (p.res_mod > 0) res_pix = pix % p.res_mod;  // broadcast rows (position embeddings)
            if (tma_st && !valid) res_pix = 0;  // rows outside the output still run the (warp-collective) store path; TMA clips them
            if constexpr (ARGMAX) {
                if (n_tile == 0) {
                    best_v = -INFINITY;
                    best_i = 0;
                }
            }

            ptx::mbar_wait(ptx::smem_u32(&tfull_bar[acc]), acc_phase);
            ptx::tc_fence_after();
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                   static_cast<uint32_t>(acc) * acc_stride;
            // arg-max epilogue: each of the two warps of a quadrant keeps the first maximum of ITS 32-column chunks (visited
            // in ascending order); the pair is merged after the last n-tile (ties -> lower index = torch.argmax)
            // Software-pipelined over 32-column chunks: the tcgen05.ld of the next chunk is in flight while the current
            // chunk's bias / residual / activation / stores issue (two warps per scheduler cannot hide it otherwise).
            constexpr int kStep = 64;
            int c = half * 32;
            uint32_t v[32];
            bool have = c < BLOCK_N && n_tile * BLOCK_N + c < Cout;  // warp-uniform
            if (have) ptx::tmem_ld_32x32b_x32(t_row + static_cast<uint32_t>(c), v);
            while (have) {
                const int col0 = n_tile * BLOCK_N + c;
                ptx::tmem_ld_wait();
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                c += kStep;
                have = c < BLOCK_N && n_tile * BLOCK_N + c < Cout;
                if (have) ptx::tmem_ld_32x32b_x32(t_row + static_cast<uint32_t>(c), v);
                if (!valid && !tma_st) continue;
                const int ncol = min(32, Cout - col0);
                if (bias_smem) {
                    const float4* b4 = reinterpret_cast<const float4*>(s_bias + col0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 b = b4[j];
                        f[4 * j] += b.x;
                        f[4 * j + 1] += b.y;
                        f[4 * j + 2] += b.z;
                        f[4 * j + 3] += b.w;
                    }
                } else if (bias != nullptr) {
                    const float4* b4 = reinterpret_cast<const float4*>(bias + col0);  // padded to 256
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 b = __ldg(b4 + j);
                        f[4 * j] += b.x;
                        f[4 * j + 1] += b.y;
                        f[4 * j + 2] += b.z;
                        f[4 * j + 3] += b.w;
                    }
                }
                if constexpr (RES_F32) {
                    if (res_mode != RES_NONE && !p.res_red) {
                        const float4* rp =
                            reinterpret_cast<const float4*>(reinterpret_cast<const float*>(res) + res_pix * res_ld + col0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (j * 4 < ncol) {
                                const float4 u = rp[j];
                                f[4 * j] += u.x;
                                f[4 * j + 1] += u.y;
                                f[4 * j + 2] += u.z;
                                f[4 * j + 3] += u.w;
                            }
                        }
                    }
                } else if (res_mode != RES_NONE) {
                    const uint4* rp = reinterpret_cast<const uint4*>(res + res_pix * res_ld + col0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (j * 8 < ncol) {
                            const uint4 u = __ldg(rp + j);
                            const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 t2 = __half22float2(h2[e]);
                                f[j * 8 + e * 2] += t2.x;
                                f[j * 8 + e * 2 + 1] += t2.y;
                            }
                        }
                    }
                    if (p.res_split_off > 0) {  // split residual: + its lo half
                        const uint4* rl = reinterpret_cast<const uint4*>(res + res_pix * res_ld + col0 + p.res_split_off);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (j * 8 < ncol) {
                                const uint4 u = __ldg(rl + j);
                                const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float2 t2 = __half22float2(h2[e]);
                                    f[j * 8 + e * 2] += t2.x;
                                    f[j * 8 + e * 2 + 1] += t2.y;
                                }
                            }
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    if constexpr (ACT == ACT_GELU) {
                        gelu_x2(f[j], f[j + 1], f[j], f[j + 1]);  // packed fp32: half the issue slots, same bits
                    } else {
                        f[j] = apply_act<ACT>(f[j]);
                        f[j + 1] = apply_act<ACT>(f[j + 1]);
                    }
                }
                if (p.post_affine) {
                    const float ps = p.post_scale, pb = p.post_bias;
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = fmaf(f[j], ps, pb);
                }
                if constexpr (ARGMAX) {
                    // torch.argmax semantics: first maximum wins (columns are visited in ascending order)
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (j < ncol && f[j] > best_v) {
                            best_v = f[j];
                            best_i = col0 + j;
                        }
                    }
                    if (p.out == nullptr) continue;
                }
                // ---- store
                int reps = 1;
                long long opix0 = pix;
                int ocol = col0;
                int orow_stride = 0;  // pixels per output row (for replication)
                if (out_mode == OUT_REPL) {
                    reps = rep;
                    orow_stride = Wo * rep;
                    opix0 = (static_cast<long long>(img) * Ho * rep + static_cast<long long>(y) * rep) * orow_stride +
                            static_cast<long long>(x) * rep;
                } else if (out_mode == OUT_SHUF2) {
                    const int cq = Cout >> 2;
                    const int quad = col0 / cq;
                    ocol = col0 - quad * cq;
                    opix0 = (static_cast<long long>(img) * Ho * 2 + 2 * y + (quad >> 1)) * (Wo * 2) + 2 * x + (quad & 1);
                }
                if constexpr (OUT_F32) {
                    float4 o[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                    for (int dy = 0; dy < reps; ++dy)
                        for (int dx = 0; dx < reps; ++dx) {
                            const long long opix = opix0 + static_cast<long long>(dy) * orow_stride + dx;
                            float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + opix * out_ld +
                                                                   out_coff + ocol);
                            if (RES_F32 && p.res_red) {
#pragma unroll
                                for (int j = 0; j < 8; ++j)
                                    if (j * 4 < ncol)
                                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(op + j), "f"(o[j].x), "f"(o[j].y),
                                                     "f"(o[j].z), "f"(o[j].w)
                                                     : "memory");
                                continue;
                            }
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                if (j * 4 < ncol) op[j] = o[j];
                        }
                } else {
                    uint4 o[4], ol[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        __half2* h2 = reinterpret_cast<__half2*>(&o[j]);
#pragma unroll
                        for (int e = 0; e < 4; ++e) h2[e] = __floats2half2_rn(f[j * 8 + e * 2], f[j * 8 + e * 2 + 1]);
                    }
                    if (tma_st) {
                        // 32 rows x 32 columns of this warp -> shared memory (64-byte rows, 16-byte piece j of row l at
                        // j ^ ((l >> 1) & 3) = SWIZZLE_64B: conflict-free) -> one cp.async.bulk.tensor store; rows / columns
                        // outside the tensor are clipped by the TMA unit
                        const uint32_t sbuf = stg_base + (n_st & 1u) * 2048u;
                        if (lane == 0) ptx::bulk_wait_read<1>();  // the store that last read this buffer (two stores ago) is done with it
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t addr = sbuf + static_cast<uint32_t>(lane) * 64u + ((static_cast<uint32_t>(j) ^ ((static_cast<uint32_t>(lane) >> 1) & 3u)) << 4);
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o[j].x), "r"(o[j].y), "r"(o[j].z), "r"(o[j].w)
                                         : "memory");
                        }
                        ptx::fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            if (mode == A_FLAT) {
                                ptx::tma_store_2d(&p.tmD, sbuf, col0, m_tile * 128 + q * 32);
                            } else {
                                const int r0 = q * 32, ly0 = r0 / TW;
                                const int t = m_tile - img * tiles_per_img;
                                const int ty = t / tiles_x, tx = t - ty * tiles_x;
                                ptx::tma_store_4d(&p.tmD, sbuf, col0, tx * TW + (r0 - ly0 * TW), ty * TH + ly0, img);
                            }
                            ptx::bulk_commit();
                        }
                        ++n_st;
                        continue;
                    }
                    const int split_off = p.split_off;
                    if (split_off > 0) {  // residual halves of the split-fp16 representation (any store pattern)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const __half2* h2 = reinterpret_cast<const __half2*>(&o[j]);
                            __half2* l2 = reinterpret_cast<__half2*>(&ol[j]);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 hi = __half22float2(h2[e]);
                                l2[e] = __floats2half2_rn(f[j * 8 + e * 2] - hi.x, f[j * 8 + e * 2 + 1] - hi.y);
                            }
                        }
                    }
                    for (int dy = 0; dy < reps; ++dy)
                        for (int dx = 0; dx < reps; ++dx) {
                            const long long opix = opix0 + static_cast<long long>(dy) * orow_stride + dx;
                            uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out) + opix * out_ld +
                                                                 out_coff + ocol);
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (j * 8 < ncol) op[j] = o[j];
                            if (split_off > 0) {
                                uint4* lp = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out) + opix * out_ld + out_coff + ocol + split_off);
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    if (j * 8 < ncol) lp[j] = ol[j];
                            }
                        }
                }
            }
            if constexpr (ARGMAX) {
                if (n_tile == p.n_tiles - 1) {
                    if (half == 1) {
                        s_argv[row] = best_v;
                        s_argi[row] = best_i;
                    }
                    asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");  // the two warps of quadrant q
                    if (half == 0 && valid) {
                        const float ov = s_argv[row];
                        const int oi = s_argi[row];
                        if (ov > best_v || (ov == best_v && oi < best_i)) {
                            best_v = ov;
                            best_i = oi;
                        }
                        p.arg_out[pix] = best_i;
                        if (p.max_out != nullptr) p.max_out[pix] = best_v;
                    }
                    asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");  // s_arg* may be rewritten for the next m-tile
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&tempty_bar[acc]));
            if (++acc == acc_stages) { acc = 0; acc_phase ^= 1u; }
        }
    }

    if (warp >= 2 && lane == 0 && p.tma_store != 0) ptx::bulk_wait_read<0>();  // shared memory must outlive the last tile stores
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace dv
