// mlp_fused_tcgen05: x += W2 * GELU(W1 * h + b1) + b2 in ONE kernel -- the pwconv1 -> GELU -> pwconv2 (+ layer_scale,
// folded into W2 / b2 at pack time) + residual tail of HF ConvNextLayer (modeling_convnext.py, used by
// convnext_vit/modeling_convnext.py:28-80) and the intermediate -> GELU -> output MLP of the ViT layers
// (modeling_vit.py:32-180).  As two conv_igemm_tcgen05 launches this pair was 43 % of the recogniser: the 4C-wide hidden
// tensor (8C bytes per token) was written by one epilogue-bound GEMM and read back by a second, HBM-bound one
// (profiles/r1t_layers_rec.txt).  Here the hidden tensor never leaves the SM:
//
//   per 128-token tile, per 64-column chunk j of the hidden dimension (NC = 4C / 64 chunks)
//     GEMM1  acc1[b]  = h_tile[128 x C] * W1[64j .. 64j+64, :]^T      tcgen05.mma M128 N64, K = C (zero-padded to 64s)
//     epi1   G[g]     = fp16(GELU(acc1[b] + b1))                      tcgen05.ld -> registers -> st.shared, written in
//                                                                     the 128-byte-swizzled K-major layout a UMMA
//                                                                     descriptor reads (16-byte chunk ^= row & 7)
//     GEMM2  acc2    += G[g][128 x 64] * W2[:, 64j .. 64j+64]^T       tcgen05.mma M128 N=C, K = 64
//   final  x_tile    += acc2 + b2                                     fp32 residual stream, in place
//
// Warp roles (608 threads): warp 0 = TMA producer of the h tile (once per tile) and the W1 chunk ring, warp 2 = TMA
// producer of the W2 chunk ring (both rings 2 deep), warp 1 = MMA issuer (event-driven: whichever of GEMM1(next) /
// GEMM2(next) has its operands first), warps 3-18 = epilogue in two groups of eight that alternate chunks.
// TMEM: acc2 at columns [0, C), acc1 double-buffered at 256 + 64 b.
#pragma once
#include <cuda.h>

#include "igemm.cuh"
#include "mlp_params.h"

namespace dv {

constexpr int kMlpEpiWarps = 16;  // two groups of eight: two warps per TMEM lane quadrant, 32 of a chunk's 64 columns each
constexpr int kMlpThreads = 96 + 32 * kMlpEpiWarps;

template <int C>
struct MlpCfg {
    static constexpr int KB1 = (C + 63) / 64;          // k-blocks of GEMM1 (TMA zero-fills columns >= C)
    static constexpr int NC = 4 * C / 64;              // hidden chunks
    static constexpr int A_BYTES = KB1 * 128 * 128;    // h tile: KB1 atoms of 128 rows x 128 B
    static constexpr int W1_SLOT = KB1 * 64 * 128;     // KB1 atoms of 64 rows x 128 B
    static constexpr int W2_SLOT = C * 128;            // one atom of C rows x 128 B
    static constexpr int G_BYTES = 128 * 128;          // one atom of 128 rows x 128 B
    static constexpr int GBUF = (C <= 192) ? 2 : 1;    // C = 256: shared memory holds one hidden buffer only
    static constexpr int SMEM = A_BYTES + 2 * W1_SLOT + 2 * W2_SLOT + GBUF * G_BYTES + 5 * C * 4 + 1024;
};

template <int C>
__global__ void __launch_bounds__(kMlpThreads, 1)
mlp_fused_tcgen05(const __grid_constant__ MlpParams p) {
    using Cfg = MlpCfg<C>;
    constexpr int KB1 = Cfg::KB1, NC = Cfg::NC, GBUF = Cfg::GBUF;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t a_full, a_empty, acc2_full, acc2_empty;
    __shared__ __align__(8) uint64_t w1_full[2], w1_empty[2], w2_full[2], w2_empty[2], acc1_full[2], acc1_empty[2];
    __shared__ __align__(8) uint64_t g_full[2], g_empty[2];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = smem_base;
    const uint32_t sW1 = sA + Cfg::A_BYTES;
    const uint32_t sW2 = sW1 + 2 * Cfg::W1_SLOT;
    const uint32_t sG = sW2 + 2 * Cfg::W2_SLOT;
    float* s_b1 = reinterpret_cast<float*>(smem_raw + (sG + GBUF * Cfg::G_BYTES - ptx::smem_u32(smem_raw)));
    float* s_b2 = s_b1 + 4 * C;

    for (int i = threadIdx.x; i < 4 * C; i += kMlpThreads) s_b1[i] = __ldg(p.b1 + i);
    for (int i = threadIdx.x; i < C; i += kMlpThreads) s_b2[i] = __ldg(p.b2 + i);
    if (threadIdx.x == 0) {
        ptx::mbar_init(ptx::smem_u32(&a_full), 1);
        ptx::mbar_init(ptx::smem_u32(&a_empty), 1);
        ptx::mbar_init(ptx::smem_u32(&acc2_full), 1);
        ptx::mbar_init(ptx::smem_u32(&acc2_empty), kMlpEpiWarps);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(ptx::smem_u32(&w1_full[i]), 1);
            ptx::mbar_init(ptx::smem_u32(&w1_empty[i]), 1);
            ptx::mbar_init(ptx::smem_u32(&w2_full[i]), 1);
            ptx::mbar_init(ptx::smem_u32(&w2_empty[i]), 1);
            ptx::mbar_init(ptx::smem_u32(&acc1_full[i]), 1);
            ptx::mbar_init(ptx::smem_u32(&acc1_empty[i]), kMlpEpiWarps / 2);
            ptx::mbar_init(ptx::smem_u32(&g_full[i]), kMlpEpiWarps / 2);
            ptx::mbar_init(ptx::smem_u32(&g_empty[i]), 1);
        }
        ptx::fence_barrier_init();
        ptx::prefetch_tmap(&p.tmA);
        ptx::prefetch_tmap(&p.tmW1);
        ptx::prefetch_tmap(&p.tmW2);
    }
    if (warp == 1) {
        ptx::tmem_alloc(ptx::smem_u32(&tmem_base_smem), 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    const int m_tiles = p.m_tiles;

    if (warp == 0) {
        // ===================== h tile + W1 producer =====================
        if (ptx::elect_one_sync()) {
            uint32_t n = 0, t = 0;  // chunk / tile counters (ring slot = n & 1, phase = (n >> 1) & 1)
            for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++t) {
                ptx::mbar_wait(ptx::smem_u32(&a_empty), (t & 1u) ^ 1u);
                const uint32_t ab = ptx::smem_u32(&a_full);
                ptx::mbar_expect_tx(ab, Cfg::A_BYTES);
#pragma unroll
                for (int kb = 0; kb < KB1; ++kb) ptx::tma_load_2d(sA + kb * 16384, &p.tmA, ab, kb * 64, tile * 128);
                for (int j = 0; j < NC; ++j, ++n) {
                    const uint32_t s = n & 1u, ph = (n >> 1) & 1u;
                    ptx::mbar_wait(ptx::smem_u32(&w1_empty[s]), ph ^ 1u);
                    const uint32_t b1b = ptx::smem_u32(&w1_full[s]);
                    if (p.dbg == 2 && n >= 2) {  // tuning aid: no weight traffic after the first two chunks (stale weights)
                        ptx::mbar_arrive(b1b);
                        continue;
                    }
                    ptx::mbar_expect_tx(b1b, Cfg::W1_SLOT);
#pragma unroll
                    for (int kb = 0; kb < KB1; ++kb)
                        ptx::tma_load_2d(sW1 + s * Cfg::W1_SLOT + kb * 8192, &p.tmW1, b1b, kb * 64, j * 64);
                }
            }
        }
    } else if (warp == 2) {
        // ===================== W2 producer =====================
        // A separate thread: W2's slot is released by GEMM2(j) (end of chunk j's GELU) while W1's is released two chunks
        // earlier; one in-order producer made every W1 load queue behind a W2 wait and the chunk period became one TMA
        // latency + the GEMMs (2.9 us per chunk at C = 192, whatever the epilogue did).
        if (ptx::elect_one_sync()) {
            uint32_t n = 0;
            for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
                for (int j = 0; j < NC; ++j, ++n) {
                    const uint32_t s = n & 1u, ph = (n >> 1) & 1u;
                    ptx::mbar_wait(ptx::smem_u32(&w2_empty[s]), ph ^ 1u);
                    const uint32_t b2b = ptx::smem_u32(&w2_full[s]);
                    if (p.dbg == 2 && n >= 2) {
                        ptx::mbar_arrive(b2b);
                        continue;
                    }
                    ptx::mbar_expect_tx(b2b, Cfg::W2_SLOT);
                    ptx::tma_load_2d(sW2 + s * Cfg::W2_SLOT, &p.tmW2, b2b, j * 64, 0);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (ptx::elect_one_sync()) {
            const uint32_t idesc1 = ptx::make_idesc_f16_m128(64), idesc2 = ptx::make_idesc_f16_m128(C);
            // Event-driven issue: GEMM1(n1) goes out as soon as its weights and its accumulator buffer are there (the buffer
            // is released when the epilogue has pulled chunk n1 - 2 into registers), GEMM2(n2) as soon as its hidden tile is
            // written -- whichever is ready first.  A fixed program order (GEMM1(j+2) behind GEMM2(j-1)) left the epilogue
            // waiting on acc1_full for 20 % of its time (profiles/r1z_mlp_ncu.txt, v6).
            const uint32_t my_tiles = m_tiles > static_cast<int>(blockIdx.x) ? (m_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
            const uint32_t total = my_tiles * NC;
            uint32_t n1 = 0, n2 = 0, spins = 0;
            while (n2 < total) {
                bool did = false;
                if (n1 < total) {
                    const uint32_t s = n1 & 1u, ph = (n1 >> 1) & 1u, j1 = n1 % NC, t1 = n1 / NC;
                    if (ptx::mbar_test_wait(ptx::smem_u32(&w1_full[s]), ph) && ptx::mbar_test_wait(ptx::smem_u32(&acc1_empty[s]), ph ^ 1u) &&
                        (j1 != 0 || ptx::mbar_test_wait(ptx::smem_u32(&a_full), t1 & 1u))) {
                        ptx::tc_fence_after();
                        const uint32_t d = tmem_base + 256u + s * 64u;
#pragma unroll
                        for (int kb = 0; kb < KB1; ++kb)
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                ptx::umma_f16_ss(d, ptx::make_kmajor_desc(sA + kb * 16384 + k * 32, 128),
                                                 ptx::make_kmajor_desc(sW1 + s * Cfg::W1_SLOT + kb * 8192 + k * 32, 128), idesc1,
                                                 (kb | k) != 0 ? 1u : 0u);
                        ptx::umma_commit(ptx::smem_u32(&w1_empty[s]));
                        ptx::umma_commit(ptx::smem_u32(&acc1_full[s]));
                        if (j1 == NC - 1) ptx::umma_commit(ptx::smem_u32(&a_empty));  // every GEMM1 of this tile has been issued
                        ++n1;
                        did = true;
                    }
                }
                if (n2 < n1) {
                    const uint32_t s = n2 & 1u, ph = (n2 >> 1) & 1u, j2 = n2 % NC, t2 = n2 / NC;
                    const uint32_t gs = n2 % GBUF;  // hidden buffer; its barriers are indexed by chunk parity (= epilogue group)
                    if (ptx::mbar_test_wait(ptx::smem_u32(&w2_full[s]), ph) && ptx::mbar_test_wait(ptx::smem_u32(&g_full[s]), ph) &&
                        (j2 != 0 || ptx::mbar_test_wait(ptx::smem_u32(&acc2_empty), (t2 & 1u) ^ 1u))) {
                        ptx::tc_fence_after();
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            ptx::umma_f16_ss(tmem_base, ptx::make_kmajor_desc(sG + gs * Cfg::G_BYTES + k * 32, 128),
                                             ptx::make_kmajor_desc(sW2 + s * Cfg::W2_SLOT + k * 32, 128), idesc2,
                                             (j2 | static_cast<uint32_t>(k)) != 0 ? 1u : 0u);
                        ptx::umma_commit(ptx::smem_u32(&w2_empty[s]));
                        ptx::umma_commit(ptx::smem_u32(&g_empty[s]));
                        if (j2 == NC - 1) ptx::umma_commit(ptx::smem_u32(&acc2_full));
                        ++n2;
                        did = true;
                    }
                }
                if (did) spins = 0;
                else if (++spins > (1u << 28)) __trap();  // protocol bug: fail the launch instead of hanging the GPU
            }
        }
    } else {
        // ===================== epilogue: 16 warps in two groups =====================
        // Group g (two warps per TMEM lane quadrant, 32 columns each) owns the chunks j = g (mod 2), i.e. accumulator
        // buffer g and (GBUF = 2) hidden buffer g.  The groups run half a chunk out of step, so one group's barrier
        // waits, TMEM round trip, shared-memory stores and proxy fence are covered by the other group's GELU math: with
        // all warps walking every chunk together those phases were exposed on every scheduler at once (issue slots
        // 44 % used, profiles/r1z_mlp_ncu.txt).
        const int q = warp & 3, sub = (warp - 3) >> 2, grp = sub & 1, half = sub >> 1;
        const int part = sub;  // column quarter of the final update
        const int row = q * 32 + lane;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        // this thread's four 16-byte pieces of its row of G: chunk index (half * 4 + i) ^ (row & 7) inside the 128-byte row
        const uint32_t g_row = static_cast<uint32_t>((row >> 3) * 1024 + (row & 7) * 128);
        uint32_t t = 0;
        constexpr int QC = C / 4;  // columns of the final update owned by this warp
        for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++t) {
            const long long grow = static_cast<long long>(tile) * 128 + row;
            float* xr = p.x + grow * C + part * QC;
            const bool valid = grow < p.M;
            for (int j = grp; j < NC; j += 2) {
                const uint32_t n = t * NC + j;  // global chunk counter (NC is even: n = grp mod 2)
                const uint32_t b = n & 1u, ph = (n >> 1) & 1u;
                const uint32_t gs = n % GBUF;
                ptx::mbar_wait(ptx::smem_u32(&acc1_full[b]), ph);
                ptx::tc_fence_after();
                uint32_t v[32];
                ptx::tmem_ld_32x32b_x32(t_lane + 256u + b * 64u + static_cast<uint32_t>(half * 32), v);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&acc1_empty[b]));  // TMEM buffer free: GEMM1(j+2) may start
                const float4* bb = reinterpret_cast<const float4*>(s_b1 + j * 64 + half * 32);
                uint4 o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float f[8];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float4 bv = bb[i * 2 + e];
                        if (p.dbg == 1 || p.dbg >= 3) {  // tuning aid: the kernel without its GELU math
                            f[e * 4 + 0] = __uint_as_float(v[i * 8 + e * 4 + 0]) + bv.x;
                            f[e * 4 + 1] = __uint_as_float(v[i * 8 + e * 4 + 1]) + bv.y;
                            f[e * 4 + 2] = __uint_as_float(v[i * 8 + e * 4 + 2]) + bv.z;
                            f[e * 4 + 3] = __uint_as_float(v[i * 8 + e * 4 + 3]) + bv.w;
                            continue;
                        }
                        gelu_bias_x2(__uint_as_float(v[i * 8 + e * 4 + 0]), __uint_as_float(v[i * 8 + e * 4 + 1]), bv.x, bv.y,
                                     f[e * 4 + 0], f[e * 4 + 1]);
                        gelu_bias_x2(__uint_as_float(v[i * 8 + e * 4 + 2]), __uint_as_float(v[i * 8 + e * 4 + 3]), bv.z, bv.w,
                                     f[e * 4 + 2], f[e * 4 + 3]);
                    }
                    __half2 h0 = __floats2half2_rn(f[0], f[1]), h1 = __floats2half2_rn(f[2], f[3]);
                    __half2 h2 = __floats2half2_rn(f[4], f[5]), h3 = __floats2half2_rn(f[6], f[7]);
                    o[i] = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                      *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
                }
                // GEMM2 of the chunk that last used this hidden buffer must be done.  The g barriers are indexed by chunk
                // parity; with two buffers that is this group's own barrier (chunk n - 2), with one buffer the other
                // group's (chunk n - 1) -- either way a group waits on consecutive phases of one barrier, never skipping one.
                if constexpr (GBUF == 2) ptx::mbar_wait(ptx::smem_u32(&g_empty[b]), ph ^ 1u);
                else if (n > 0) ptx::mbar_wait(ptx::smem_u32(&g_empty[b ^ 1u]), ((n - 1) >> 1) & 1u);
                const uint32_t gb = sG + gs * Cfg::G_BYTES + g_row;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (p.dbg == 3) break;  // tuning aid: no hidden-tile stores
                    const uint32_t addr = gb + ((static_cast<uint32_t>(half * 4 + i) ^ static_cast<uint32_t>(row & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o[i].x), "r"(o[i].y), "r"(o[i].z),
                                 "r"(o[i].w)
                                 : "memory");
                }
                if (p.dbg != 3) ptx::fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async proxy
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&g_full[b]));
            }
            // ---- final: x += acc2 + b2 (this warp: columns [part * C/4, (part + 1) * C/4) of its 32 rows); the residual
            // loads are all in flight before the wait for the last GEMM2
            // x += acc2 + b2 as fire-and-forget vector reductions (red.global.add.v4.f32): the fp32 sum is formed at L2, bit for
            // bit the value a load / add / store would give, but no thread ever waits for x.  With the read-modify-write
            // form this stage was 40 % of the kernel (0.111 -> 0.068 ms at C = 192 with it stubbed out, tools note in
            // profiles/r2f_mlp_stage_costs.md): sixteen warps stalled on two DRAM round trips per tile.
            if (p.dbg == 4) {  // tuning aid: no residual update
                ptx::mbar_wait(ptx::smem_u32(&acc2_full), t & 1u);
                ptx::tc_fence_after();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&acc2_empty));
                continue;
            }
            ptx::mbar_wait(ptx::smem_u32(&acc2_full), t & 1u);
            ptx::tc_fence_after();
            constexpr int BATCH = QC > 32 ? QC / 2 : QC;  // 24 / 24 / 32 columns per TMEM round trip
#pragma unroll
            for (int c0 = 0; c0 < QC; c0 += BATCH) {
                uint32_t w[BATCH / 8][8];
#pragma unroll
                for (int c = 0; c < BATCH; c += 8) ptx::tmem_ld_32x32b_x8(t_lane + static_cast<uint32_t>(part * QC + c0 + c), w[c / 8]);
                ptx::tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int i = 0; i < BATCH / 4; ++i) {
                        const float4 bv = *reinterpret_cast<const float4*>(s_b2 + part * QC + c0 + i * 4);
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(xr + c0 + i * 4),
                                     "f"(__uint_as_float(w[i / 2][(i & 1) * 4 + 0]) + bv.x), "f"(__uint_as_float(w[i / 2][(i & 1) * 4 + 1]) + bv.y),
                                     "f"(__uint_as_float(w[i / 2][(i & 1) * 4 + 2]) + bv.z), "f"(__uint_as_float(w[i / 2][(i & 1) * 4 + 3]) + bv.w)
                                     : "memory");
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&acc2_empty));
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace dv
