// conv_win_tcgen05 -- convolution of SMALL-CHANNEL, full-resolution layers (DLA base 7x7 3->16, level0 3x3 16->16, the 7x7
// stride-2 stems) on tcgen05 with a load/store PRODUCER instead of TMA.
//
// Why: for these layers one A-operand row is a window of 8 pixels x 8 channels (or 4 x 16) = 128 contiguous bytes of a
// zero-bordered NHWC image, and consecutive output pixels read windows that overlap by 7/8.  Fetched by TMA (A_STEM /
// A_PATCH modes of conv_igemm_tcgen05) every 128-byte row is its own L2 request and the kernel is bound by the TMA row
// rate (~4.4 clk per row per SM, profiles/r1m, r1p: tensor pipe 8 % active, 6000 clk per 128-pixel tile).  Here 128
// producer threads (one per output pixel of the tile) read their window with eight 16-byte ld.global.nc -- neighbouring
// windows hit in L1 -- and store it into the canonical K-major SWIZZLE_128B layout (chunk c of row m at
// m*128 + ((c ^ (m & 7)) << 4)), fence.proxy.async, and arrive on the stage's mbarrier.  The filter (KR k-blocks of
// [BLOCK_N][64]) is loaded ONCE per CTA by TMA and stays resident.  One k-block = one filter row.
//
// Roles: warps 0-3 producers, warp 4 TMEM allocator + MMA issuer + weight loader, warps 5-8 epilogue (TMEM quadrant =
// warp % 4).  Persistent CTAs, NS-deep A ring, 2-deep TMEM accumulator ring.
#pragma once
#include "igemm_params.h"
#include "win_conv_params.h"
#include "ptx.cuh"

namespace dv {

static constexpr int kWinThreads = 9 * 32;

template <int ACT>
__device__ __forceinline__ float win_act(float x) {
    if constexpr (ACT == ACT_RELU) return fmaxf(x, 0.f);
    return x;
}

template <int ACT>
__global__ void __launch_bounds__(kWinThreads, 1)
conv_win_tcgen05(const __grid_constant__ WinConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[12];
    __shared__ __align__(8) uint64_t empty_bar[12];
    __shared__ __align__(8) uint64_t tfull_bar[2];
    __shared__ __align__(8) uint64_t tempty_bar[2];
    __shared__ __align__(8) uint64_t w_bar;
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float s_bias[64];
    __shared__ uint64_t s_adesc[28], s_bdesc[28];  // patch mode: per-(filter row, K16 slice) descriptors, tabulated once
    __shared__ uint32_t s_dsel[28], s_accum[28];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_bytes = static_cast<uint32_t>(p.BLOCK_N) * 128u;
    const uint32_t a_base = smem_base + static_cast<uint32_t>(p.KR) * b_bytes;  // B (resident) first, then the A ring
    constexpr uint32_t kABytes = 128u * 128u;
    const int num_stages = p.num_stages;
    // 2 tile buffers x kSplit partial accumulators (filter rows are dealt round-robin to kSplit INDEPENDENT accumulation
    // chains, summed by the epilogue: a chain of dependent tcgen05.mma costs ~150 clk per instruction regardless of N)
    constexpr uint32_t kSplit = 2;
    const uint32_t tmem_cols = p.BLOCK_N <= 16 ? 64u : p.BLOCK_N <= 32 ? 128u : 256u;

    if (threadIdx.x < 64) s_bias[threadIdx.x] = threadIdx.x < p.BLOCK_N && p.bias ? __ldg(p.bias + threadIdx.x) : 0.f;
    if (threadIdx.x == 0) {
        for (int i = 0; i < num_stages; ++i) {
            ptx::mbar_init(ptx::smem_u32(&full_bar[i]), 128);
            ptx::mbar_init(ptx::smem_u32(&empty_bar[i]), 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(ptx::smem_u32(&tfull_bar[i]), 1);
            ptx::mbar_init(ptx::smem_u32(&tempty_bar[i]), 4);
        }
        ptx::mbar_init(ptx::smem_u32(&w_bar), 1);
        ptx::fence_barrier_init();
        ptx::prefetch_tmap(&p.tmB);
    }
    if (warp == 4) {
        ptx::tmem_alloc(ptx::smem_u32(&tmem_base_smem), tmem_cols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const uint32_t acc_cols = tmem_cols >> 1;            // per tile buffer
    const uint32_t part_cols = acc_cols / kSplit;        // per partial accumulator (>= BLOCK_N)

    if (warp < 4) {
        // ===================== producers: thread t owns A row t of every stage =====================
        // cp.async (LDGSTS) straight into the swizzled slot, kDepth k-blocks in flight per thread; a k-block is published
        // (fence.proxy.async + mbarrier arrive) once its group has landed.  A first version with ld.global -> registers ->
        // st.shared exposed one L2 round trip per k-block (base: 2.95 ms, slower than TMA).
        constexpr int kDepth = 4;
        if (p.patch) {
            const int chunks = p.planes * p.PR * 16;
            const uint32_t stage_bytes = static_cast<uint32_t>(chunks) * 16u;
            int stage = 0, pub_stage = 0, issued = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
                const int img = tile / tiles_per_img;
                const int t = tile - img * tiles_per_img;
                const int ty = t / p.tiles_x;
                const int y0 = ty * p.TH, x0 = (t - ty * p.tiles_x) * p.TW;
                ptx::mbar_wait(ptx::smem_u32(&empty_bar[stage]), phase ^ 1u);
                const uint32_t sbase = a_base + static_cast<uint32_t>(stage) * stage_bytes;
                for (int i = threadIdx.x; i < chunks; i += 128) {
                    const int plane = i / (p.PR * 16);
                    const int rem = i - plane * p.PR * 16;
                    const int prow = rem >> 4, px = rem & 15;
                    const int y = y0 + prow, x = x0 + px;
                    const uint32_t nb = (y < p.Hp && x < p.Wp) ? 16u : 0u;
                    const __half* g = p.in + ((static_cast<long long>(img) * p.Hp + min(y, p.Hp - 1)) * p.Wp + min(x, p.Wp - 1)) * p.cpp + plane * 8;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(sbase + static_cast<uint32_t>(i) * 16u), "l"(g), "r"(nb)
                                 : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                if (++stage == num_stages) { stage = 0; phase ^= 1u; }
                if (++issued > kDepth) {
                    asm volatile("cp.async.wait_group %0;" ::"n"(kDepth) : "memory");
                    ptx::fence_proxy_async_smem();
                    ptx::mbar_arrive(ptx::smem_u32(&full_bar[pub_stage]));
                    if (++pub_stage == num_stages) pub_stage = 0;
                }
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            ptx::fence_proxy_async_smem();
            for (int k = issued > kDepth ? kDepth : issued; k > 0; --k) {
                ptx::mbar_arrive(ptx::smem_u32(&full_bar[pub_stage]));
                if (++pub_stage == num_stages) pub_stage = 0;
            }
        } else {
        const int m = threadIdx.x;
        const int ly = m / p.TW, lx = m - ly * p.TW;
        const uint32_t row_off = static_cast<uint32_t>(m) * 128u;
        const uint32_t sw = static_cast<uint32_t>(m & 7);
        const long long row_pitch = static_cast<long long>(p.Wp) * p.cpp;  // halves per padded image row
        int stage = 0, pub_stage = 0, issued = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
            const int img = tile / tiles_per_img;
            const int t = tile - img * tiles_per_img;
            const int ty = t / p.tiles_x;
            const int oy = ty * p.TH + ly, ox = (t - ty * p.tiles_x) * p.TW + lx;
            const uint32_t src_bytes = (oy < p.Ho && ox < p.Wo) ? 16u : 0u;  // 0: the slot is zero-filled
            const int cy = min(oy, p.Ho - 1), cx = min(ox, p.Wo - 1);      // keep the (ignored) address in bounds
            const __half* src = p.in + ((static_cast<long long>(img) * p.Hp + static_cast<long long>(cy) * p.stride) * p.Wp +
                                        static_cast<long long>(cx) * p.stride) * p.cpp;
            for (int r = 0; r < p.KR; ++r) {
                ptx::mbar_wait(ptx::smem_u32(&empty_bar[stage]), phase ^ 1u);
                const uint32_t dst = a_base + static_cast<uint32_t>(stage) * kABytes + row_off;
                const char* g = reinterpret_cast<const char*>(src + r * row_pitch);
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst + ((static_cast<uint32_t>(c) ^ sw) << 4)),
                                 "l"(g + c * 16), "r"(src_bytes)
                                 : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
                if (++stage == num_stages) { stage = 0; phase ^= 1u; }
                if (++issued > kDepth) {
                    asm volatile("cp.async.wait_group %0;" ::"n"(kDepth) : "memory");
                    ptx::fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async proxy
                    ptx::mbar_arrive(ptx::smem_u32(&full_bar[pub_stage]));
                    if (++pub_stage == num_stages) pub_stage = 0;
                }
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        ptx::fence_proxy_async_smem();
        for (int k = issued > kDepth ? kDepth : issued; k > 0; --k) {
            ptx::mbar_arrive(ptx::smem_u32(&full_bar[pub_stage]));
            if (++pub_stage == num_stages) pub_stage = 0;
        }
        }
    } else if (warp == 4) {
        // ===================== weight loader + MMA issuer (one thread) =====================
        if (ptx::elect_one_sync()) {
            const uint32_t wb = ptx::smem_u32(&w_bar);
            ptx::mbar_expect_tx(wb, static_cast<uint32_t>(p.KR) * b_bytes);
            for (int r = 0; r < p.KR; ++r) ptx::tma_load_2d(smem_base + r * b_bytes, &p.tmB, wb, r * 64, 0);
            ptx::mbar_wait(wb, 0);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            const uint32_t idesc = ptx::make_idesc_f16_m128(static_cast<uint32_t>(p.BLOCK_N));
            const uint32_t patch_stage_bytes = static_cast<uint32_t>(p.planes * p.PR) * 256u;
            const int n_mma = p.KR * 4;
            if (p.patch) {
                const int per_plane = 4 / p.planes;  // MMAs (2 pixels x 8 channels each) per plane and filter row
                for (int r = 0; r < p.KR; ++r)
                    for (int k = 0; k < 4; ++k) {
                        const int plane = k / per_plane, j = k - plane * per_plane;
                        const uint32_t off = static_cast<uint32_t>(plane * p.PR) * 256u + static_cast<uint32_t>(r * 16 + 2 * j) * 16u;
                        s_adesc[r * 4 + k] = ptx::make_kmajor_desc_noswz(off, 16u, 256u);  // + (stage base >> 4) per tile
                        s_bdesc[r * 4 + k] = ptx::make_kmajor_desc(smem_base + r * b_bytes, 128) + 2ull * k;
                        s_dsel[r * 4 + k] = static_cast<uint32_t>(r & 1) * part_cols;
                        s_accum[r * 4 + k] = ((r >> 1) | k) != 0 ? 1u : 0u;
                    }
            }
            for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
                ptx::mbar_wait(ptx::smem_u32(&tempty_bar[acc]), acc_phase ^ 1u);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc) * acc_cols;
                if (p.patch) {
                    ptx::mbar_wait(ptx::smem_u32(&full_bar[stage]), phase);
                    ptx::tc_fence_after();
                    // descriptors differ from tile to tile only by the stage base: everything else was tabulated once
                    const uint64_t abase = static_cast<uint64_t>((a_base + static_cast<uint32_t>(stage) * patch_stage_bytes) >> 4);
                    for (int i = 0; i < n_mma; ++i)
                        ptx::umma_f16_ss(d_tmem + s_dsel[i], s_adesc[i] + abase, s_bdesc[i], idesc, s_accum[i]);
                    ptx::umma_commit(ptx::smem_u32(&empty_bar[stage]));
                    if (++stage == num_stages) { stage = 0; phase ^= 1u; }
                    ptx::umma_commit(ptx::smem_u32(&tfull_bar[acc]));
                    acc ^= 1;
                    if (acc == 0) acc_phase ^= 1u;
                    continue;
                }
                for (int r = 0; r < p.KR; ++r) {
                    ptx::mbar_wait(ptx::smem_u32(&full_bar[stage]), phase);
                    ptx::tc_fence_after();
                    const uint64_t adesc = ptx::make_kmajor_desc(a_base + static_cast<uint32_t>(stage) * kABytes, 128);
                    const uint64_t bdesc = ptx::make_kmajor_desc(smem_base + r * b_bytes, 128);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::umma_f16_ss(d_tmem + static_cast<uint32_t>(r & 1) * part_cols, adesc + 2ull * k, bdesc + 2ull * k, idesc, ((r >> 1) | k) != 0 ? 1u : 0u);
                    ptx::umma_commit(ptx::smem_u32(&empty_bar[stage]));
                    if (++stage == num_stages) { stage = 0; phase ^= 1u; }
                }
                ptx::umma_commit(ptx::smem_u32(&tfull_bar[acc]));
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else {
        // ===================== epilogue (4 warps) =====================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int ly = row / p.TW, lx = row - ly * p.TW;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
            const int img = tile / tiles_per_img;
            const int t = tile - img * tiles_per_img;
            const int ty = t / p.tiles_x;
            const int oy = ty * p.TH + ly, ox = (t - ty * p.tiles_x) * p.TW + lx;
            const bool valid = oy < p.Ho && ox < p.Wo;
            __half* op = p.out + ((static_cast<long long>(img) * p.oHp + oy + p.opad) * p.oWp + ox + p.opad) * p.out_ld;
            ptx::mbar_wait(ptx::smem_u32(&tfull_bar[acc]), acc_phase);
            ptx::tc_fence_after();
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc) * acc_cols;
            for (int c = 0; c < p.BLOCK_N; c += 16) {
                uint32_t v[16], v2[16];
                ptx::tmem_ld_32x32b_x16(t_row + static_cast<uint32_t>(c), v);
                ptx::tmem_ld_32x32b_x16(t_row + part_cols + static_cast<uint32_t>(c), v2);
                ptx::tmem_ld_wait();
                if (!valid || c >= p.Cout) continue;
                if (p.KR > 1) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
                }
                uint4 o[2];
                __half2* h2 = reinterpret_cast<__half2*>(o);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    h2[j] = __floats2half2_rn(win_act<ACT>(__uint_as_float(v[2 * j]) + s_bias[c + 2 * j]),
                                              win_act<ACT>(__uint_as_float(v[2 * j + 1]) + s_bias[c + 2 * j + 1]));
                uint4* dst = reinterpret_cast<uint4*>(op + c);
                dst[0] = o[0];
                if (c + 8 < p.Cout) dst[1] = o[1];
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&tempty_bar[acc]));
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, tmem_cols);
    }
}

}  // namespace dv
