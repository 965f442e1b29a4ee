// dcn_fused_tcgen05: the modulated deformable 3x3 convolution of Lore's DLA-up / IDA-up nodes (DCN, lore/dcnv2.py:71-86 ->
// torchvision.ops.deform_conv2d; the reference's own CUDA spec: DCNv2_latest/src/cuda/dcn_v2_im2col_cuda.cu:121-191) + folded
// BatchNorm + ReLU as ONE kernel.  The sampled-column matrix (`columns` in the reference kernel, 9C fp16 per pixel: 1.2 GB
// per 64-channel 256 x 256 layer of a 16-image batch) never exists in global memory: producer warps gather the four
// bilinear corners of every (pixel, tap), blend them with the mask-scaled weights and write the result straight into the
// 128-byte-swizzled K-major shared-memory tile a UMMA descriptor reads; the weights stream through a TMA ring; one thread
// issues tcgen05.mma into a double-buffered TMEM accumulator; four epilogue warps add the bias, apply ReLU and store NHWC.
//
//   tile   = 8 x 16 output pixels of one image (128 MMA rows): the 36 samples of a pixel and those of its neighbours hit
//            the same few input pixels, so a spatially compact tile keeps the gather in L1
//   chunk  = 64 of the K = 9C columns = (tap, 64-channel block): A stage 128 x 128 B, B stage cout x 128 B
//   table  = per tile, per (tap, pixel): corner-0 byte offset + the two clamped steps, four mask-scaled fp16 weights --
//            computed ONCE (sigmoid, floor, clamps) by four set-up warps that run one tile ahead of the gather through a
//            double-buffered shared-memory table
//
// The arithmetic is that of k_dcn_im2col + the flat GEMM it replaces (fp16 weights, fma.rn.f32.f16 blend, fp16 rounding,
// K walked in the same order), so the two paths are bit-identical (tests/test_gpu_lore.py, DV_DCN_FUSED=0 selects the
// three-launch path).
#pragma once
#include <cuda.h>

#include "dcn_params.h"
#include "igemm_params.h"
#include "ptx.cuh"

namespace dv {

constexpr int kDcnSetupWarps = 4;  // one thread per tile row: the nine (tap, pixel) jobs of its pixel
constexpr int kDcnProdWarps = 16;
constexpr int kDcnFirstSetup = 6;  // warp 0 = weight TMA, 1 = MMA issuer, 2-5 = epilogue, 6-9 = set-up, 10.. = gather producers
constexpr int kDcnFirstProd = kDcnFirstSetup + kDcnSetupWarps;
constexpr int kDcnThreads = (kDcnFirstProd + kDcnProdWarps) * 32;
constexpr int kDcnProdThreads = kDcnProdWarps * 32;
constexpr int kDcnMaxStages = 6;
constexpr int kDcnJobs = 9 * 128;  // (tap, pixel) jobs per tile, one uint4 each

__host__ __device__ constexpr int dcn_smem_bytes(int stages, int cout) {
    return stages * (16384 + cout * 128) + 2 * kDcnJobs * 16 + cout * 4 + 1024;
}

__device__ __forceinline__ float dcn_fhfma(unsigned short a, unsigned short b, float c) {
    asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(c) : "h"(a), "h"(b));
    return c;
}

// blend of the four corners of eight channels: fp16 sample x fp16 (mask-scaled) weight + fp32 accumulator, rounded to fp16
__device__ __forceinline__ uint4 dcn_blend8(const uint4 (&u)[4], uint32_t w01, uint32_t w23) {
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t wp = (k < 2) ? w01 : w23;
        const unsigned short wk = static_cast<unsigned short>((k & 1) ? (wp >> 16) : (wp & 0xffffu));
        const uint32_t* h = reinterpret_cast<const uint32_t*>(&u[k]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            acc[2 * e] = dcn_fhfma(static_cast<unsigned short>(h[e] & 0xffffu), wk, acc[2 * e]);
            acc[2 * e + 1] = dcn_fhfma(static_cast<unsigned short>(h[e] >> 16), wk, acc[2 * e + 1]);
        }
    }
    uint4 out;
    __half2* ho = reinterpret_cast<__half2*>(&out);
#pragma unroll
    for (int e = 0; e < 4; ++e) ho[e] = __floats2half2_rn(acc[2 * e], acc[2 * e + 1]);
    return out;
}

// tuning aid (DV_DCN_DEBUG=3): the same blend on packed fp16 FMAs (fp16 accumulation: not the shipped arithmetic)
__device__ __forceinline__ uint4 dcn_blend8_h2(const uint4 (&u)[4], uint32_t w01, uint32_t w23) {
    uint4 out = make_uint4(0u, 0u, 0u, 0u);
    __half2* ho = reinterpret_cast<__half2*>(&out);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t wp = (k < 2) ? w01 : w23;
        const __half2 wv = (k & 1) ? __high2half2(*reinterpret_cast<const __half2*>(&wp)) : __low2half2(*reinterpret_cast<const __half2*>(&wp));
        const __half2* h = reinterpret_cast<const __half2*>(&u[k]);
#pragma unroll
        for (int e = 0; e < 4; ++e) ho[e] = __hfma2(h[e], wv, ho[e]);
    }
    return out;
}

__global__ void __launch_bounds__(kDcnThreads, 1)
dcn_fused_tcgen05(const __grid_constant__ DcnParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_a[kDcnMaxStages], full_b[kDcnMaxStages], empty[kDcnMaxStages], acc_full[2], acc_empty[2];
    __shared__ __align__(8) uint64_t tbl_full[2], tbl_empty[2];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.stages, cout = p.cout;
    const uint32_t b_slot = static_cast<uint32_t>(cout) * 128u;
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = smem_base;
    const uint32_t sB = sA + static_cast<uint32_t>(S) * 16384u;
    uint8_t* gen_base = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
    uint4* table = reinterpret_cast<uint4*>(gen_base + static_cast<size_t>(S) * (16384 + b_slot));
    float* s_bias = reinterpret_cast<float*>(table + 2 * kDcnJobs);

    for (int i = threadIdx.x; i < cout; i += kDcnThreads) s_bias[i] = p.bias ? __ldg(p.bias + i) : 0.f;
    if (threadIdx.x == 0) {
        for (int i = 0; i < S; ++i) {
            ptx::mbar_init(ptx::smem_u32(&full_a[i]), kDcnProdWarps);
            ptx::mbar_init(ptx::smem_u32(&full_b[i]), 1);
            ptx::mbar_init(ptx::smem_u32(&empty[i]), 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(ptx::smem_u32(&acc_full[i]), 1);
            ptx::mbar_init(ptx::smem_u32(&acc_empty[i]), 4);
            ptx::mbar_init(ptx::smem_u32(&tbl_full[i]), kDcnSetupWarps);
            ptx::mbar_init(ptx::smem_u32(&tbl_empty[i]), kDcnProdWarps);
        }
        ptx::fence_barrier_init();
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
    const int CB = p.C >> 6, KC = 9 * CB;  // 64-channel blocks per tap, chunks per tile
    const int tiles_per_img = p.tiles_x * p.tiles_y;

    if (warp == 0) {
        // ===================== weight producer (TMA) =====================
        if (ptx::elect_one_sync()) {
            uint32_t s = 0, ph = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                for (int kc = 0; kc < KC; ++kc) {
                    ptx::mbar_wait(ptx::smem_u32(&empty[s]), ph ^ 1u);
                    const uint32_t fb = ptx::smem_u32(&full_b[s]);
                    ptx::mbar_expect_tx(fb, b_slot);
                    ptx::tma_load_2d(sB + s * b_slot, &p.tmB, fb, kc * 64, 0);
                    if (++s == static_cast<uint32_t>(S)) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (ptx::elect_one_sync()) {
            const uint32_t idesc = ptx::make_idesc_f16_m128(static_cast<uint32_t>(cout));
            uint32_t s = 0, ph = 0, t = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++t) {
                const uint32_t buf = t & 1u;
                ptx::mbar_wait(ptx::smem_u32(&acc_empty[buf]), ((t >> 1) & 1u) ^ 1u);
                ptx::tc_fence_after();
                const uint32_t d = tmem_base + buf * static_cast<uint32_t>(cout);
                for (int kc = 0; kc < KC; ++kc) {
                    ptx::mbar_wait(ptx::smem_u32(&full_b[s]), ph);
                    ptx::mbar_wait(ptx::smem_u32(&full_a[s]), ph);
                    ptx::tc_fence_after();
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::umma_f16_ss(d, ptx::make_kmajor_desc(sA + s * 16384u + k * 32, 128),
                                         ptx::make_kmajor_desc(sB + s * b_slot + k * 32, 128), idesc, (kc | k) != 0 ? 1u : 0u);
                    ptx::umma_commit(ptx::smem_u32(&empty[s]));
                    if (++s == static_cast<uint32_t>(S)) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
                ptx::umma_commit(ptx::smem_u32(&acc_full[buf]));
            }
        }
    } else if (warp < kDcnFirstSetup) {
        // ===================== epilogue: bias + ReLU + NHWC store =====================
        const int q = warp & 3, row = q * 32 + lane;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        uint32_t t = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++t) {
            const uint32_t buf = t & 1u;
            const int img = tile / tiles_per_img, rem = tile - img * tiles_per_img;
            const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
            const int y = ty * kDcnTH + (row >> 4), x = tx * kDcnTW + (row & 15);
            const bool valid = y < p.H && x < p.W;
            __half* orow = p.out + (static_cast<size_t>(img) * p.H * p.W + static_cast<size_t>(y) * p.W + x) * p.ldo;
            ptx::mbar_wait(ptx::smem_u32(&acc_full[buf]), (t >> 1) & 1u);
            ptx::tc_fence_after();
            for (int c0 = 0; c0 < cout; c0 += 32) {
                uint32_t v[32];
                ptx::tmem_ld_32x32b_x32(t_lane + buf * static_cast<uint32_t>(cout) + static_cast<uint32_t>(c0), v);
                ptx::tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint4 o;
                        __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float a = __uint_as_float(v[g * 8 + 2 * i]) + s_bias[c0 + g * 8 + 2 * i];
                            float b = __uint_as_float(v[g * 8 + 2 * i + 1]) + s_bias[c0 + g * 8 + 2 * i + 1];
                            if (p.act == ACT_RELU) {
                                a = fmaxf(a, 0.f);
                                b = fmaxf(b, 0.f);
                            }
                            ho[i] = __floats2half2_rn(a, b);
                        }
                        *reinterpret_cast<uint4*>(orow + c0 + g * 8) = o;
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&acc_empty[buf]));
        }
    } else if (warp < kDcnFirstProd) {
        // ===================== set-up: the (tap, pixel) job table of each tile, one tile ahead of the gather =====================
        // entry = {byte offset of the clamped corner (y0, x0) in its image, bit 0 / 1 = the x / y neighbour is a different
        // pixel (not clamped away), half2(w00, w01), half2(w10, w11)}; weights are mask-scaled and zero where a corner falls
        // outside the image (torchvision bilinear_interpolate), fp16 like k_dcn_im2col's
        const int r = threadIdx.x - kDcnFirstSetup * 32;
        const int H = p.H, W = p.W;
        uint32_t t = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++t) {
            const uint32_t tb = t & 1u;
            const int img = tile / tiles_per_img, rem = tile - img * tiles_per_img;
            const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
            const int y = ty * kDcnTH + (r >> 4), x = tx * kDcnTW + (r & 15);
            const bool valid = y < H && x < W;
            float o[28];
            if (valid) {
                const float4* op = reinterpret_cast<const float4*>(p.om + (static_cast<size_t>(img) * H * W + static_cast<size_t>(y) * W + x) * 32);
#pragma unroll
                for (int i = 0; i < 7; ++i) {
                    const float4 f = __ldg(op + i);
                    o[4 * i] = f.x;
                    o[4 * i + 1] = f.y;
                    o[4 * i + 2] = f.z;
                    o[4 * i + 3] = f.w;
                }
            }
            ptx::mbar_wait(ptx::smem_u32(&tbl_empty[tb]), ((t >> 1) & 1u) ^ 1u);
            uint4* tbl = table + tb * kDcnJobs + r;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                uint4 e = make_uint4(0u, 0u, 0u, 0u);
                if (valid) {
                    const float dy = o[2 * tap], dx = o[2 * tap + 1];
                    const float mask = 1.f / (1.f + expf(-o[18 + tap]));
                    const int ky = tap / 3, kx = tap - 3 * ky;
                    const float py = static_cast<float>(y + ky - 1) + dy;
                    const float px = static_cast<float>(x + kx - 1) + dx;
                    const bool inside = py > -1.f && py < static_cast<float>(H) && px > -1.f && px < static_cast<float>(W);
                    const float fy = floorf(py), fx = floorf(px);
                    const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
                    const float ly = py - fy, lx = px - fx, hy = 1.f - ly, hx = 1.f - lx;
                    const bool oky0 = inside && y0 >= 0 && y0 <= H - 1, oky1 = inside && y0 + 1 >= 0 && y0 + 1 <= H - 1;
                    const bool okx0 = x0 >= 0 && x0 <= W - 1, okx1 = x0 + 1 >= 0 && x0 + 1 <= W - 1;
                    const float w00 = (oky0 && okx0) ? hy * hx * mask : 0.f, w01 = (oky0 && okx1) ? hy * lx * mask : 0.f;
                    const float w10 = (oky1 && okx0) ? ly * hx * mask : 0.f, w11 = (oky1 && okx1) ? ly * lx * mask : 0.f;
                    const int cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y0 + 1, 0), H - 1);
                    const int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x0 + 1, 0), W - 1);
                    e.x = static_cast<uint32_t>(cy0 * W + cx0) * static_cast<uint32_t>(p.ldi) * 2u;
                    e.y = static_cast<uint32_t>(cx1 - cx0) | (static_cast<uint32_t>(cy1 - cy0) << 1);
                    const __half2 h01 = __floats2half2_rn(w00, w01), h23 = __floats2half2_rn(w10, w11);
                    e.z = *reinterpret_cast<const uint32_t*>(&h01);
                    e.w = *reinterpret_cast<const uint32_t*>(&h23);
                }
                tbl[tap * 128] = e;
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&tbl_full[tb]));
        }
    } else {
        // ===================== gather producers =====================
        // thread = (row, 16-byte channel group): the eight lanes of a row read one whole 128-byte line per corner -- one L1
        // wavefront per (pixel, tap, corner), the floor of this gather (two half-line requests per corner measured 2x slower)
        const int ptid = threadIdx.x - kDcnFirstProd * 32;
        const int c8 = ptid & 7, rr = ptid >> 3;
        constexpr int ROWS = kDcnProdThreads / 8, ITEMS = 128 / ROWS;
        const uint32_t sxb = static_cast<uint32_t>(p.ldi) * 2u, syb = static_cast<uint32_t>(p.W) * sxb;
        uint32_t s = 0, ph = 0, t = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++t) {
            const uint32_t tb = t & 1u;
            const int img = tile / tiles_per_img;
            const uint4* tbl = table + tb * kDcnJobs + rr;
            const char* img_base = reinterpret_cast<const char*>(p.in + static_cast<size_t>(img) * p.H * p.W * p.ldi + c8 * 8);
            ptx::mbar_wait(ptx::smem_u32(&tbl_full[tb]), (t >> 1) & 1u);
            int tap = 0, cb = 0;
            for (int kc = 0; kc < KC; ++kc) {
                uint4 u[ITEMS][4];
                uint32_t w01[ITEMS], w23[ITEMS];
                // all corner loads of this thread's rows are in flight before the stage wait
#pragma unroll
                for (int i = 0; i < ITEMS; ++i) {
                    const uint4 e = tbl[tap * 128 + i * ROWS];
                    w01[i] = e.z;
                    w23[i] = e.w;
                    const uint32_t o00 = e.x + static_cast<uint32_t>(cb) * 128u;
                    const uint32_t o01 = o00 + ((e.y & 1u) ? sxb : 0u);
                    const uint32_t o10 = o00 + ((e.y & 2u) ? syb : 0u);
                    const uint32_t o11 = o10 + ((e.y & 1u) ? sxb : 0u);
                    if (p.dbg == 1) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) u[i][k] = make_uint4(0, 0, 0, 0);
                    } else {
                        u[i][0] = __ldg(reinterpret_cast<const uint4*>(img_base + o00));
                        u[i][1] = __ldg(reinterpret_cast<const uint4*>(img_base + o01));
                        u[i][2] = __ldg(reinterpret_cast<const uint4*>(img_base + o10));
                        u[i][3] = __ldg(reinterpret_cast<const uint4*>(img_base + o11));
                    }
                }
                ptx::mbar_wait(ptx::smem_u32(&empty[s]), ph ^ 1u);
#pragma unroll
                for (int i = 0; i < ITEMS; ++i) {
                    const int r = rr + i * ROWS;
                    const uint4 out = (p.dbg == 2) ? u[i][0] : (p.dbg == 3) ? dcn_blend8_h2(u[i], w01[i], w23[i]) : dcn_blend8(u[i], w01[i], w23[i]);
                    const uint32_t addr = sA + s * 16384u + static_cast<uint32_t>((r >> 3) * 1024 + (r & 7) * 128) +
                                          ((static_cast<uint32_t>(c8) ^ static_cast<uint32_t>(r & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(out.x), "r"(out.y), "r"(out.z), "r"(out.w)
                                 : "memory");
                }
                ptx::fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async proxy
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&full_a[s]));
                if (++cb == CB) {
                    cb = 0;
                    ++tap;
                }
                if (++s == static_cast<uint32_t>(S)) {
                    s = 0;
                    ph ^= 1u;
                }
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&tbl_empty[tb]));  // this warp has read its last entry of the table
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
