// Non-GEMM kernels of the ConvNextViT text-line recogniser:
//   k_cnv_patchify_ln  RGB->gray + 4x4 s4 patchify conv (1->96) + LayerNorm      (HF ConvNextEmbeddings)
//   k_dwconv7_ln       depthwise 7x7 + bias + LayerNorm(eps 1e-6), fp32 stream -> fp16 GEMM operand
//   k_ln_rows          row LayerNorm / cast with the (2,1) down-sampling re-layout and the 3x75->201 stitch
//   k_attn75           softmax(QK^T)V for the 75-token ViT (3 heads x 64), one CTA per chunk
// Reference: convnext_vit/modeling_convnext_vit.py:37-45, modeling_convnext.py:28-80, modeling_vit.py:32-180 and the
// HF blocks they instantiate (ConvNextLayer: dwconv -> LN -> pwconv1 -> GELU -> pwconv2 -> layer_scale -> +res).
// The residual stream stays fp32 in HBM (it is only C floats per token; the 4C-wide hidden tensors are fp16).
#include <stdlib.h>

#include "engine.h"

namespace dv {

namespace {

// fp32x (split-fp16) operands: a value is stored as hi = fp16(v) and lo = fp16(v - hi); hi + lo carries ~21 significand bits
// and the consumer GEMM forms A_hi W_hi + A_hi W_lo + A_lo W_hi (igemm_host.cu plan_linear, ConvSpec::split).
__device__ __forceinline__ void store_split(__half* hi_ptr, int lo_off, float v) {
    const __half h = __float2half_rn(v);
    *hi_ptr = h;
    if (lo_off) hi_ptr[lo_off] = __float2half_rn(v - __half2float(h));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------
// chunks fp32 NCHW [B,3,32,300] in [0,1]  ->  x fp32 [B,8,75,96]
// gray = R*0.2989 + G*0.5870 + B*0.1140 (modeling_convnext_vit.py:40), conv 4x4 stride 4, LN eps 1e-6.
// U8 variant: `in` is uint8 HWC crops [B/3, 32, crop_w, 3] (height 32, zero padded to a common width <= 804);
// chunk k of a crop is the column window [252k, 252k+300) with zeros beyond crop_w, and /255 is applied first
// -- OCRRecognitionPreprocessor.__call__ (ocr_recognition/processor_ocr_recognition.py:57-61, 104-112) fused in.
template <bool U8>
__global__ void __launch_bounds__(256)
k_cnv_patchify_ln(const void* __restrict__ in_raw, int crop_w, int B, const float* __restrict__ w /*[16][96]*/,
                  const float* __restrict__ bias, const float* __restrict__ lnw, const float* __restrict__ lnb,
                  float* __restrict__ out) {
    __shared__ float sw[16 * 96];
    for (int i = threadIdx.x; i < 16 * 96; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long pix = static_cast<long long>(blockIdx.x) * 8 + warp;
    if (pix >= static_cast<long long>(B) * 600) return;
    const int b = static_cast<int>(pix / 600);
    const int r = static_cast<int>(pix - static_cast<long long>(b) * 600);
    const int py = r / 75, px = r - py * 75;
    float g = 0.f;
    if (lane < 16) {
        const int iy = py * 4 + (lane >> 2), ix = px * 4 + (lane & 3);
        float c0, c1, c2;
        if constexpr (U8) {
            const int crop = b / 3, col = 252 * (b - crop * 3) + ix;
            c0 = c1 = c2 = 0.f;
            if (col < crop_w) {
                const uint8_t* ip = reinterpret_cast<const uint8_t*>(in_raw) +
                                    ((static_cast<long long>(crop) * 32 + iy) * crop_w + col) * 3;
                c0 = __fdiv_rn(static_cast<float>(ip[0]), 255.f);
                c1 = __fdiv_rn(static_cast<float>(ip[1]), 255.f);
                c2 = __fdiv_rn(static_cast<float>(ip[2]), 255.f);
            }
        } else {
            const float* ip = reinterpret_cast<const float*>(in_raw) + (static_cast<long long>(b) * 3 * 32 + iy) * 300 + ix;
            c0 = ip[0];
            c1 = ip[32 * 300];
            c2 = ip[2 * 32 * 300];
        }
        // same evaluation order as the reference expression: (R*a + G*b) + B*c, no FMA contraction
        g = __fadd_rn(__fadd_rn(__fmul_rn(c0, 0.2989f), __fmul_rn(c1, 0.5870f)), __fmul_rn(c2, 0.1140f));
    }
    float acc[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) acc[j] = bias[lane + 32 * j];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const float gk = __shfl_sync(0xffffffffu, g, k);
#pragma unroll
        for (int j = 0; j < 3; ++j) acc[j] = fmaf(gk, sw[k * 96 + lane + 32 * j], acc[j]);
    }
    const float mean = warp_sum(acc[0] + acc[1] + acc[2]) * (1.f / 96.f);
    float d0 = acc[0] - mean, d1 = acc[1] - mean, d2 = acc[2] - mean;
    const float var = warp_sum(d0 * d0 + d1 * d1 + d2 * d2) * (1.f / 96.f);
    const float rstd = rsqrtf(var + 1e-6f);
    float* op = out + pix * 96;
    op[lane] = d0 * rstd * lnw[lane] + lnb[lane];
    op[lane + 32] = d1 * rstd * lnw[lane + 32] + lnb[lane + 32];
    op[lane + 64] = d2 * rstd * lnw[lane + 64] + lnb[lane + 64];
}

// ------------------------------------------------------------------------------------------------
// Depthwise 7x7 (pad 3) + bias + LayerNorm over C.  x fp32 [B,H,75,C] -> h fp16 [B,H,75,C].
// One CTA = one chunk b, one strip of kStrip output columns, all H rows.  The (H x (kStrip+6) x C) input
// tile is staged once in shared memory; thread (y, c) produces the kStrip outputs of its row with a
// register-blocked sliding window (each staged value feeds up to 7 outputs), then the tile memory is
// re-used to exchange the conv outputs for the per-pixel LayerNorm done by whole warps.
constexpr int kStrip = 15;  // 75 = 5 strips
constexpr int kTileW = kStrip + 6;

template <int C, int H>
__global__ void __launch_bounds__(C* H)
k_dwconv7_ln(const float* __restrict__ x, const float* __restrict__ w /*[49][C]*/, const float* __restrict__ bias,
             const float* __restrict__ lnw, const float* __restrict__ lnb, __half* __restrict__ out, int split) {
    extern __shared__ float tile[];  // [H][kTileW][C]
    const int b = blockIdx.x / 5;
    const int x0 = (blockIdx.x - b * 5) * kStrip;
    const int tid = threadIdx.x;
    const float* xb = x + static_cast<long long>(b) * H * 75 * C;
    // ---- stage (zero-filled outside the image); C % 4 == 0 -> float4
    constexpr int C4 = C / 4;
    for (int i = tid; i < H * kTileW * C4; i += C * H) {
        const int c4 = i % C4;
        const int col = (i / C4) % kTileW;
        const int row = i / (C4 * kTileW);
        const int gx = x0 - 3 + col;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gx >= 0 && gx < 75) v = *reinterpret_cast<const float4*>(xb + (static_cast<long long>(row) * 75 + gx) * C + c4 * 4);
        *reinterpret_cast<float4*>(tile + (static_cast<long long>(row) * kTileW + col) * C + c4 * 4) = v;
    }
    __syncthreads();
    const int c = tid % C;
    const int y = tid / C;
    float acc[kStrip];
    const float bv = bias[c];
#pragma unroll
    for (int i = 0; i < kStrip; ++i) acc[i] = bv;
#pragma unroll
    for (int r = 0; r < 7; ++r) {
        const int iy = y + r - 3;
        if (iy < 0 || iy >= H) continue;  // zero padding rows
        float wr[7];
#pragma unroll
        for (int s = 0; s < 7; ++s) wr[s] = __ldg(w + (r * 7 + s) * C + c);
        const float* trow = tile + static_cast<long long>(iy) * kTileW * C + c;
#pragma unroll
        for (int col = 0; col < kTileW; ++col) {
            const float v = trow[col * C];
#pragma unroll
            for (int s = 0; s < 7; ++s) {
                const int o = col - s;  // output column fed by tile column `col` through tap s
                if (o >= 0 && o < kStrip) acc[o] = fmaf(v, wr[s], acc[o]);
            }
        }
    }
    __syncthreads();  // everyone is done reading the tile -> reuse it as [H][kStrip][C] conv outputs
#pragma unroll
    for (int i = 0; i < kStrip; ++i) tile[(static_cast<long long>(y) * kStrip + i) * C + c] = acc[i];
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    constexpr int NW = C * H / 32;
    constexpr int CPL = C / 32;
    for (int p = warp; p < H * kStrip; p += NW) {
        const float* tp = tile + static_cast<long long>(p) * C;
        float v[CPL];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            v[j] = tp[lane + 32 * j];
            s += v[j];
        }
        const float mean = warp_sum(s) * (1.f / C);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            v[j] -= mean;
            q += v[j] * v[j];
        }
        const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + 1e-6f);
        const int py = p / kStrip, pxs = p - py * kStrip;
        // split: rows of [hi(C) | lo(C)]
        __half* op = out + ((static_cast<long long>(b) * H + py) * 75 + x0 + pxs) * (split ? 2 * C : C);
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            const int cc = lane + 32 * j;
            store_split(op + cc, split ? C : 0, v[j] * rstd * __ldg(lnw + cc) + __ldg(lnb + cc));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Row LayerNorm (or plain cast) fp32 [rows, C] -> fp16 with an output row mapping:
//   LN_IDENT  out row = in row
//   LN_DOWN   in [B,H,75,C] -> out [B,H/2,75,2C]: the (2,1)/(2,1) down-sampling conv becomes a flat GEMM with
//             K = 2C (kernel row p = y&1 selects the half)           (modeling_convnext.py:44-53)
//   LN_STITCH in [3n,75,C] -> out [n,201,C]: chunk0[0:69] | chunk1[6:69] | chunk2[6:75]  (modeling_vit.py:135-139)
enum { LN_IDENT = 0, LN_DOWN = 1, LN_STITCH = 2 };

template <int C>
__global__ void __launch_bounds__(256)
k_ln_rows(const float* __restrict__ in, long long rows, const float* __restrict__ lnw, const float* __restrict__ lnb,
          float eps, int normalise, int map, int H, __half* __restrict__ out, int split) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = static_cast<long long>(blockIdx.x) * 8 + warp;
    if (row >= rows) return;
    long long orow = row;
    int ocoff = 0, old = C;
    if (map == LN_DOWN) {
        const int xq = static_cast<int>(row % 75);
        const long long by = row / 75;  // b*H + y
        const int y = static_cast<int>(by % H);
        const long long bb = by / H;
        orow = (bb * (H / 2) + (y >> 1)) * 75 + xq;
        ocoff = (y & 1) * C;
        old = 2 * C;
    } else if (map == LN_STITCH) {
        const int t = static_cast<int>(row % 75);
        const long long chunk = row / 75;
        const int k = static_cast<int>(chunk % 3);
        const long long n = chunk / 3;
        int pos;
        if (k == 0) {
            if (t >= 69) return;
            pos = t;
        } else if (k == 1) {
            if (t < 6 || t >= 69) return;
            pos = 69 + t - 6;
        } else {
            if (t < 6) return;
            pos = 132 + t - 6;
        }
        orow = n * 201 + pos;
    }
    constexpr int CPL = C / 32;
    const float* ip = in + row * C;
    float v[CPL];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
        v[j] = ip[lane + 32 * j];
        s += v[j];
    }
    const int lo_off = split ? old : 0;  // split: output rows are [hi(old) | lo(old)]
    __half* op = out + orow * (split ? 2 * old : old) + ocoff;
    if (normalise) {
        const float mean = warp_sum(s) * (1.f / C);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            v[j] -= mean;
            q += v[j] * v[j];
        }
        const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + eps);
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            const int cc = lane + 32 * j;
            store_split(op + cc, lo_off, v[j] * rstd * __ldg(lnw + cc) + __ldg(lnb + cc));
        }
    } else {
#pragma unroll
        for (int j = 0; j < CPL; ++j) store_split(op + lane + 32 * j, lo_off, v[j]);
    }
}

// ------------------------------------------------------------------------------------------------
// ViT self-attention over the 75 tokens of one chunk: qkv fp16 [B*75, 576] (q|k|v, head-major 3 x 64; the
// 1/sqrt(64) scale is folded into the packed q weights) -> ctx fp16 [B*75, 192].
// One CTA per chunk, 3 warp-triples of 96 threads = one (head, query) per thread; K and V of all heads are
// staged in shared memory as fp32, scores live in a thread-private shared column.
constexpr int kAttnThreads = 288;
constexpr int kAttnSmem = (2 * 75 * 192 + 75 * kAttnThreads) * 4;

// SPLIT (fp32x mode): qkv rows are [hi(576) | lo(576)] and ctx rows [hi(192) | lo(192)]; q, k, v = hi + lo in fp32.
template <bool SPLIT>
__global__ void __launch_bounds__(kAttnThreads, 1)
k_attn75(const __half* __restrict__ qkv, __half* __restrict__ ctx) {
    extern __shared__ float sm[];
    float* sK = sm;                 // [75][192]
    float* sV = sm + 75 * 192;      // [75][192]
    float* sS = sm + 2 * 75 * 192;  // [75][288]
    const int b = blockIdx.x;
    constexpr int RS = SPLIT ? 1152 : 576;  // halves per qkv row
    const __half* base = qkv + static_cast<long long>(b) * 75 * RS;
    for (int i = threadIdx.x; i < 75 * 48; i += kAttnThreads) {  // 48 x 8 halves = K|V of one token
        const int t = i / 48, c8 = i - t * 48;
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(base + t * RS + 192 + c8 * 8));
        const __half2* h = reinterpret_cast<const __half2*>(&u);
        float* dst = (c8 < 24 ? sK + t * 192 + c8 * 8 : sV + t * 192 + (c8 - 24) * 8);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(h[e]);
            dst[2 * e] = f.x;
            dst[2 * e + 1] = f.y;
        }
        if constexpr (SPLIT) {
            const uint4 ul = __ldg(reinterpret_cast<const uint4*>(base + t * RS + 576 + 192 + c8 * 8));
            const __half2* hl = reinterpret_cast<const __half2*>(&ul);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __half22float2(hl[e]);
                dst[2 * e] += f.x;
                dst[2 * e + 1] += f.y;
            }
        }
    }
    __syncthreads();
    const int head = threadIdx.x / 96;
    const int qi = threadIdx.x - head * 96;
    if (qi >= 75) return;
    float q[64];
    {
        const uint4* qp = reinterpret_cast<const uint4*>(base + qi * RS + head * 64);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint4 u = __ldg(qp + i);
            const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __half22float2(h[e]);
                q[i * 8 + 2 * e] = f.x;
                q[i * 8 + 2 * e + 1] = f.y;
            }
            if constexpr (SPLIT) {
                const uint4 ul = __ldg(qp + 72 + i);  // + 576 halves
                const __half2* hl = reinterpret_cast<const __half2*>(&ul);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 f = __half22float2(hl[e]);
                    q[i * 8 + 2 * e] += f.x;
                    q[i * 8 + 2 * e + 1] += f.y;
                }
            }
        }
    }
    float* myS = sS + threadIdx.x;
    float mx = -INFINITY;
    for (int j = 0; j < 75; ++j) {
        const float4* kp = reinterpret_cast<const float4*>(sK + j * 192 + head * 64);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int d = 0; d < 16; ++d) {
            const float4 k4 = kp[d];
            a0 = fmaf(q[4 * d], k4.x, a0);
            a1 = fmaf(q[4 * d + 1], k4.y, a1);
            a2 = fmaf(q[4 * d + 2], k4.z, a2);
            a3 = fmaf(q[4 * d + 3], k4.w, a3);
        }
        const float s = (a0 + a1) + (a2 + a3);
        myS[j * kAttnThreads] = s;
        mx = fmaxf(mx, s);
    }
    float o[64];
#pragma unroll
    for (int d = 0; d < 64; ++d) o[d] = 0.f;
    float denom = 0.f;
    for (int j = 0; j < 75; ++j) {
        const float pj = expf(myS[j * kAttnThreads] - mx);
        denom += pj;
        const float4* vp = reinterpret_cast<const float4*>(sV + j * 192 + head * 64);
#pragma unroll
        for (int d = 0; d < 16; ++d) {
            const float4 v4 = vp[d];
            o[4 * d] = fmaf(pj, v4.x, o[4 * d]);
            o[4 * d + 1] = fmaf(pj, v4.y, o[4 * d + 1]);
            o[4 * d + 2] = fmaf(pj, v4.z, o[4 * d + 2]);
            o[4 * d + 3] = fmaf(pj, v4.w, o[4 * d + 3]);
        }
    }
    const float inv = 1.f / denom;
    uint4* op = reinterpret_cast<uint4*>(ctx + (static_cast<long long>(b) * 75 + qi) * (SPLIT ? 384 : 192) + head * 64);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint4 u, ul;
        __half2* h = reinterpret_cast<__half2*>(&u);
        __half2* hl = reinterpret_cast<__half2*>(&ul);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float v0 = o[i * 8 + 2 * e] * inv, v1 = o[i * 8 + 2 * e + 1] * inv;
            h[e] = __floats2half2_rn(v0, v1);
            if constexpr (SPLIT) {
                const float2 f = __half22float2(h[e]);
                hl[e] = __floats2half2_rn(v0 - f.x, v1 - f.y);
            }
        }
        op[i] = u;
        if constexpr (SPLIT) op[24 + i] = ul;  // + 192 halves
    }
}

// ------------------------------------------------------------------------------------------------
// k_attn75_mma: the same 75-token, 3-head attention on the legacy warp-level tensor path (mma.sync m16n8k16; the
// problem -- 75 x 75 x 64 per head -- is far below one tcgen05 tile, and the CUDA-core version above had become the
// second largest kernel of the recogniser).  One CTA per chunk, 15 warps: warp = (head, 16-query tile).  Q, K, V are
// staged once as fp16 in padded shared memory; S = QK^T stays in registers (10 n-tiles), softmax in fp32 on the quad,
// P is re-used in place as the fp16 A fragments of the PV product (the accumulator layout of m16n8 equals the A layout
// of m16k16), V fragments come from ldmatrix.trans.
constexpr int kAttnRow = 200;                 // halves per staged row (192 + 8 pad: conflict-free ldmatrix)
constexpr int kAttnMmaThreads = 15 * 32;
constexpr int kAttnMmaSmem = 3 * 80 * kAttnRow * 2;

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// Persistent and double-buffered: a CTA walks chunks b = blockIdx.x, + gridDim.x, ... and the q | k | v rows of the NEXT chunk
// arrive by cp.async while the current one is computed.  (One CTA per chunk with a synchronous staging loop spent most of
// its ~9 us on the load round trip: 128 registers x 480 threads allow one CTA per SM, so nothing else covered it.)
__global__ void __launch_bounds__(kAttnMmaThreads, 1)
k_attn75_mma(const __half* __restrict__ qkv, __half* __restrict__ ctx, int B) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    constexpr int kStageHalves = 3 * 80 * kAttnRow;
    __half* stage0 = reinterpret_cast<__half*>(sm_raw);
    // rows 75..79 of every staged matrix are zero padding: written once, never touched by the copies
    for (int i = threadIdx.x; i < 2 * 3 * 5 * (kAttnRow / 8); i += kAttnMmaThreads) {
        const int m = i / (5 * (kAttnRow / 8)), r = i - m * (5 * (kAttnRow / 8));
        *reinterpret_cast<uint4*>(stage0 + m * 80 * kAttnRow + 75 * kAttnRow + r * 8) = make_uint4(0, 0, 0, 0);
    }
    auto prefetch = [&](int chunk, int st) {
        if (chunk < B) {
            const __half* base = qkv + static_cast<long long>(chunk) * 75 * 576;
            __half* sq = stage0 + st * kStageHalves;
            for (int i = threadIdx.x; i < 75 * 72; i += kAttnMmaThreads) {  // 72 x 16 B per token row (q | k | v)
                const int t = i / 72, c8 = i - t * 72;
                const __half* dst = sq + (c8 / 24) * 80 * kAttnRow + t * kAttnRow + (c8 % 24) * 8;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst))),
                             "l"(base + t * 576 + c8 * 8)
                             : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");  // an empty group keeps the wait count uniform
    };
    prefetch(blockIdx.x, 0);
    int it = 0;
    for (int b = blockIdx.x; b < B; b += gridDim.x, ++it) {
    const int st = it & 1;
    prefetch(b + gridDim.x, st ^ 1);
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    __half* sQ = stage0 + st * kStageHalves;
    __half* sK = sQ + 80 * kAttnRow;
    __half* sV = sK + 80 * kAttnRow;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int head = warp / 5, mt = warp - head * 5;
    const uint32_t q_s = static_cast<uint32_t>(__cvta_generic_to_shared(sQ));
    const uint32_t k_s = static_cast<uint32_t>(__cvta_generic_to_shared(sK));
    const uint32_t v_s = static_cast<uint32_t>(__cvta_generic_to_shared(sV));
    // ---- S = Q K^T (Q pre-scaled by 1/8 in the packed weights)
    uint32_t qa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const int row = mt * 16 + (lane & 15), col = head * 64 + ks * 16 + (lane >> 4) * 8;
        ldsm_x4(q_s + (row * kAttnRow + col) * 2, qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
    }
    float s[10][4];
#pragma unroll
    for (int nt = 0; nt < 10; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {  // 32 head dims per ldmatrix.x4
            uint32_t b0, b1, b2, b3;
            const int row = nt * 8 + (lane & 7), col = head * 64 + kp * 32 + (lane >> 3) * 8;
            ldsm_x4(k_s + (row * kAttnRow + col) * 2, b0, b1, b2, b3);
            mma16816(s[nt], qa[2 * kp], b0, b1);
            mma16816(s[nt], qa[2 * kp + 1], b2, b3);
        }
    }
    // ---- softmax over the 75 keys (rows r0 = lane/4 and r0 + 8 of the tile; a row lives on the 4 lanes of a quad)
    const int cq = 2 * (lane & 3);
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 10; ++nt) {
        if (nt * 8 + cq >= 75) s[nt][0] = s[nt][2] = -INFINITY;
        if (nt * 8 + cq + 1 >= 75) s[nt][1] = s[nt][3] = -INFINITY;
        m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
        m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.f, l1 = 0.f;
    uint32_t pa[5][4];
#pragma unroll
    for (int nt = 0; nt < 10; ++nt) {
        const float p0 = __expf(s[nt][0] - m0), p1 = __expf(s[nt][1] - m0);
        const float p2 = __expf(s[nt][2] - m1), p3 = __expf(s[nt][3] - m1);
        l0 += p0 + p1;
        l1 += p2 + p3;
        pa[nt >> 1][(nt & 1) * 2] = pack_h2(p0, p1);
        pa[nt >> 1][(nt & 1) * 2 + 1] = pack_h2(p2, p3);
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    // ---- O = P V
    float o[8][4];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 5; ++kk) {
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {  // two 8-wide dim tiles per ldmatrix.x4.trans
            uint32_t b0, b1, b2, b3;
            const int row = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, col = head * 64 + dp * 16 + (lane >> 4) * 8;
            ldsm_x4_trans(v_s + (row * kAttnRow + col) * 2, b0, b1, b2, b3);
            mma16816(o[2 * dp], pa[kk], b0, b1);
            mma16816(o[2 * dp + 1], pa[kk], b2, b3);
        }
    }
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    const int r0 = mt * 16 + (lane >> 2), r1 = r0 + 8;
    __half* out = ctx + static_cast<long long>(b) * 75 * 192 + head * 64 + cq;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
        if (r0 < 75) *reinterpret_cast<__half2*>(out + r0 * 192 + dt * 8) = __floats2half2_rn(o[dt][0] * i0, o[dt][1] * i0);
        if (r1 < 75) *reinterpret_cast<__half2*>(out + r1 * 192 + dt * 8) = __floats2half2_rn(o[dt][2] * i1, o[dt][3] * i1);
    }
    __syncthreads();  // this stage is overwritten by the prefetch of the iteration after next
    }
}

template <int C, int H>
int launch_dw(Engine* e, const float* x, int B, const float* w, const float* b, const float* lnw, const float* lnb,
              __half* out, const char* layer, int split) {
    const size_t smem = static_cast<size_t>(H) * kTileW * C * sizeof(float);
    static DeviceOnce attr_once;
    if (attr_once.need(e->device)) {
        DV_CUDA(e, cudaFuncSetAttribute(k_dwconv7_ln<C, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        attr_once.mark(e->device);
    }
    const double elems = static_cast<double>(B) * H * 75 * C;
    e->launch_begin("k_dwconv7_ln", layer, 2.0 * 49 * elems, elems * (4 + (split ? 4 : 2)));
    k_dwconv7_ln<C, H><<<B * 5, C * H, smem, e->stream>>>(x, w, b, lnw, lnb, out, split);
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

}  // namespace

int op_cnv_patchify_ln(Engine* e, const float* chunks, const uint8_t* crops_u8, int crop_w, int B, const float* w,
                       const float* bias, const float* lnw, const float* lnb, float* out) {
    const long long pix = static_cast<long long>(B) * 600;
    const double in_bytes = crops_u8 ? static_cast<double>(B / 3) * 32 * crop_w * 3 : static_cast<double>(B) * 3 * 32 * 300 * 4;
    e->launch_begin("k_cnv_patchify_ln", "patchify", 2.0 * 16 * 96 * pix, in_bytes + pix * 96.0 * 4);
    if (crops_u8)
        k_cnv_patchify_ln<true><<<static_cast<int>((pix + 7) / 8), 256, 0, e->stream>>>(crops_u8, crop_w, B, w, bias, lnw, lnb, out);
    else
        k_cnv_patchify_ln<false><<<static_cast<int>((pix + 7) / 8), 256, 0, e->stream>>>(chunks, 0, B, w, bias, lnw, lnb, out);
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

int op_dwconv7_ln(Engine* e, const float* x, int B, int H, int C, const float* w, const float* b, const float* lnw,
                  const float* lnb, __half* out, const char* layer, int split) {
    if (C == 96 && H == 8) return launch_dw<96, 8>(e, x, B, w, b, lnw, lnb, out, layer, split);
    if (C == 192 && H == 4) return launch_dw<192, 4>(e, x, B, w, b, lnw, lnb, out, layer, split);
    if (C == 256 && H == 2) return launch_dw<256, 2>(e, x, B, w, b, lnw, lnb, out, layer, split);
    if (C == 512 && H == 1) return launch_dw<512, 1>(e, x, B, w, b, lnw, lnb, out, layer, split);
    return set_err(e, DV_ERR_UNSUPPORTED, "dwconv7_ln: unsupported (C=%d, H=%d)", C, H);
}

int op_ln_rows(Engine* e, const float* in, long long rows, int C, const float* lnw, const float* lnb, float eps,
               int normalise, int map, int H, __half* out, const char* layer, int split) {
    const int grid = static_cast<int>((rows + 7) / 8);
    e->launch_begin("k_ln_rows", layer, 0.0, static_cast<double>(rows) * C * (split ? 8 : 6));
    switch (C) {
        case 96: k_ln_rows<96><<<grid, 256, 0, e->stream>>>(in, rows, lnw, lnb, eps, normalise, map, H, out, split); break;
        case 192: k_ln_rows<192><<<grid, 256, 0, e->stream>>>(in, rows, lnw, lnb, eps, normalise, map, H, out, split); break;
        case 256: k_ln_rows<256><<<grid, 256, 0, e->stream>>>(in, rows, lnw, lnb, eps, normalise, map, H, out, split); break;
        case 512: k_ln_rows<512><<<grid, 256, 0, e->stream>>>(in, rows, lnw, lnb, eps, normalise, map, H, out, split); break;
        default: e->launch_end(); return set_err(e, DV_ERR_UNSUPPORTED, "ln_rows: C=%d", C);
    }
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

int op_attn75(Engine* e, const __half* qkv, int B, __half* ctx, const char* layer, int split) {
    static DeviceOnce attr_once;
    static const bool mma_env = !(getenv("DV_ATTN_SIMT") && atoi(getenv("DV_ATTN_SIMT")));
    const bool use_mma = mma_env && !split;  // fp32x mode: q, k, v = hi + lo in fp32 on the CUDA cores
    if (attr_once.need(e->device)) {
        DV_CUDA(e, cudaFuncSetAttribute(k_attn75<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
        DV_CUDA(e, cudaFuncSetAttribute(k_attn75<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
        DV_CUDA(e, cudaFuncSetAttribute(k_attn75_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kAttnMmaSmem));
        attr_once.mark(e->device);
    }
    e->launch_begin(use_mma ? "k_attn75_mma" : "k_attn75", layer, 4.0 * 75 * 75 * 192 * B, static_cast<double>(B) * 75 * (576 + 192) * 2);
    if (use_mma) k_attn75_mma<<<B < e->num_sms ? B : e->num_sms, kAttnMmaThreads, 2 * kAttnMmaSmem, e->stream>>>(qkv, ctx, B);
    else if (split) k_attn75<true><<<B, kAttnThreads, kAttnSmem, e->stream>>>(qkv, ctx);
    else k_attn75<false><<<B, kAttnThreads, kAttnSmem, e->stream>>>(qkv, ctx);
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

}  // namespace dv
