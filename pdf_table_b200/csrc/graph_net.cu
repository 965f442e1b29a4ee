// Graph-described network executor: runs the op program that pdf_table_b200/picodet_graph.py lowers PicoDet to
// (LCNet-x1.0 + CSP-PAN + PicoHead; reference picodet/lcnet.py:159-263, picodet/csp_pan.py:233-360,
// picodet/pico_head.py:37-167, 1108-1138).  The CPU mirror is oracle/picodet_net_ref.py.
//
//   OP_STEM  uint8 HWC (or fp32 NCHW) image -> [normalise] -> 3x3 stride-2 conv 3 -> 16 + folded BN + hardswish
//            (OCRPicodetPreProcessor.normalize, picodet/processor_picodet.py:66-70, fused: fp32, numpy's operation order)
//   OP_DW    depthwise k x k (3 / 5), stride 1 / 2, folded BN, activation          CUDA cores, HBM-bound
//   OP_PW    1x1 conv + folded BN + activation                                     conv_igemm_tcgen05 (flat GEMM)
//   OP_SE    x * hardsigmoid(W2 relu(W1 avgpool(x) + b1) + b2)                     (LCNet SEModule :130-156)
//   OP_UP2   nearest 2x up-sampling into a concatenation slice                     (CSPPAN.forward :322-325)
//   OP_ADD   element-wise sum                                                      (the extra top level :343-345)
//   OP_HEAD  1x1 conv 128 -> C + 32 (fp32) -> sigmoid(class scores) [N,HW,C] + raw DFL logits [N,HW,32]
// and, for the PP-OCRv4 recogniser (pdf_table_b200/pp_rec_graph.py: PPLCNetV3-0.95 + SVTR neck + CTC head, SURVEY.md a5;
// CPU mirror oracle/pp_rec_ref.py):
//   OP_DW / OP_PW also take a post-activation affine (LearnableAffineBlock scalars, `w{id}.pa`), OP_DW a (sh, sw) stride,
//            OP_PW a residual tensor (aux) and the Swish activation
//   OP_AVGPOOL  kh x kw average pool, stride = kernel (the eval tail avg_pool2d(x, [3, 2]))
//   OP_UNFOLD3  [N,1,T,C] -> [N,1,T,3C] = [x[t-1] | x[t] | x[t+1]] (zeros outside): the (1,3) convs of the SVTR neck as GEMMs
//   OP_LN       LayerNorm over the channels of each position
//   OP_ATTN     softmax(q k^T) v per (image, head) over the T positions of a text line (global mixer, 8 heads x 15)
//   OP_CTC      Linear -> fp32 logits -> softmax: probabilities [N,T,C] (optional) + per-step arg-max / max probability
// and, for the PP-OCRv4 mobile detector (pdf_table_b200/pp_det_graph.py: PPLCNetV3-0.75 + RSE-FPN + DBHead; CPU
// mirror oracle/pp_det_ref.py):
//   OP_SE with k = 2     the RSELayer shortcut x + x * s (db_fpn.py RSELayer), output may be a concatenation slice
//   OP_UP2 with aux      nearest up-sampling + addend: the top-down sums out_k = in_k + up2(out_k+1)
//   OP_CONV     dense k x k conv, stride 1, + bias + activation                       conv_igemm_tcgen05 (A_PATCH)
//   OP_DECONV2  ConvTranspose 2x2 stride 2 + folded BN + activation                   conv_igemm_tcgen05, pixel-shuffle store
//   OP_DBHEAD   ConvTranspose 2x2 stride 2 C -> 1 + sigmoid -> fp32 probability map [N,1,H,W]
// Every tensor is NHWC fp16; an operand may be a channel slice (coff, c) of a wider buffer, which is how the CSP
// concatenations exist without copies.  The plan (buffers, TMA descriptors) is built once per input shape.
#include <stdlib.h>

#include "engine.h"

namespace dv {

namespace {

enum { OP_STEM = 0, OP_DW, OP_PW, OP_SE, OP_UP2, OP_ADD, OP_HEAD, OP_AVGPOOL, OP_UNFOLD3, OP_LN, OP_ATTN, OP_CTC, OP_CONV, OP_DECONV2, OP_DBHEAD };

__device__ __forceinline__ float act_f(float x, int act) {
    if (act == ACT_HSWISH) return x * fminf(fmaxf(x + 3.f, 0.f), 6.f) * (1.f / 6.f);
    if (act == ACT_RELU) return fmaxf(x, 0.f);
    if (act == ACT_SWISH) return x / (1.f + __expf(-x));
    return x;
}

// Element access for the two activation storages of a graph program: fp16 (default) and fp32 (the fp32x mode: "precision" blob
// entry, see GraphNet::precise).  8 consecutive channels per vector access.
__device__ __forceinline__ void ld8(const __half* p, float (&v)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h[i]);
        v[2 * i] = f.x, v[2 * i + 1] = f.y;
    }
}
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
}
__device__ __forceinline__ void st8(__half* p, const float (&v)[8]) {
    uint4 o;
    __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) ho[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = o;
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ float ld1(const __half* p) { return __half2float(*p); }
__device__ __forceinline__ float ld1(const float* p) { return *p; }
__device__ __forceinline__ void st1(__half* p, float v) { *p = __float2half_rn(v); }
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }

// image -> 16 channels at half resolution.  One thread per output pixel; weights [27][16] (tap-major: r, s, c) in smem.
template <typename T>
__global__ void __launch_bounds__(256)
k_stem3x3s2(const uint8_t* __restrict__ u8, const float* __restrict__ f32, int N, int H, int W, int Ho, int Wo, float3 mean,
            float3 stdv, float scale, int flip, const float* __restrict__ w, const float* __restrict__ bias, int act,
            T* __restrict__ out, const int32_t* __restrict__ widths) {
    __shared__ float sw[27 * 16];
    __shared__ float sb[16];
    __shared__ float lut[3][256];  // u8 input: the normalised value of every (channel, byte), evaluated ONCE per CTA with the
                                   // reference's own operation sequence (a tap otherwise costs two IEEE divisions per channel)
    for (int i = threadIdx.x; i < 27 * 16; i += blockDim.x) sw[i] = w[i];
    if (threadIdx.x < 16) sb[threadIdx.x] = bias[threadIdx.x];
    if (u8) {
        for (int i = threadIdx.x; i < 768; i += blockDim.x) {
            const int ci = i >> 8;
            float c = static_cast<float>(i & 255);
            // PP rec (resize_norm_img): x / 255 is a DIVISION there, `scale` carries the divisor; no FMA contraction anywhere
            c = (flip & 2) ? __fdiv_rn(c, scale) : __fmul_rn(c, scale);
            lut[ci][i & 255] = __fdiv_rn(__fsub_rn(c, ci == 0 ? mean.x : ci == 1 ? mean.y : mean.z), ci == 0 ? stdv.x : ci == 1 ? stdv.y : stdv.z);
        }
    }
    __syncthreads();
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= static_cast<long long>(N) * Ho * Wo) return;
    unsigned t = static_cast<unsigned>(idx);  // 32-bit index decode (a launch covers < 2^32 pixels)
    const int ox = static_cast<int>(t % Wo);
    t /= Wo;
    const int oy = static_cast<int>(t % Ho), n = static_cast<int>(t / Ho);
    float acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = sb[c];
    for (int r = 0; r < 3; ++r) {
        const int iy = 2 * oy - 1 + r;
        if (iy < 0 || iy >= H) continue;
        for (int s = 0; s < 3; ++s) {
            const int ix = 2 * ox - 1 + s;
            if (ix < 0 || ix >= W) continue;
            float v[3];
            if (u8 && widths != nullptr && ix >= widths[n]) {
                v[0] = v[1] = v[2] = 0.f;  // PP rec: zero padding AFTER the normalisation (resize_norm_img)
            } else if (u8) {
                const uint8_t* ip = u8 + ((static_cast<long long>(n) * H + iy) * W + ix) * 3;
                const int b0 = ip[0], b1 = ip[1], b2 = ip[2];
                v[0] = lut[0][(flip & 1) ? b2 : b0];  // flip & 1: the channel-flipped image (processor_ocr_db_pp.py:124)
                v[1] = lut[1][b1];
                v[2] = lut[2][(flip & 1) ? b0 : b2];
            } else {
                const long long plane = static_cast<long long>(H) * W;
                const float* ip = f32 + static_cast<long long>(n) * 3 * plane + static_cast<long long>(iy) * W + ix;
                v[0] = ip[0], v[1] = ip[plane], v[2] = ip[2 * plane];
            }
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
                const float* wp = sw + ((r * 3 + s) * 3 + ci) * 16;
#pragma unroll
                for (int c = 0; c < 16; ++c) acc[c] = fmaf(v[ci], wp[c], acc[c]);
            }
        }
    }
    float lo8[8], hi8[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) lo8[c] = act_f(acc[c], act), hi8[c] = act_f(acc[8 + c], act);
    st8(out + idx * 16, lo8);
    st8(out + idx * 16 + 8, hi8);
}

// depthwise k x k, pad (k-1)/2, stride s; w fp32 [k*k][C] (BN scale folded), b fp32 [C]; one thread = (pixel, 8 channels)
template <typename T>
__global__ void __launch_bounds__(256)
k_dwconv(const T* __restrict__ in, int N, int H, int W, int C, int ldi, int k, int sh, int sw, int Ho, int Wo,
         const float* __restrict__ w, const float* __restrict__ b, int act, float ps, float pb, T* __restrict__ out, int ldo) {
    const int cv = C >> 3;
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= static_cast<long long>(N) * Ho * Wo * cv) return;
    unsigned t = static_cast<unsigned>(idx);  // 32-bit index decode (every launch covers < 2^32 work items): a 64-bit division is ~100 SASS instructions
    const int c8 = static_cast<int>(t % cv);
    t /= cv;
    const int ox = static_cast<int>(t % Wo);
    t /= Wo;
    const int oy = static_cast<int>(t % Ho), n = static_cast<int>(t / Ho);
    float acc[8];
    {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + c8 * 8)), b1 = __ldg(reinterpret_cast<const float4*>(b + c8 * 8 + 4));
        acc[0] = b0.x, acc[1] = b0.y, acc[2] = b0.z, acc[3] = b0.w, acc[4] = b1.x, acc[5] = b1.y, acc[6] = b1.z, acc[7] = b1.w;
    }
    const int pad = (k - 1) >> 1;
    for (int r = 0; r < k; ++r) {
        const int iy = oy * sh - pad + r;
        if (iy < 0 || iy >= H) continue;
        for (int s = 0; s < k; ++s) {
            const int ix = ox * sw - pad + s;
            if (ix < 0 || ix >= W) continue;
            float v[8];
            ld8(in + ((static_cast<long long>(n) * H + iy) * W + ix) * ldi + c8 * 8, v);
            const float4* wp = reinterpret_cast<const float4*>(w + static_cast<long long>(r * k + s) * C + c8 * 8);
            const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
            acc[0] = fmaf(v[0], w0.x, acc[0]);
            acc[1] = fmaf(v[1], w0.y, acc[1]);
            acc[2] = fmaf(v[2], w0.z, acc[2]);
            acc[3] = fmaf(v[3], w0.w, acc[3]);
            acc[4] = fmaf(v[4], w1.x, acc[4]);
            acc[5] = fmaf(v[5], w1.y, acc[5]);
            acc[6] = fmaf(v[6], w1.z, acc[6]);
            acc[7] = fmaf(v[7], w1.w, acc[7]);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(act_f(acc[i], act), ps, pb);
    st8(out + ((static_cast<long long>(n) * Ho + oy) * Wo + ox) * ldo + c8 * 8, acc);
}

__device__ __forceinline__ uint64_t pk2(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// Depthwise k x k, stride 1 along W (any stride along H): one thread = 8 channels x P consecutive output pixels of a row.
// The K + P - 1 input pixels a filter row touches are loaded ONCE into registers and slid across the P outputs, and every
// filter tap's weights are loaded once per thread instead of once per output: 3.0x (k = 3) / 3.3x (k = 5) fewer load
// instructions per output than k_dwconv, which was LSU-bound at 0.7 TB/s on the PP-OCRv4 backbone (profiles/r3h_bench.json).
template <int K, int P>
__global__ void __launch_bounds__(128)
k_dwconv_row(const __half* __restrict__ in, int N, int H, int W, int C, int ldi, int sh, int Ho, int Wo,
             const float* __restrict__ w, const float* __restrict__ b, int act, float ps, float pb, __half* __restrict__ out, int ldo) {
    const int cv = C >> 3;
    const int wq = (Wo + P - 1) / P;
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= static_cast<long long>(N) * Ho * wq * cv) return;
    const int c8 = static_cast<int>(idx % cv);  // 64-bit decode kept here: the 32-bit form made ptxas schedule this kernel 15 % slower (profiles/r5i)
    long long t = idx / cv;
    const int ox0 = static_cast<int>(t % wq) * P;
    t /= wq;
    const int oy = static_cast<int>(t % Ho), n = static_cast<int>(t / Ho);
    constexpr int PAD = (K - 1) / 2;
    // accumulators, pixels and weights live as packed fp32 pairs: Blackwell's fma.rn.f32x2 issues two IEEE FMAs per slot (same
    // bits as two fmaf), and this kernel is issue-bound (ncu profiles/r3m_dwconv_ncu.txt: 62 % issue slots busy, L1 73 %, DRAM 16 %)
    uint64_t acc[P][4];
    {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + c8 * 8)), b1 = __ldg(reinterpret_cast<const float4*>(b + c8 * 8 + 4));
#pragma unroll
        for (int p = 0; p < P; ++p) {
            acc[p][0] = pk2(b0.x, b0.y), acc[p][1] = pk2(b0.z, b0.w), acc[p][2] = pk2(b1.x, b1.y), acc[p][3] = pk2(b1.z, b1.w);
        }
    }
#pragma unroll
    for (int r = 0; r < K; ++r) {
        const int iy = oy * sh - PAD + r;
        if (iy < 0 || iy >= H) continue;
        const __half* row = in + (static_cast<long long>(n) * H + iy) * W * ldi + c8 * 8;
        // the K + P - 1 pixels of this filter row, converted to fp32 ONCE (each feeds up to K outputs)
        uint64_t fx[K + P - 1][4];
#pragma unroll
        for (int j = 0; j < K + P - 1; ++j) {
            const int ix = ox0 - PAD + j;
            const uint4 u = (ix >= 0 && ix < W) ? __ldg(reinterpret_cast<const uint4*>(row + static_cast<long long>(ix) * ldi)) : make_uint4(0u, 0u, 0u, 0u);
            const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 v = __half22float2(h[i]);
                fx[j][i] = pk2(v.x, v.y);
            }
        }
#pragma unroll
        for (int s = 0; s < K; ++s) {
            const float4* wp = reinterpret_cast<const float4*>(w + static_cast<long long>(r * K + s) * C + c8 * 8);
            const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
            const uint64_t wv[4] = {pk2(w0.x, w0.y), pk2(w0.z, w0.w), pk2(w1.x, w1.y), pk2(w1.z, w1.w)};
#pragma unroll
            for (int p = 0; p < P; ++p)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[p][i] = fma2(fx[p + s][i], wv[i], acc[p][i]);
        }
    }
#pragma unroll
    for (int p = 0; p < P; ++p) {
        if (ox0 + p >= Wo) break;
        uint4 o;
        __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float a0, a1;
            upk2(acc[p][i], a0, a1);
            ho[i] = __floats2half2_rn(fmaf(act_f(a0, act), ps, pb), fmaf(act_f(a1, act), ps, pb));
        }
        *reinterpret_cast<uint4*>(out + ((static_cast<long long>(n) * Ho + oy) * Wo + ox0 + p) * ldo + c8 * 8) = o;
    }
}

// Depthwise k x k, stride 1 along W, stride SH along H: one thread = 2 channels x P consecutive output pixels x R output rows,
// ALL k * k filter taps of its channel pair held in registers.  ncu on k_dwconv_row (profiles/r3m_dwconv_ncu.txt) showed the
// L1 data path at 73 % with DRAM at 16 %: per output it re-loaded every tap's weights (32 B) and K input pixels per filter row.
// Here a warp's 32 lanes read 32 consecutive channel pairs of one pixel (one 128-byte line per load instruction), weights cost
// no loads inside the loop, an input row is loaded once for the R output rows it feeds and each pixel feeds up to K outputs:
// ~0.3 B of L1 traffic per FMA instead of ~1.1.
template <int K, int P, int R, int SH>
__global__ void __launch_bounds__(128)
k_dwconv_c2(const __half* __restrict__ in, int N, int H, int W, int C, int ldi, int Ho, int Wo, const float* __restrict__ w,
            const float* __restrict__ b, int act, float ps, float pb, __half* __restrict__ out, int ldo) {
    const int cp = C >> 1;
    const int wq = (Wo + P - 1) / P, hq = (Ho + R - 1) / R;
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= static_cast<long long>(N) * hq * wq * cp) return;
    unsigned t = static_cast<unsigned>(idx);  // 32-bit index decode (every launch covers < 2^32 work items): a 64-bit division is ~100 SASS instructions
    const int c2 = static_cast<int>(t % cp);
    t /= cp;
    const int ox0 = static_cast<int>(t % wq) * P;
    t /= wq;
    const int oy0 = static_cast<int>(t % hq) * R, n = static_cast<int>(t / hq);
    constexpr int PAD = (K - 1) / 2;
    uint64_t wv[K * K];
#pragma unroll
    for (int i = 0; i < K * K; ++i) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(w + static_cast<long long>(i) * C + c2 * 2));
        wv[i] = pk2(v.x, v.y);
    }
    uint64_t acc[R][P];
    {
        const float2 bv = __ldg(reinterpret_cast<const float2*>(b + c2 * 2));
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int p = 0; p < P; ++p) acc[r][p] = pk2(bv.x, bv.y);
    }
    constexpr int ROWS = (R - 1) * SH + K;  // input rows feeding the R output rows
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
        const int iy = oy0 * SH - PAD + j;
        if (iy < 0 || iy >= H) continue;
        const __half* row = in + (static_cast<long long>(n) * H + iy) * W * ldi + c2 * 2;
        uint64_t fx[K + P - 1];
#pragma unroll
        for (int q = 0; q < K + P - 1; ++q) {
            const int ix = ox0 - PAD + q;
            float2 v = make_float2(0.f, 0.f);
            if (ix >= 0 && ix < W) v = __half22float2(__ldg(reinterpret_cast<const __half2*>(row + static_cast<long long>(ix) * ldi)));
            fx[q] = pk2(v.x, v.y);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int fr = j - r * SH;  // filter row of output row r that reads input row j (compile-time after unrolling)
            if (fr < 0 || fr >= K) continue;
#pragma unroll
            for (int s2 = 0; s2 < K; ++s2)
#pragma unroll
                for (int p = 0; p < P; ++p) acc[r][p] = fma2(fx[p + s2], wv[fr * K + s2], acc[r][p]);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (oy0 + r >= Ho) break;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            if (ox0 + p >= Wo) break;
            float a0, a1;
            upk2(acc[r][p], a0, a1);
            *reinterpret_cast<__half2*>(out + ((static_cast<long long>(n) * Ho + oy0 + r) * Wo + ox0 + p) * ldo + c2 * 2) =
                __floats2half2_rn(fmaf(act_f(a0, act), ps, pb), fmaf(act_f(a1, act), ps, pb));
        }
    }
}

// kh x kw average pool with stride = kernel (floor output size), 8 channels per thread
template <typename T>
__global__ void __launch_bounds__(256)
k_avgpool(const T* __restrict__ in, int N, int H, int W, int C, int ldi, int kh, int kw, int Ho, int Wo, T* __restrict__ out, int ldo) {
    const int cv = C >> 3;
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= static_cast<long long>(N) * Ho * Wo * cv) return;
    unsigned t = static_cast<unsigned>(idx);  // 32-bit index decode (every launch covers < 2^32 work items): a 64-bit division is ~100 SASS instructions
    const int c8 = static_cast<int>(t % cv);
    t /= cv;
    const int ox = static_cast<int>(t % Wo);
    t /= Wo;
    const int oy = static_cast<int>(t % Ho), n = static_cast<int>(t / Ho);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < kh; ++r)
        for (int s = 0; s < kw; ++s) {
            float v[8];
            ld8(in + ((static_cast<long long>(n) * H + oy * kh + r) * W + ox * kw + s) * ldi + c8 * 8, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += v[i];
        }
    const float inv = 1.f / static_cast<float>(kh * kw);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] *= inv;
    st8(out + ((static_cast<long long>(n) * Ho + oy) * Wo + ox) * ldo + c8 * 8, acc);
}

// [N, T, C] (row stride ldi) -> [N, T, 3C]: columns [x[t-1] | x[t] | x[t+1]], zeros outside the line
template <typename T_>
__global__ void __launch_bounds__(256)
k_unfold3(const T_* __restrict__ in, int N, int T, int C, int ldi, T_* __restrict__ out) {
    const int cv = C >> 3;
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= static_cast<long long>(N) * T * 3 * cv) return;
    unsigned r = static_cast<unsigned>(idx);  // 32-bit index decode
    const int c8 = static_cast<int>(r % cv);
    r /= cv;
    const int tap = static_cast<int>(r % 3);
    r /= 3;
    const int t = static_cast<int>(r % T);
    const long long n = r / T;
    const int ts = t + tap - 1;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (ts >= 0 && ts < T) ld8(in + (n * T + ts) * ldi + c8 * 8, v);
    st8(out + ((n * T + t) * 3 + tap) * C + c8 * 8, v);  // fp16 -> fp32 -> fp16 is exact: a plain copy either way
}

// LayerNorm over the C channels of each row (one warp per row, any C % 8 == 0 up to 1024), fp16 in / out
template <typename T>
__global__ void __launch_bounds__(256)
k_ln_c(const T* __restrict__ in, long long rows, int C, int ldi, const float* __restrict__ g, const float* __restrict__ b, float eps,
       T* __restrict__ out, int ldo) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = static_cast<long long>(blockIdx.x) * 8 + warp;
    if (row >= rows) return;
    const T* ip = in + row * ldi;
    float v[32];
    float s = 0.f;
    const int per = (C + 31) / 32;
    for (int j = 0; j < per; ++j) {
        const int c = lane + 32 * j;
        v[j] = c < C ? ld1(ip + c) : 0.f;
        s += v[j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / static_cast<float>(C);
    float q = 0.f;
    for (int j = 0; j < per; ++j) {
        const int c = lane + 32 * j;
        const float d = c < C ? v[j] - mean : 0.f;
        v[j] = d;
        q += d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / static_cast<float>(C) + eps);
    T* op = out + row * ldo;
    for (int j = 0; j < per; ++j) {
        const int c = lane + 32 * j;
        if (c < C) st1(op + c, v[j] * rstd * __ldg(g + c) + __ldg(b + c));
    }
}

// Global-mixer attention of the SVTR neck: qkv [N, T, 3D] (q | k | v, each head-major heads x hd; the 1/sqrt(hd) scale is
// folded into the packed q weights) -> ctx [N, T, D].  One CTA per image; K and V of all heads staged in shared memory as
// fp32, one thread per (head, query) with an online softmax (no score buffer, any T that fits shared memory).
// HD > 0: the head width is a compile-time constant (15 for the PP-OCRv4 neck), so q / o live in registers and the inner loops
// unroll; with the run-time width (HD = 0) the two arrays were indexed dynamically and went to local memory (0.33 ms per launch
// against 0.05).
template <typename T_, int HD>
__global__ void __launch_bounds__(256)
k_attn_small(const T_* __restrict__ qkv, int T, int D, int heads, T_* __restrict__ ctx) {
    extern __shared__ float sm[];  // K [T][D] | V [T][D]
    float* sK = sm;
    float* sV = sm + static_cast<size_t>(T) * D;
    const int n = blockIdx.x;
    const int hd = HD > 0 ? HD : D / heads;
    const T_* base = qkv + static_cast<long long>(n) * T * 3 * D;
    for (int i = threadIdx.x; i < T * D; i += blockDim.x) {
        const int t = i / D, c = i - t * D;
        sK[i] = ld1(base + static_cast<long long>(t) * 3 * D + D + c);
        sV[i] = ld1(base + static_cast<long long>(t) * 3 * D + 2 * D + c);
    }
    __syncthreads();
    constexpr int kMaxHd = HD > 0 ? HD : 16;
    for (int w = threadIdx.x; w < heads * T; w += blockDim.x) {
        const int h = w / T, qi = w - h * T;
        float q[kMaxHd], o[kMaxHd];
#pragma unroll
        for (int d = 0; d < kMaxHd; ++d) {
            q[d] = d < hd ? ld1(base + static_cast<long long>(qi) * 3 * D + h * hd + d) : 0.f;
            o[d] = 0.f;
        }
        float m = -INFINITY, l = 0.f;
        for (int j = 0; j < T; ++j) {
            const float* kp = sK + j * D + h * hd;
            float sc = 0.f;
#pragma unroll
            for (int d = 0; d < kMaxHd; ++d)
                if (d < hd) sc = fmaf(q[d], kp[d], sc);
            const float mn = fmaxf(m, sc);
            const float corr = __expf(m - mn), pj = __expf(sc - mn);
            l = l * corr + pj;
            const float* vp = sV + j * D + h * hd;
#pragma unroll
            for (int d = 0; d < kMaxHd; ++d)
                if (d < hd) o[d] = fmaf(o[d], corr, pj * vp[d]);
            m = mn;
        }
        const float inv = 1.f / l;
        T_* op = ctx + (static_cast<long long>(n) * T + qi) * D + h * hd;
#pragma unroll
        for (int d = 0; d < kMaxHd; ++d)
            if (d < hd) st1(op + d, o[d] * inv);
    }
}

// fp32 logits [M, ld] -> softmax over the first C columns: probs [M, C] (optional), arg-max id and max probability per row
__global__ void __launch_bounds__(256)
k_softmax_rows(const float* __restrict__ logits, long long M, int ld, int C, float* __restrict__ probs, int32_t* __restrict__ ids,
               float* __restrict__ maxp, float* __restrict__ raw) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = static_cast<long long>(blockIdx.x) * 8 + warp;
    if (row >= M) return;
    const float* ip = logits + row * ld;
    float mx = -INFINITY;
    int arg = 0;
    for (int c = lane; c < C; c += 32) {
        const float v = ip[c];
        if (v > mx) {
            mx = v;
            arg = c;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, mx, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ov > mx || (ov == mx && oa < arg)) {
            mx = ov;
            arg = oa;
        }
    }
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(ip[c] - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float inv = 1.f / s;
    if (probs != nullptr)
        for (int c = lane; c < C; c += 32) probs[row * C + c] = expf(ip[c] - mx) * inv;
    if (raw != nullptr)
        for (int c = lane; c < C; c += 32) raw[row * C + c] = ip[c];
    if (lane == 0) {
        if (ids != nullptr) ids[row] = arg;
        if (maxp != nullptr) maxp[row] = inv;  // exp(0) / sum
    }
}

// The two SE matrices, one warp per output row: the lanes stride over the row of the weight matrix (coalesced; a thread per
// row would read w1 / w2 with a stride of a whole row) and the partial sums meet in a shuffle tree.
__device__ __forceinline__ void se_fc(const float* avg, float* hid, int C, int R, const float* __restrict__ w1, const float* __restrict__ b1,
                                      const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ scale_out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int r = warp; r < R; r += nwarps) {
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s = fmaf(__ldg(w1 + r * C + c), avg[c], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) hid[r] = fmaxf(s + b1[r], 0.f);
    }
    __syncthreads();
    for (int c = warp; c < C; c += nwarps) {
        float s = 0.f;
        for (int r = lane; r < R; r += 32) s = fmaf(__ldg(w2 + c * R + r), hid[r], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) scale_out[c] = fminf(fmaxf((s + b2[c]) * (1.f / 6.f) + 0.5f, 0.f), 1.f);  // F.hardsigmoid
    }
}

// SE squeeze: one CTA per image.  avg[c] -> hidden = relu(W1 avg + b1) -> scale[c] = hardsigmoid(W2 hidden + b2)
template <typename T>
__global__ void __launch_bounds__(512)
k_se_scale(const T* __restrict__ in, int HW, int C, const float* __restrict__ w1, const float* __restrict__ b1,
           const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ scale) {
    extern __shared__ float sm[];  // avg[C] | hidden[C/4]
    float* avg = sm;
    float* hid = sm + C;
    const int n = blockIdx.x, R = C >> 2;
    const T* base = in + static_cast<long long>(n) * HW * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int p = 0; p < HW; ++p) s += ld1(base + static_cast<long long>(p) * C + c);
        avg[c] = s / static_cast<float>(HW);
    }
    __syncthreads();
    se_fc(avg, hid, C, R, w1, b1, w2, b2, scale + n * C);
}

// SE squeeze over a LARGE map (the detector's RSE layers pool 240 x 240 x 96 per page): partial channel sums of kSePoolRows
// pixels per CTA, grid (chunks, N), summed in a fixed order by k_se_scale_p -- deterministic, and 300x the single-CTA loop's
// 45 GB/s (profiles/r4g_layers_ppdet.txt: 13.7 of 20.1 ms per 32 pages before).
constexpr int kSePoolRows = 512;
constexpr int kSePoolMin = 256;  // maps with fewer pixels keep the one-kernel squeeze
template <typename T>
__global__ void __launch_bounds__(256)
k_se_pool(const T* __restrict__ in, int HW, int C, float* __restrict__ partial) {
    extern __shared__ float sm[];  // [slots][C]
    const int cv = C >> 3, slots = 256 / cv;
    const int n = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
    const int c8 = threadIdx.x % cv, slot = threadIdx.x / cv;
    const int p0 = chunk * kSePoolRows, p1 = min(p0 + kSePoolRows, HW);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (slot < slots) {
        const T* base = in + (static_cast<long long>(n) * HW) * C + c8 * 8;
        for (int p = p0 + slot; p < p1; p += slots) {
            float v[8];
            ld8(base + static_cast<long long>(p) * C, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += v[i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) sm[slot * C + c8 * 8 + i] = acc[i];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < slots; ++k) s += sm[k * C + c];
        partial[(static_cast<long long>(n) * nchunks + chunk) * C + c] = s;
    }
}

__global__ void __launch_bounds__(512)
k_se_scale_p(const float* __restrict__ partial, int nchunks, int HW, int C, const float* __restrict__ w1, const float* __restrict__ b1,
             const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ scale) {
    extern __shared__ float sm[];  // avg[C] | hidden[C/4]
    float* avg = sm;
    float* hid = sm + C;
    const int n = blockIdx.x, R = C >> 2;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < nchunks; ++k) s += partial[(static_cast<long long>(n) * nchunks + k) * C + c];
        avg[c] = s / static_cast<float>(HW);
    }
    __syncthreads();
    se_fc(avg, hid, C, R, w1, b1, w2, b2, scale + n * C);
}

template <typename T>
__global__ void __launch_bounds__(256)
k_se_apply(const T* __restrict__ in, long long total8, int HW, int C, const float* __restrict__ scale, float shortcut,
           T* __restrict__ out, int ldo) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= total8) return;
    const int cv = C >> 3;
    const unsigned i32 = static_cast<unsigned>(idx);  // 32-bit index decode (a launch covers < 2^32 work items)
    const int c8 = static_cast<int>(i32 % cv);
    const long long pix = i32 / cv;
    const int n = static_cast<int>((i32 / cv) / HW);
    float v[8];
    ld8(in + idx * 8, v);
    const float* sp = scale + n * C + c8 * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] *= sp[i] + shortcut;  // shortcut = 1: x + x * s
    st8(out + pix * ldo + c8 * 8, v);
}

template <typename T>
__global__ void __launch_bounds__(256)
k_up2(const T* __restrict__ in, int N, int h, int w, int C, int ldi, int Ho, int Wo, const T* __restrict__ add, T* __restrict__ out, int ldo) {
    const int cv = C >> 3;
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= static_cast<long long>(N) * Ho * Wo * cv) return;
    unsigned t = static_cast<unsigned>(idx);  // 32-bit index decode (every launch covers < 2^32 work items): a 64-bit division is ~100 SASS instructions
    const int c8 = static_cast<int>(t % cv);
    t /= cv;
    const int ox = static_cast<int>(t % Wo);
    t /= Wo;
    const int oy = static_cast<int>(t % Ho), n = static_cast<int>(t / Ho);
    // F.interpolate(mode="nearest") to an explicit size: src = floor(dst * in / out)
    const int iy = min(static_cast<int>(static_cast<long long>(oy) * h / Ho), h - 1), ix = min(static_cast<int>(static_cast<long long>(ox) * w / Wo), w - 1);
    float v[8];
    ld8(in + ((static_cast<long long>(n) * h + iy) * w + ix) * ldi + c8 * 8, v);
    if (add != nullptr) {  // dense addend of the output's shape (the top-down sum of an FPN)
        float a[8];
        ld8(add + ((static_cast<long long>(n) * Ho + oy) * Wo + ox) * C + c8 * 8, a);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += a[i];
    }
    st8(out + ((static_cast<long long>(n) * Ho + oy) * Wo + ox) * ldo + c8 * 8, v);
}

// ConvTranspose2d(C -> 1, k 2, s 2) + sigmoid: fp16 [N,H,W,C] -> fp32 probability map [N,1,2H,2W]; w fp32 [C][dy*2+dx].
// One thread per input pixel (DBHead.binarize conv3, det_db_head.py).
template <typename T>
__global__ void __launch_bounds__(128)
k_dbhead(const T* __restrict__ in, long long npix, int H, int W, int C, const float* __restrict__ w, const float* __restrict__ bias,
         float* __restrict__ out) {
    __shared__ float sw[64 * 4];
    for (int i = threadIdx.x; i < C * 4; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= npix) return;
    const int x = static_cast<int>(idx % W), y = static_cast<int>((idx / W) % H);
    const long long n = idx / (static_cast<long long>(W) * H);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int j = 0; j < (C >> 3); ++j) {
        float f[8];
        ld8(in + idx * C + j * 8, f);
#pragma unroll
        for (int e2 = 0; e2 < 8; ++e2) {
            const int c = j * 8 + e2;
            a0 = fmaf(f[e2], sw[c * 4 + 0], a0);
            a1 = fmaf(f[e2], sw[c * 4 + 1], a1);
            a2 = fmaf(f[e2], sw[c * 4 + 2], a2);
            a3 = fmaf(f[e2], sw[c * 4 + 3], a3);
        }
    }
    const float b = __ldg(bias);
    const int W2 = 2 * W;
    float* op = out + (n * 2 * H + 2 * y) * W2 + 2 * x;
    *reinterpret_cast<float2*>(op) = make_float2(1.f / (1.f + expf(-(a0 + b))), 1.f / (1.f + expf(-(a1 + b))));
    *reinterpret_cast<float2*>(op + W2) = make_float2(1.f / (1.f + expf(-(a2 + b))), 1.f / (1.f + expf(-(a3 + b))));
}

__global__ void __launch_bounds__(256) k_add(const __half* __restrict__ a, const __half* __restrict__ b, long long total8, __half* __restrict__ out) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= total8) return;
    const uint4 ua = __ldg(reinterpret_cast<const uint4*>(a) + idx), ub = __ldg(reinterpret_cast<const uint4*>(b) + idx);
    const __half2 *ha = reinterpret_cast<const __half2*>(&ua), *hb = reinterpret_cast<const __half2*>(&ub);
    uint4 o;
    __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 x = __half22float2(ha[i]), y = __half22float2(hb[i]);
        ho[i] = __floats2half2_rn(x.x + y.x, x.y + y.y);
    }
    reinterpret_cast<uint4*>(out)[idx] = o;
}

// raw fp32 [M, ld] head output -> scores [M, C] = sigmoid(raw[:, :C]), dfl [M, R] = raw[:, C:C+R]
__global__ void __launch_bounds__(256)
k_head_split(const float* __restrict__ raw, long long M, int ld, int C, int R, float* __restrict__ scores, float* __restrict__ dfl) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= M * (C + R)) return;
    const long long m = idx / (C + R);
    const int j = static_cast<int>(idx % (C + R));
    const float v = raw[m * ld + j];
    if (j < C) scores[m * C + j] = 1.f / (1.f + expf(-v));
    else dfl[m * R + (j - C)] = v;
}

// fp32x mode: fp32 rows [M, ld] (columns [0, K)) -> the split-fp16 A operand [M, 2 * Kp]: hi = fp16(x) at [0, Kp) (zero beyond K),
// lo = fp16(x - hi) at [Kp, 2 * Kp); with weights packed [W_hi | W_lo | W_hi] conv_igemm_tcgen05 accumulates the three partial
// products in fp32 (plan_linear, ConvSpec::split)
__global__ void __launch_bounds__(256)
k_split_f32(const float* __restrict__ in, long long M, int K, int Kp, int ld, __half* __restrict__ out) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const int kv = Kp >> 3;
    if (idx >= M * kv) return;
    const long long r = idx / kv;
    const int c = static_cast<int>(idx - r * kv) * 8;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c < K) ld8(in + r * ld + c, v);  // K % 8 == 0
    uint4 hi, lo;
    __half2 *hh = reinterpret_cast<__half2*>(&hi), *hl = reinterpret_cast<__half2*>(&lo);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        hh[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        const float2 f = __half22float2(hh[i]);
        hl[i] = __floats2half2_rn(v[2 * i] - f.x, v[2 * i + 1] - f.y);
    }
    *reinterpret_cast<uint4*>(out + r * 2 * Kp + c) = hi;
    *reinterpret_cast<uint4*>(out + r * 2 * Kp + Kp + c) = lo;
}

struct GOp {
    int code, in_t, in_coff, in_c, out_t, out_coff, out_c, k, stride, act, w, aux;
    ConvPlan plan;  // OP_PW / OP_HEAD / OP_CTC / OP_CONV / OP_DECONV2
    const float *f0 = nullptr, *f1 = nullptr, *f2 = nullptr, *f3 = nullptr;
    float ps = 1.f, pb = 0.f;  // post-activation affine (w{id}.pa), identity when absent
    int has_pa = 0;
    float eps = 1e-5f;         // OP_LN (w{id}.eps)
    int kp = 0;                // fp32x: padded K of the split A operand
    int sh() const { return stride < 256 ? stride : (stride & 255); }  // stride = sh | sw << 8 when the two differ
    int sw() const { return stride < 256 ? stride : (stride >> 8); }
};

struct GraphNet : Model {
    Engine* e = nullptr;
    int N = 0, H = 0, W = 0;
    int kind = 0;  // graph.meta[5]: 0 = PicoDet, 1 = PP-OCR recogniser, 2 = PULC classifier, 3 = PP-OCR detector
    // fp32x mode (blob entry "precision", pp_rec_graph.pack_pp_rec(precise=True)): every activation buffer is fp32, the CUDA-core
    // kernels read / write fp32, and every GEMM takes its A operand through k_split_f32 as a split-fp16 pair against weights
    // packed [W_hi | W_lo | W_hi] (three MMAs per product, fp32 accumulation, fp32 output).  Probabilities of the PP-OCRv4
    // recogniser land within 1e-4 of the fp32 oracle (tests/test_gpu_pp_rec.py).
    bool precise = false;
    __half* split_buf = nullptr;
    int num_classes = 0, reg_bins = 32, head_ld = 40;
    std::vector<int> tc, tdh, tdw, tph, tpw;  // channels; size = floor(ceil(H / dh) / ph) x floor(ceil(W / dw) / pw)
    std::vector<GOp> ops;
    std::vector<Tensor> tens;
    std::vector<void*> mem;
    float* se_scale = nullptr;
    float* se_partial = nullptr;  // k_se_pool partial sums [N][chunks][C] (large maps only)
    float* head_raw = nullptr;
    double flops = 0;
    // plans of other input shapes (the recogniser alternates between a full pass and a tail pass, and between padded widths):
    // a shape change parks the current plan here instead of freeing it
    struct Saved {
        int N, H, W;
        std::vector<Tensor> tens;
        std::vector<void*> mem;
        std::vector<GOp> ops;
        float *se_scale, *head_raw, *se_partial;
        __half* split_buf;
        double flops;
    };
    std::vector<Saved> cache;
    double last_call_flops = 0;
    int pass_n = 2048;  // images per pass of the recogniser: bounds the activation workspace (~10 MB per crop); smaller passes are
                        // SLOWER on the B200 (profiles/r3j_pp_rec_pass_sweep.json: the kernels are issue-bound, not HBM-bound)
    ~GraphNet() override {
        for (void* p : mem) cudaFree(p);
        for (Saved& sv : cache)
            for (void* p : sv.mem) cudaFree(p);
    }
    void park() {
        if (N == 0) return;
        cache.push_back(Saved{N, H, W, std::move(tens), std::move(mem), ops, se_scale, head_raw, se_partial, split_buf, flops});
        mem.clear();
        tens.clear();
        N = H = W = 0;
        if (cache.size() > 6) {
            for (void* p : cache.front().mem) cudaFree(p);
            cache.erase(cache.begin());
        }
    }
    bool restore(int n, int h, int w) {
        for (size_t i = 0; i < cache.size(); ++i)
            if (cache[i].N == n && cache[i].H == h && cache[i].W == w) {
                Saved sv = std::move(cache[i]);
                cache.erase(cache.begin() + i);
                N = sv.N, H = sv.H, W = sv.W;
                tens = std::move(sv.tens);
                mem = std::move(sv.mem);
                ops = std::move(sv.ops);
                se_scale = sv.se_scale, head_raw = sv.head_raw, se_partial = sv.se_partial, split_buf = sv.split_buf, flops = sv.flops;
                return true;
            }
        return false;
    }
    int alloc(void** p, size_t bytes) {
        cudaError_t st = cudaMalloc(p, bytes ? bytes : 16);
        if (st != cudaSuccess) return set_err(e, DV_ERR_CUDA, "graph: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(st));
        mem.push_back(*p);
        return 0;
    }
};

const float* wf32(Engine* e, int id, const char* field, size_t elems, int* rc) {
    const std::string name = "w" + std::to_string(id) + "." + field;
    const BlobTensor* t = e->find(name);
    if (!t || t->dtype != 0 || t->nbytes < elems * 4) {
        *rc = set_err(e, DV_ERR_WEIGHTS, "graph: missing / short fp32 tensor '%s'", name.c_str());
        return nullptr;
    }
    return reinterpret_cast<const float*>(t->dptr);
}

int build(Engine* e, GraphNet* m, int N, int H, int W) {
    for (void* p : m->mem) cudaFree(p);
    m->mem.clear();
    m->e = e;
    m->N = N;
    m->H = H;
    m->W = W;
    m->flops = 0;
    const size_t nt = m->tc.size();
    m->tens.assign(nt, Tensor());
    for (size_t i = 1; i < nt; ++i) {  // tensor 0 is the input image
        Tensor& t = m->tens[i];
        t.N = N;
        // pool factor 0 = globally pooled along that axis
        t.H = m->tph[i] == 0 ? 1 : ((H + m->tdh[i] - 1) / m->tdh[i]) / m->tph[i];
        t.W = m->tpw[i] == 0 ? 1 : ((W + m->tdw[i] - 1) / m->tdw[i]) / m->tpw[i];
        if (t.H <= 0 || t.W <= 0) return set_err(e, DV_ERR_ARG, "graph: input %dx%d too small for tensor %zu", H, W, i);
        t.C = m->tc[i];
        void* p = nullptr;
        DV_TRY(m->alloc(&p, t.elems() * (m->precise ? sizeof(float) : sizeof(__half))));
        t.p = reinterpret_cast<__half*>(p);  // fp32x: the buffer holds floats (see F32 below)
    }
    size_t max_head_rows = 0, max_se_partial = 0;
    int max_se_c = 0;
    for (GOp& op : m->ops) {
        int rc = 0;
        const Tensor& in = m->tens[op.in_t];
        const Tensor& out = m->tens[op.out_t];
        switch (op.code) {
            case OP_STEM:
                op.f0 = wf32(e, op.w, "sw", 27 * 16, &rc);
                op.f1 = wf32(e, op.w, "sb", 16, &rc);
                m->flops += 2.0 * N * out.H * out.W * 27 * 16;
                break;
            case OP_DW:
                op.f0 = wf32(e, op.w, "dw", static_cast<size_t>(op.k) * op.k * op.in_c, &rc);
                op.f1 = wf32(e, op.w, "db", op.in_c, &rc);
                m->flops += 2.0 * N * out.H * out.W * op.k * op.k * op.in_c;
                break;
            case OP_LN: {
                op.f0 = wf32(e, op.w, "lnw", op.in_c, &rc);
                op.f1 = wf32(e, op.w, "lnb", op.in_c, &rc);
                const float* ep = wf32(e, op.w, "eps", 1, &rc);
                if (!rc) DV_CUDA(e, cudaMemcpy(&op.eps, ep, 4, cudaMemcpyDeviceToHost));
                if (op.in_c > 1024) rc = set_err(e, DV_ERR_UNSUPPORTED, "graph: LayerNorm over %d channels", op.in_c);
                break;
            }
            case OP_ATTN:
                if (op.k <= 0 || (op.out_c % op.k) || op.out_c / op.k > 16 || op.in_c != 3 * op.out_c || in.C != op.in_c || in.H != 1)
                    rc = set_err(e, DV_ERR_UNSUPPORTED, "graph: attention needs [N,1,T,3D] input, head dim <= 16");
                m->flops += 4.0 * N * in.W * in.W * op.out_c;
                break;
            case OP_UNFOLD3:
                if (in.H != 1 || out.C != 3 * op.in_c) rc = set_err(e, DV_ERR_UNSUPPORTED, "graph: unfold3 needs a [N,1,T,C] line and a dense [.,3C] output");
                break;
            case OP_CTC:
                if (static_cast<size_t>(N) * in.H * in.W > max_head_rows) max_head_rows = static_cast<size_t>(N) * in.H * in.W;
                break;
            case OP_SE:
                op.f0 = wf32(e, op.w, "s1w", static_cast<size_t>(op.in_c) * op.in_c / 4, &rc);
                op.f1 = wf32(e, op.w, "s1b", op.in_c / 4, &rc);
                op.f2 = wf32(e, op.w, "s2w", static_cast<size_t>(op.in_c) * op.in_c / 4, &rc);
                op.f3 = wf32(e, op.w, "s2b", op.in_c, &rc);
                if (op.in_c > max_se_c) max_se_c = op.in_c;
                if (in.C != op.in_c || op.in_coff != 0) rc = set_err(e, DV_ERR_UNSUPPORTED, "graph: SE needs a dense input");
                if (in.H * in.W >= kSePoolMin) {
                    const size_t need = static_cast<size_t>(N) * ((in.H * in.W + kSePoolRows - 1) / kSePoolRows) * op.in_c;
                    if (need > max_se_partial) max_se_partial = need;
                    if (op.in_c > 2048) rc = set_err(e, DV_ERR_UNSUPPORTED, "graph: SE over %d channels", op.in_c);
                }
                break;
            case OP_HEAD:
                if (static_cast<size_t>(N) * in.H * in.W > max_head_rows) max_head_rows = static_cast<size_t>(N) * in.H * in.W;
                break;
            case OP_DBHEAD:
                if (op.in_c > 64 || in.C != op.in_c) rc = set_err(e, DV_ERR_UNSUPPORTED, "graph: DB head over %d channels", op.in_c);
                op.f0 = wf32(e, op.w, "hw", static_cast<size_t>(op.in_c) * 4, &rc);
                op.f1 = wf32(e, op.w, "hb", 1, &rc);
                m->flops += 2.0 * N * in.H * in.W * op.in_c * 4;
                break;
            default: break;
        }
        if (rc) return rc;
        if (op.w >= 0 && (op.code == OP_DW || op.code == OP_PW)) {  // optional post-activation affine
            const BlobTensor* pa = e->find("w" + std::to_string(op.w) + ".pa");
            if (pa && pa->dtype == 0 && pa->nbytes >= 8) {
                float h2[2];
                DV_CUDA(e, cudaMemcpy(h2, pa->dptr, 8, cudaMemcpyDeviceToHost));
                op.ps = h2[0];
                op.pb = h2[1];
                op.has_pa = 1;
            }
        }
    }
    if (max_se_c) {
        void* p = nullptr;
        DV_TRY(m->alloc(&p, static_cast<size_t>(N) * max_se_c * 4));
        m->se_scale = reinterpret_cast<float*>(p);
    }
    if (max_se_partial) {
        void* p = nullptr;
        DV_TRY(m->alloc(&p, max_se_partial * 4));
        m->se_partial = reinterpret_cast<float*>(p);
    }
    {
        void* p = nullptr;
        DV_TRY(m->alloc(&p, max_head_rows * m->head_ld * 4 + 16));
        m->head_raw = reinterpret_cast<float*>(p);
    }
    if (m->precise) {  // one split-operand scratch for all GEMMs of the program (they run one after the other)
        size_t need = 0;
        for (const GOp& op : m->ops) {
            if (op.code != OP_PW && op.code != OP_HEAD && op.code != OP_CTC && op.code != OP_CONV && op.code != OP_DECONV2) continue;
            if (op.code == OP_HEAD) return set_err(e, DV_ERR_UNSUPPORTED, "graph: the fp32x mode does not cover the PicoDet head");
            const BlobTensor* w = e->find("w" + std::to_string(op.w) + ".w");
            const Tensor& in = m->tens[op.in_t];
            const int taps = op.code == OP_CONV ? op.k * op.k : 1;
            if (w && w->ndim == 2) need = std::max(need, static_cast<size_t>(N) * in.H * in.W * 2 * (w->dims[1] / 3 / taps));
        }
        void* p = nullptr;
        DV_TRY(m->alloc(&p, need * sizeof(__half)));
        m->split_buf = reinterpret_cast<__half*>(p);
    }
    for (GOp& op : m->ops) {
        if (op.code != OP_CONV && op.code != OP_DECONV2) continue;
        const Tensor& in = m->tens[op.in_t];
        const Tensor& out = m->tens[op.out_t];
        const std::string wn = "w" + std::to_string(op.w);
        const BlobTensor* w = e->find(wn + ".w");
        const BlobTensor* b = e->find(wn + ".b");
        const bool dec = op.code == OP_DECONV2;
        const int parts = m->precise ? 3 : 1;  // fp32x: [W_hi | W_lo | W_hi] per filter tap
        const int taps = dec ? 1 : op.k * op.k, rows = dec ? 4 * op.out_c : op.out_c;
        if (!w || !b || w->dtype != 1 || b->dtype != 0 || w->ndim != 2 || static_cast<int>(w->dims[0]) != rows || (w->dims[1] % (taps * parts)) ||
            static_cast<int>(w->dims[1]) / (taps * parts) < op.in_c || ((w->dims[1] / (taps * parts)) % 16))
            return set_err(e, DV_ERR_WEIGHTS, "graph: bad conv weights '%s' (want [%d, %d x >= %d])", wn.c_str(), rows, taps, op.in_c);
        const int kpad = static_cast<int>(w->dims[1]) / (taps * parts);
        op.kp = kpad;
        ConvSpec cs;
        cs.KH = cs.KW = dec ? 1 : op.k;
        cs.pad = dec ? 0 : op.k / 2;
        cs.stride = 1;
        cs.Cin = m->precise ? kpad : op.in_c;
        cs.Cin_pad = (m->precise && dec) ? 3 * kpad : kpad;  // the flat (1x1) path counts all three parts, the patch path one
        cs.split = m->precise;
        cs.Cout = rows;
        cs.BK = (cs.Cin_pad % 64 == 0) ? 64 : (cs.Cin_pad % 32 == 0) ? 32 : 16;
        cs.w = reinterpret_cast<const __half*>(w->dptr);
        cs.bias = reinterpret_cast<const float*>(b->dptr);
        EpiSpec es;
        es.act = op.act;
        es.out = out.p;
        es.out_ld = out.C;
        es.out_coff = op.out_coff;
        es.out_f32 = m->precise ? 1 : 0;
        if (dec) {
            es.out_mode = OUT_SHUF2;
            if (out.H != 2 * in.H || out.W != 2 * in.W || op.out_coff != 0 || out.C != op.out_c)
                return set_err(e, DV_ERR_WEIGHTS, "graph: transposed conv '%s' needs a dense output at twice the input size", wn.c_str());
        } else if (out.H != in.H || out.W != in.W) {
            return set_err(e, DV_ERR_WEIGHTS, "graph: conv '%s' changes the map size", wn.c_str());
        }
        Tensor a = in;  // the op's channel slice of the input buffer
        a.p = in.p + op.in_coff;
        a.ld = in.C;
        a.C = op.in_c;
        if (m->precise) {  // the split scratch as an NHWC tensor of [hi(Kp) | lo(Kp)] pixels (k_split_f32 fills it before the launch)
            a.p = m->split_buf;
            a.C = kpad;
            a.ld = 2 * kpad;
            a.lo = kpad;
        }
        DV_TRY(plan_conv(e, a, cs, es, in.H, in.W, &op.plan, wn.c_str()));
        m->mem.push_back(e->owned.back());
        e->owned.pop_back();
        m->flops += op.plan.flops;
    }
    for (GOp& op : m->ops) {
        if (op.code != OP_PW && op.code != OP_HEAD && op.code != OP_CTC) continue;
        const Tensor& in = m->tens[op.in_t];
        const Tensor& out = m->tens[op.out_t];
        const std::string wn = "w" + std::to_string(op.w);
        const BlobTensor* w = e->find(wn + ".w");
        const BlobTensor* b = e->find(wn + ".b");
        // the packed K axis is padded to a multiple of 16 (weights.cin_pad_of) or, for the wide layers, of 64 (pp_rec_graph.
        // _gemm_weight: 64-wide k-blocks = 128-byte TMA rows and half the barrier round trips); the A operand keeps its own
        // width, the TMA unit zero-fills the columns beyond it
        const int kw_total = w ? static_cast<int>(w->dims[1]) : 0;
        const int kpad = m->precise ? kw_total / 3 : kw_total;  // fp32x: [W_hi | W_lo | W_hi], each Kp wide
        // OP_PW with k > 1: k consecutive pixels per GEMM row against diag(w, ..., w) (pp_rec_graph.pw_pack_factor) -- a pure
        // re-interpretation of the dense NHWC buffers as [M / k, k * C]
        const int pack = (op.code == OP_PW && op.k > 1) ? op.k : 1;
        if (pack > 1 && (m->precise || op.in_coff || op.out_coff || in.C != op.in_c || out.C != op.out_c || op.aux >= 0 ||
                         (static_cast<long long>(N) * in.H * in.W) % pack))
            return set_err(e, DV_ERR_UNSUPPORTED, "graph: packed 1x1 '%s' needs dense whole tensors, no residual and a pixel count divisible by %d", wn.c_str(), pack);
        const int kin = op.in_c * pack, nout = op.out_c * pack;
        if (!w || !b || w->dtype != 1 || b->dtype != 0 || w->ndim != 2 || static_cast<int>(w->dims[0]) != nout || kpad < kin ||
            (m->precise ? kw_total != 3 * kpad : false) || kpad - kin >= 64 || (kpad % 16))
            return set_err(e, DV_ERR_WEIGHTS, "graph: bad 1x1 weights '%s' (want [%d,%d])", wn.c_str(), nout, kin);
        ConvSpec cs;
        cs.KH = cs.KW = 1;
        cs.Cin = kin;
        cs.Cin_pad = kw_total;
        cs.Cout = nout;
        cs.BK = (kpad % 64 == 0) ? 64 : (kpad % 32 == 0) ? 32 : 16;
        cs.w = reinterpret_cast<const __half*>(w->dptr);
        cs.bias = reinterpret_cast<const float*>(b->dptr);
        cs.flat = true;
        cs.split = m->precise;
        op.kp = kpad;
        EpiSpec es;
        es.act = op.act;
        if (op.code == OP_HEAD || op.code == OP_CTC) {
            es.out = m->head_raw;
            es.out_ld = m->head_ld;
            es.out_f32 = 1;
        } else {
            es.out = out.p;
            es.out_ld = out.C * pack;
            es.out_coff = op.out_coff;
            es.out_f32 = m->precise ? 1 : 0;
            es.post_affine = op.has_pa;
            es.post_scale = op.ps;
            es.post_bias = op.pb;
            if (op.aux >= 0) {  // residual: a dense tensor of the output's shape
                const Tensor& r = m->tens[op.aux];
                if (r.C != op.out_c || r.H != out.H || r.W != out.W) return set_err(e, DV_ERR_WEIGHTS, "graph: residual shape mismatch at '%s'", wn.c_str());
                es.res = r.p;
                es.res_mode = RES_SAME;
                es.res_ld = r.C;
                es.res_f32 = m->precise ? 1 : 0;
                if (m->precise && op.act != ACT_NONE) return set_err(e, DV_ERR_UNSUPPORTED, "graph: fp32x residual with an activation at '%s'", wn.c_str());
            }
        }
        if (m->precise) {
            const int M = N * in.H * in.W;
            DV_TRY(plan_linear(e, m->split_buf, M, kpad, cs, es, &op.plan, wn.c_str(), 2 * kpad, kpad));
            m->mem.push_back(e->owned.back());
            e->owned.pop_back();
            m->flops += op.plan.flops;
            continue;
        }
        const int M = N * in.H * in.W / pack;
        DV_TRY(plan_linear(e, in.p + op.in_coff, M, kin, cs, es, &op.plan, wn.c_str(), in.C * pack));
        m->mem.push_back(e->owned.back());
        e->owned.pop_back();
        op.plan.flops /= pack;  // the off-diagonal zero blocks are not algorithmic work
        m->flops += op.plan.flops;
    }
    return 0;
}

static inline int grid_for(long long n, int block) { return static_cast<int>((n + block - 1) / block); }

}  // namespace

int graph_create(Engine* e) {
    GraphNet* m = new GraphNet();
    m->e = e;
    const BlobTensor* tt = e->find("graph.tensors");
    const BlobTensor* to = e->find("graph.ops");
    const BlobTensor* tm = e->find("graph.meta");
    if (!tt || !to || !tm || tt->dtype != 2 || to->dtype != 2 || tm->dtype != 2 || to->dims[1] != 12 || (tt->dims[1] != 2 && tt->dims[1] != 5)) {
        delete m;
        return set_err(e, DV_ERR_WEIGHTS, "graph model: missing graph.tensors / graph.ops / graph.meta");
    }
    std::vector<int32_t> ht(tt->nbytes / 4), ho(to->nbytes / 4), hm(tm->nbytes / 4);
    cudaMemcpy(ht.data(), tt->dptr, tt->nbytes, cudaMemcpyDeviceToHost);
    cudaMemcpy(ho.data(), to->dptr, to->nbytes, cudaMemcpyDeviceToHost);
    cudaMemcpy(hm.data(), tm->dptr, tm->nbytes, cudaMemcpyDeviceToHost);
    m->num_classes = hm[0];
    m->reg_bins = hm[1];
    m->head_ld = hm[2];
    m->kind = hm.size() > 5 ? hm[5] : 0;
    m->precise = e->find("precision") != nullptr;
    if (m->precise) m->pass_n = 512;  // fp32 buffers + the split scratch: ~4x the workspace per crop
    if (const char* ps = getenv("DV_REC_PASS")) {
        if (atoi(ps) > 0) m->pass_n = atoi(ps);
    }
    const int tcols = static_cast<int>(tt->dims[1]);  // 2: (c, down); 5: (c, down_h, down_w, pool_h, pool_w)
    for (size_t i = 0; i < tt->dims[0]; ++i) {
        const int32_t* r = &ht[tcols * i];
        m->tc.push_back(r[0]);
        m->tdh.push_back(r[1]);
        m->tdw.push_back(tcols == 5 ? r[2] : r[1]);
        m->tph.push_back(tcols == 5 ? r[3] : 1);
        m->tpw.push_back(tcols == 5 ? r[4] : 1);
        if (m->tdh.back() <= 0 || m->tdw.back() <= 0 || m->tph.back() < 0 || m->tpw.back() < 0) {
            delete m;
            return set_err(e, DV_ERR_WEIGHTS, "graph model: malformed tensor table");
        }
    }
    for (size_t i = 0; i < to->dims[0]; ++i) {
        const int32_t* o = &ho[12 * i];
        GOp op;
        op.code = o[0], op.in_t = o[1], op.in_coff = o[2], op.in_c = o[3], op.out_t = o[4], op.out_coff = o[5], op.out_c = o[6];
        op.k = o[7], op.stride = o[8], op.act = o[9], op.w = o[10], op.aux = o[11];
        const int nt = static_cast<int>(m->tc.size());
        if (op.in_t < 0 || op.in_t >= nt || op.out_t < 0 || op.out_t >= nt || (op.in_coff % 8) || (op.out_coff % 8) || (op.in_c % 8 && op.code != OP_STEM)) {
            delete m;
            return set_err(e, DV_ERR_WEIGHTS, "graph model: malformed op %zu", i);
        }
        m->ops.push_back(op);
    }
    e->model.reset(m);
    return 0;
}

double graph_flops(Engine* e) {
    GraphNet* m = dynamic_cast<GraphNet*>(e->model.get());
    return m ? (m->kind == 1 ? m->last_call_flops : m->flops) : 0.0;
}

int graph_num_classes(Engine* e) {
    GraphNet* m = dynamic_cast<GraphNet*>(e->model.get());
    return m ? m->num_classes : 0;
}

int graph_debug_tensor(Engine* e, int tensor_id, float* out_nchw, int* dims4) {
    GraphNet* m = dynamic_cast<GraphNet*>(e->model.get());
    if (!m) return set_err(e, DV_ERR_STATE, "not a graph-model handle");
    if (tensor_id <= 0 || tensor_id >= static_cast<int>(m->tens.size()) || !m->tens[tensor_id].p) return set_err(e, DV_ERR_ARG, "no tensor %d", tensor_id);
    const Tensor& t = m->tens[tensor_id];
    if (m->precise) return set_err(e, DV_ERR_UNSUPPORTED, "graph: tensor dumps are not available in the fp32x mode");
    if (dims4) {
        dims4[0] = t.N, dims4[1] = t.C, dims4[2] = t.H, dims4[3] = t.W;
    }
    if (out_nchw) return op_nhwc_f16_to_nchw_f32(e, t.p, t.N, t.C, t.H, t.W, out_nchw);
    return 0;
}

namespace {

// Outputs of a graph program: PicoDet heads (scores / dfl per level) or the recogniser's CTC head (probs / ids / maxp).
struct GraphOut {
    float* const* scores = nullptr;
    float* const* dfl = nullptr;
    float* probs = nullptr;
    int32_t* ids = nullptr;
    float* maxp = nullptr;
    float* logits = nullptr;
    float* prob = nullptr;  // OP_DBHEAD: fp32 probability map [N,1,H,W]
};

int run_graph(Engine* e, GraphNet* m, const float* in_nchw, const uint8_t* in_u8, const int32_t* widths, const float* mean3, const float* std3,
              float scale, int flip, int N, int H, int W, const GraphOut& go) {
    float* const* scores_out = go.scores;
    float* const* dfl_out = go.dfl;
    if (m->N != N || m->H != H || m->W != W) {
        m->park();
        if (!m->restore(N, H, W)) DV_TRY(build(e, m, N, H, W));
    }
    cudaStream_t s = e->stream;
    const bool pr = m->precise;
    auto F32 = [](const Tensor& t, int coff) { return reinterpret_cast<float*>(t.p) + coff; };  // fp32x: the buffers hold floats
    for (GOp& op : m->ops) {
        const Tensor& in = m->tens[op.in_t];
        const Tensor& out = m->tens[op.out_t];
        if (pr && (op.code == OP_ADD || op.code == OP_HEAD)) return set_err(e, DV_ERR_UNSUPPORTED, "graph: the fp32x mode does not cover PicoDet's ops");
        switch (op.code) {
            case OP_STEM: {
                const long long total = static_cast<long long>(N) * out.H * out.W;
                float3 mean = make_float3(0, 0, 0), stdv = make_float3(1, 1, 1);
                if (in_u8) {
                    mean = make_float3(mean3[0], mean3[1], mean3[2]);
                    stdv = make_float3(std3[0], std3[1], std3[2]);
                }
                e->launch_begin("k_stem3x3s2", "conv1", 2.0 * total * 27 * 16, total * (12.0 * (in_u8 ? 1 : 4) + 32.0));
                if (pr) k_stem3x3s2<float><<<grid_for(total, 256), 256, 0, s>>>(in_u8, in_nchw, N, H, W, out.H, out.W, mean, stdv, scale, flip, op.f0, op.f1, op.act,
                                                                                F32(out, 0), widths);
                else k_stem3x3s2<__half><<<grid_for(total, 256), 256, 0, s>>>(in_u8, in_nchw, N, H, W, out.H, out.W, mean, stdv, scale, flip, op.f0, op.f1, op.act,
                                                                              out.p, widths);
                e->launch_end();
                break;
            }
            case OP_DW: {
                static const bool profile_layers = getenv("DV_PROFILE_LAYERS") != nullptr;  // per-shape labels for tools/cascade_profile.py
                const long long total = static_cast<long long>(N) * out.H * out.W * (op.in_c / 8);
                // flops = 0: a depthwise conv is judged against the HBM roofline (k * k MACs per 4 bytes moved)
                e->launch_begin("k_dwconv", profile_layers ? "dw c" + std::to_string(op.in_c) + " k" + std::to_string(op.k) + " s" + std::to_string(op.sh()) + "x" +
                                                                  std::to_string(op.sw()) + " @" + std::to_string(out.H) + "x" + std::to_string(out.W)
                                                            : std::string("dw"),
                                0.0, total * 8 * 2.0 * (1.0 + 1.0 * (op.sh() * op.sw())));
                static const bool row_kernel = !(getenv("DV_DWROW") && atoi(getenv("DV_DWROW")) == 0);
                static const int dw_mode = getenv("DV_DWMODE") ? atoi(getenv("DV_DWMODE")) : 1;  // 1: k_dwconv_row (default), 2: k_dwconv_c2 (2x slower on the B200: 4-byte loads), 0: k_dwconv
                if (pr) {
                    k_dwconv<float><<<grid_for(total, 256), 256, 0, s>>>(F32(in, op.in_coff), N, in.H, in.W, op.in_c, in.C, op.k, op.sh(), op.sw(), out.H, out.W,
                                                                         op.f0, op.f1, op.act, op.ps, op.pb, F32(out, op.out_coff), out.C);
                } else if (row_kernel && dw_mode == 2 && op.sw() == 1 && (op.k == 3 || op.k == 5) && (op.sh() == 1 || op.sh() == 2)) {
#define DV_DWC2(KK, PP, RR, SS)                                                                                                                   \
    k_dwconv_c2<KK, PP, RR, SS><<<grid_for(static_cast<long long>(N) * ((out.H + RR - 1) / RR) * ((out.W + PP - 1) / PP) * (op.in_c / 2), 128), 128, 0, s>>>( \
        in.p + op.in_coff, N, in.H, in.W, op.in_c, in.C, out.H, out.W, op.f0, op.f1, op.act, op.ps, op.pb, out.p + op.out_coff, out.C)
                    if (op.k == 3 && op.sh() == 1) DV_DWC2(3, 8, 2, 1);
                    else if (op.k == 3) DV_DWC2(3, 8, 2, 2);
                    else if (op.sh() == 1) DV_DWC2(5, 8, 2, 1);
                    else DV_DWC2(5, 8, 2, 2);
#undef DV_DWC2
                } else if (row_kernel && op.sw() == 1 && (op.k == 3 || op.k == 5)) {
                    static const int dwp = getenv("DV_DWP") ? atoi(getenv("DV_DWP")) : 8;
#define DV_DWROW(KK, PP)                                                                                                                        \
    k_dwconv_row<KK, PP><<<grid_for(static_cast<long long>(N) * out.H * ((out.W + PP - 1) / PP) * (op.in_c / 8), 128), 128, 0, s>>>(              \
        in.p + op.in_coff, N, in.H, in.W, op.in_c, in.C, op.sh(), out.H, out.W, op.f0, op.f1, op.act, op.ps, op.pb, out.p + op.out_coff, out.C)
                    if (op.k == 3) {
                        if (dwp == 8) DV_DWROW(3, 8);
                        else DV_DWROW(3, 4);
                    } else {
                        if (dwp == 8) DV_DWROW(5, 8);
                        else DV_DWROW(5, 4);
                    }
#undef DV_DWROW
                } else {
                    k_dwconv<__half><<<grid_for(total, 256), 256, 0, s>>>(in.p + op.in_coff, N, in.H, in.W, op.in_c, in.C, op.k, op.sh(), op.sw(), out.H, out.W, op.f0,
                                                                  op.f1, op.act, op.ps, op.pb, out.p + op.out_coff, out.C);
                }
                e->launch_end();
                break;
            }
            case OP_PW:
                if (pr) {
                    const long long M = static_cast<long long>(N) * in.H * in.W;
                    e->launch_begin("k_split_f32", "split", 0.0, M * (4.0 * op.in_c + 4.0 * op.kp));
                    k_split_f32<<<grid_for(M * (op.kp / 8), 256), 256, 0, s>>>(F32(in, op.in_coff), M, op.in_c, op.kp, in.C, m->split_buf);
                    e->launch_end();
                }
                DV_TRY(launch_conv(e, op.plan));
                break;
            case OP_CONV:
            case OP_DECONV2:
                if (pr) {
                    const long long M = static_cast<long long>(N) * in.H * in.W;
                    e->launch_begin("k_split_f32", "split", 0.0, M * (4.0 * op.in_c + 4.0 * op.kp));
                    k_split_f32<<<grid_for(M * (op.kp / 8), 256), 256, 0, s>>>(F32(in, op.in_coff), M, op.in_c, op.kp, in.C, m->split_buf);
                    e->launch_end();
                }
                DV_TRY(launch_conv(e, op.plan));
                break;
            case OP_DBHEAD: {
                if (!go.prob) return set_err(e, DV_ERR_ARG, "graph: no probability-map output");
                const long long npix = static_cast<long long>(N) * in.H * in.W;
                e->launch_begin("k_dbhead", "head", 2.0 * npix * op.in_c * 4, npix * (2.0 * op.in_c + 16.0));
                if (pr) k_dbhead<float><<<grid_for(npix, 128), 128, 0, s>>>(F32(in, 0), npix, in.H, in.W, op.in_c, op.f0, op.f1, go.prob);
                else k_dbhead<__half><<<grid_for(npix, 128), 128, 0, s>>>(in.p, npix, in.H, in.W, op.in_c, op.f0, op.f1, go.prob);
                e->launch_end();
                break;
            }
            case OP_SE: {
                const int HW = in.H * in.W;
                if (HW >= kSePoolMin) {
                    const int nchunks = (HW + kSePoolRows - 1) / kSePoolRows, cv = op.in_c / 8;
                    e->launch_begin("k_se_pool", "se", 0.0, static_cast<double>(N) * HW * op.in_c * 2.0);
                    if (pr) k_se_pool<float><<<dim3(nchunks, N), 256, static_cast<size_t>(256 / cv) * op.in_c * sizeof(float), s>>>(F32(in, 0), HW, op.in_c, m->se_partial);
                    else k_se_pool<__half><<<dim3(nchunks, N), 256, static_cast<size_t>(256 / cv) * op.in_c * sizeof(float), s>>>(in.p, HW, op.in_c, m->se_partial);
                    e->launch_end();
                    e->launch_begin("k_se_scale", "se", 0.0, static_cast<double>(N) * nchunks * op.in_c * 4.0);
                    k_se_scale_p<<<N, 512, (op.in_c + op.in_c / 4) * sizeof(float), s>>>(m->se_partial, nchunks, HW, op.in_c, op.f0, op.f1, op.f2, op.f3,
                                                                                        m->se_scale);
                    e->launch_end();
                } else {
                    e->launch_begin("k_se_scale", "se", 0.0, static_cast<double>(N) * HW * op.in_c * 2.0);
                    if (pr) k_se_scale<float><<<N, 512, (op.in_c + op.in_c / 4) * sizeof(float), s>>>(F32(in, 0), HW, op.in_c, op.f0, op.f1, op.f2, op.f3, m->se_scale);
                    else k_se_scale<__half><<<N, 512, (op.in_c + op.in_c / 4) * sizeof(float), s>>>(in.p, HW, op.in_c, op.f0, op.f1, op.f2, op.f3, m->se_scale);
                    e->launch_end();
                }
                const long long total8 = static_cast<long long>(N) * HW * (op.in_c / 8);
                e->launch_begin("k_se_apply", "se", 0.0, total8 * 32.0);
                if (pr) k_se_apply<float><<<grid_for(total8, 256), 256, 0, s>>>(F32(in, 0), total8, HW, op.in_c, m->se_scale, op.k == 2 ? 1.f : 0.f, F32(out, op.out_coff), out.C);
                else k_se_apply<__half><<<grid_for(total8, 256), 256, 0, s>>>(in.p, total8, HW, op.in_c, m->se_scale, op.k == 2 ? 1.f : 0.f, out.p + op.out_coff, out.C);
                e->launch_end();
                break;
            }
            case OP_UP2: {
                const long long total = static_cast<long long>(N) * out.H * out.W * (op.in_c / 8);
                e->launch_begin("k_up2", "up", 0.0, total * 16.0 * 1.25);
                if (pr) k_up2<float><<<grid_for(total, 256), 256, 0, s>>>(F32(in, op.in_coff), N, in.H, in.W, op.in_c, in.C, out.H, out.W,
                                                                          op.aux >= 0 ? F32(m->tens[op.aux], 0) : nullptr, F32(out, op.out_coff), out.C);
                else k_up2<__half><<<grid_for(total, 256), 256, 0, s>>>(in.p + op.in_coff, N, in.H, in.W, op.in_c, in.C, out.H, out.W,
                                                                        op.aux >= 0 ? m->tens[op.aux].p : nullptr, out.p + op.out_coff, out.C);
                e->launch_end();
                break;
            }
            case OP_ADD: {
                const Tensor& b = m->tens[op.aux];
                const long long total8 = static_cast<long long>(out.elems() / 8);
                e->launch_begin("k_add", "add", 0.0, total8 * 48.0);
                k_add<<<grid_for(total8, 256), 256, 0, s>>>(in.p, b.p, total8, out.p);
                e->launch_end();
                break;
            }
            case OP_HEAD: {
                DV_TRY(launch_conv(e, op.plan));
                const long long M = static_cast<long long>(N) * in.H * in.W;
                if (!scores_out || !dfl_out || op.aux < 0 || op.aux > 3 || !scores_out[op.aux] || !dfl_out[op.aux])
                    return set_err(e, DV_ERR_ARG, "picodet_forward: null output for level %d", op.aux);
                e->launch_begin("k_head_split", "head", 0.0, M * (m->num_classes + m->reg_bins) * 8.0);
                k_head_split<<<grid_for(M * (m->num_classes + m->reg_bins), 256), 256, 0, s>>>(m->head_raw, M, m->head_ld, m->num_classes, m->reg_bins,
                                                                                               scores_out[op.aux], dfl_out[op.aux]);
                e->launch_end();
                break;
            }
            case OP_AVGPOOL: {
                const int kh = op.k == 0 ? in.H : (op.k & 255), kw = op.k == 0 ? in.W : (op.k >> 8);  // k = 0: global average pool
                const long long total = static_cast<long long>(N) * out.H * out.W * (op.in_c / 8);
                e->launch_begin("k_avgpool", "pool", 0.0, total * 16.0 * (kh * kw + 1));
                if (pr) k_avgpool<float><<<grid_for(total, 256), 256, 0, s>>>(F32(in, op.in_coff), N, in.H, in.W, op.in_c, in.C, kh, kw, out.H, out.W, F32(out, op.out_coff), out.C);
                else k_avgpool<__half><<<grid_for(total, 256), 256, 0, s>>>(in.p + op.in_coff, N, in.H, in.W, op.in_c, in.C, kh, kw, out.H, out.W, out.p + op.out_coff, out.C);
                e->launch_end();
                break;
            }
            case OP_UNFOLD3: {
                const long long total = static_cast<long long>(N) * in.W * 3 * (op.in_c / 8);
                e->launch_begin("k_unfold3", "unfold", 0.0, total * 32.0);
                if (pr) k_unfold3<float><<<grid_for(total, 256), 256, 0, s>>>(F32(in, op.in_coff), N, in.W, op.in_c, in.C, F32(out, 0));
                else k_unfold3<__half><<<grid_for(total, 256), 256, 0, s>>>(in.p + op.in_coff, N, in.W, op.in_c, in.C, out.p);
                e->launch_end();
                break;
            }
            case OP_LN: {
                const long long rows = static_cast<long long>(N) * in.H * in.W;
                e->launch_begin("k_ln_c", "ln", 0.0, rows * op.in_c * 4.0);
                if (pr) k_ln_c<float><<<grid_for(rows, 8), 256, 0, s>>>(F32(in, op.in_coff), rows, op.in_c, in.C, op.f0, op.f1, op.eps, F32(out, op.out_coff), out.C);
                else k_ln_c<__half><<<grid_for(rows, 8), 256, 0, s>>>(in.p + op.in_coff, rows, op.in_c, in.C, op.f0, op.f1, op.eps, out.p + op.out_coff, out.C);
                e->launch_end();
                break;
            }
            case OP_ATTN: {
                const int T = in.W, D = op.out_c;
                const size_t smem = static_cast<size_t>(2) * T * D * sizeof(float);
                if (smem > 200 * 1024) return set_err(e, DV_ERR_UNSUPPORTED, "graph: attention over %d positions exceeds shared memory", T);
                static DeviceOnce attr_once;
                if (attr_once.need(e->device)) {
                    DV_CUDA(e, cudaFuncSetAttribute(k_attn_small<__half, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                    DV_CUDA(e, cudaFuncSetAttribute(k_attn_small<float, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                    DV_CUDA(e, cudaFuncSetAttribute(k_attn_small<__half, 15>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                    DV_CUDA(e, cudaFuncSetAttribute(k_attn_small<float, 15>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                    attr_once.mark(e->device);
                }
                if (op.k <= 0 || D % op.k || D / op.k > 16) return set_err(e, DV_ERR_UNSUPPORTED, "graph: attention head width %d > 16", op.k > 0 ? D / op.k : 0);
                e->launch_begin("k_attn_small", "attn", 4.0 * N * T * T * D, static_cast<double>(N) * T * D * 8.0);
                if (D / op.k == 15) {
                    if (pr) k_attn_small<float, 15><<<N, 256, smem, s>>>(F32(in, 0), T, D, op.k, F32(out, 0));
                    else k_attn_small<__half, 15><<<N, 256, smem, s>>>(in.p, T, D, op.k, out.p);
                } else {
                    if (pr) k_attn_small<float, 0><<<N, 256, smem, s>>>(F32(in, 0), T, D, op.k, F32(out, 0));
                    else k_attn_small<__half, 0><<<N, 256, smem, s>>>(in.p, T, D, op.k, out.p);
                }
                e->launch_end();
                break;
            }
            case OP_CTC: {
                if (pr) {
                    const long long Mr = static_cast<long long>(N) * in.H * in.W;
                    e->launch_begin("k_split_f32", "split", 0.0, Mr * (4.0 * op.in_c + 4.0 * op.kp));
                    k_split_f32<<<grid_for(Mr * (op.kp / 8), 256), 256, 0, s>>>(F32(in, op.in_coff), Mr, op.in_c, op.kp, in.C, m->split_buf);
                    e->launch_end();
                }
                DV_TRY(launch_conv(e, op.plan));
                const long long M = static_cast<long long>(N) * in.H * in.W;
                if (!go.ids && !go.probs && !go.maxp && !go.logits) return set_err(e, DV_ERR_ARG, "graph head: no output requested");
                e->launch_begin("k_softmax_rows", "ctc", 0.0, M * m->num_classes * (go.probs ? 8.0 : 4.0));
                k_softmax_rows<<<grid_for(M, 8), 256, 0, s>>>(m->head_raw, M, m->head_ld, m->num_classes, go.probs, go.ids, go.maxp, go.logits);
                e->launch_end();
                break;
            }
            default: return set_err(e, DV_ERR_WEIGHTS, "graph model: unknown opcode %d", op.code);
        }
    }
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

}  // namespace

// scores_out[l] fp32 [N, HW_l, C], dfl_out[l] fp32 [N, HW_l, 32] (device), l = 0..3
int picodet_forward(Engine* e, const float* in_nchw, const uint8_t* in_u8, const float* mean3, const float* std3, float scale, int flip,
                    int N, int H, int W, float* const* scores_out, float* const* dfl_out) {
    GraphNet* m = dynamic_cast<GraphNet*>(e->model.get());
    if (!m || m->kind != 0) return set_err(e, DV_ERR_STATE, "handle was not created as a picodet model");
    if (N <= 0 || H <= 0 || W <= 0 || (!in_nchw && !in_u8) || !scores_out || !dfl_out) return set_err(e, DV_ERR_ARG, "picodet_forward: bad arguments");
    GraphOut go;
    go.scores = scores_out;
    go.dfl = dfl_out;
    return run_graph(e, m, in_nchw, in_u8, nullptr, mean3, std3, scale, flip, N, H, W, go);
}

// PP-OCR recogniser: fp32 [N,3,H,W] (PPOcrRecPreProcessor's batch) or uint8 [N,H,W,3] crops resized to height H, left aligned
// and valid up to widths[n] (normalisation (x/255 - 0.5)/0.5 and the zero padding fused into the stem) ->
// probs fp32 [N,T,C] (optional), ids int32 [N,T], maxp fp32 [N,T]; T = rec_time_steps(W).
int rec_forward(Engine* e, const float* in_nchw, const uint8_t* in_u8, const int32_t* widths, int N, int H, int W, float* probs, int32_t* ids,
                float* maxp) {
    GraphNet* m = dynamic_cast<GraphNet*>(e->model.get());
    if (!m || m->kind != 1) return set_err(e, DV_ERR_STATE, "handle was not created as a pp_rec model");
    if (N <= 0 || H <= 0 || W <= 0 || (!in_nchw && !in_u8)) return set_err(e, DV_ERR_ARG, "rec_forward: bad arguments");
    const float half3[3] = {0.5f, 0.5f, 0.5f};
    const int T = rec_time_steps(e, H, W);
    if (T <= 0) return set_err(e, DV_ERR_ARG, "rec_forward: a %dx%d input is too small for the network", H, W);
    // passes of at most pass_n crops (workspace bound); the crops are dealt evenly so that at most two plan shapes alternate
    // (both stay cached)
    const int passes = (N + m->pass_n - 1) / m->pass_n;
    const int chunk = (N + passes - 1) / passes;
    double flops = 0;
    for (int done = 0; done < N; done += chunk) {
        const int cur = N - done < chunk ? N - done : chunk;
        GraphOut go;
        go.probs = probs ? probs + static_cast<long long>(done) * T * m->num_classes : nullptr;
        go.ids = ids ? ids + static_cast<long long>(done) * T : nullptr;
        go.maxp = maxp ? maxp + static_cast<long long>(done) * T : nullptr;
        DV_TRY(run_graph(e, m, in_nchw ? in_nchw + static_cast<long long>(done) * 3 * H * W : nullptr,
                         in_u8 ? in_u8 + static_cast<long long>(done) * H * W * 3 : nullptr, widths ? widths + done : nullptr, half3, half3, 255.0f,
                         /*flip: divide by scale*/ 2, cur, H, W, go));
        flops += m->flops;
    }
    m->last_call_flops = flops;
    return 0;
}

// PULC image classifiers (PP-LCNet, cls/cls_pp_lcnet.py:164-293), model kind "pplcnet_cls": fp32 [N,3,H,W] (the image processor's
// pixel_values) -> logits fp32 [N,C] (PPLCNet.forward's return value) and / or softmax probabilities [N,C].
int cls_forward(Engine* e, const float* in_nchw, int N, int H, int W, float* logits, float* probs) {
    GraphNet* m = dynamic_cast<GraphNet*>(e->model.get());
    if (!m || m->kind != 2) return set_err(e, DV_ERR_STATE, "handle was not created as a pplcnet_cls model");
    if (N <= 0 || H <= 0 || W <= 0 || !in_nchw || (!logits && !probs)) return set_err(e, DV_ERR_ARG, "cls_forward: bad arguments");
    const float one3[3] = {1.f, 1.f, 1.f}, zero3[3] = {0.f, 0.f, 0.f};
    GraphOut go;
    go.logits = logits;
    go.probs = probs;
    return run_graph(e, m, in_nchw, nullptr, nullptr, zero3, one3, 1.f, 0, N, H, W, go);
}

// PP-OCRv4 detector: fp32 [N,3,H,W] (pre-processed) or uint8 [N,H,W,3] pages (flip / scale / mean / std fused into the stem as
// dbnet_forward does) -> probability map fp32 [N,1,H,W]; H and W multiples of 32 (DetResizeForTest).
int ppdet_forward(Engine* e, const float* in_nchw, const uint8_t* in_u8, const float* mean3, const float* std3, float scale, int flip, int N,
                  int H, int W, float* prob_out) {
    GraphNet* m = dynamic_cast<GraphNet*>(e->model.get());
    if (!m || m->kind != 3) return set_err(e, DV_ERR_STATE, "handle was not created as a pp_det model");
    if (N <= 0 || H <= 0 || W <= 0 || (H % 32) || (W % 32) || (!in_nchw && !in_u8) || !prob_out)
        return set_err(e, DV_ERR_ARG, "ppdet_forward: bad arguments (H and W must be multiples of 32)");
    const float one3[3] = {1.f, 1.f, 1.f}, zero3[3] = {0.f, 0.f, 0.f};
    GraphOut go;
    go.prob = prob_out;
    return run_graph(e, m, in_nchw, in_u8, nullptr, in_u8 ? mean3 : zero3, in_u8 ? std3 : one3, in_u8 ? scale : 1.f, in_u8 ? flip : 0, N, H, W, go);
}

int rec_time_steps(Engine* e, int H, int W) {
    GraphNet* m = dynamic_cast<GraphNet*>(e->model.get());
    if (!m || m->kind != 1 || m->ops.empty()) return 0;
    const int t = m->ops.back().in_t;
    return ((H + m->tdh[t] - 1) / m->tdh[t]) / m->tph[t] * (((W + m->tdw[t] - 1) / m->tdw[t]) / m->tpw[t]);
}

}  // namespace dv
