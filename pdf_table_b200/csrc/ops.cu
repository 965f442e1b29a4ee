// HBM-bound helper kernels around the tensor-core convolution: layout / dtype conversion,
// page normalisation (a1), 3x3 s2 max-pool, and the final ConvTranspose(64->1)+sigmoid of the DB head.
#include "engine.h"

namespace dv {

static inline int grid_for(long long n, int block) { return static_cast<int>((n + block - 1) / block); }

// fp32 NCHW [N,3,H,W] -> zero-bordered fp16 [N,H+6,W+8,4] (interior at +3,+3; channel 3 = 0)
// fp32x: lo > 0 = also write the residual halves fp16(v - hi) into a second image batch `lo` elements further
__device__ __forceinline__ void store_stem_px(__half* op, long long lo, float v0, float v1, float v2) {
    const __half2 a = __floats2half2_rn(v0, v1);
    const __half2 b = __floats2half2_rn(v2, 0.f);
    uint2 u;
    u.x = *reinterpret_cast<const uint32_t*>(&a);
    u.y = *reinterpret_cast<const uint32_t*>(&b);
    *reinterpret_cast<uint2*>(op) = u;
    if (lo > 0) {
        const float2 fa = __half22float2(a);
        const __half2 la = __floats2half2_rn(v0 - fa.x, v1 - fa.y);
        const __half2 lb = __floats2half2_rn(v2 - __low2float(b), 0.f);
        u.x = *reinterpret_cast<const uint32_t*>(&la);
        u.y = *reinterpret_cast<const uint32_t*>(&lb);
        *reinterpret_cast<uint2*>(op + lo) = u;
    }
}

__global__ void k_nchw_f32_to_stem(const float* __restrict__ in, int N, int H, int W, __half* __restrict__ out, long long lo) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long total = static_cast<long long>(N) * H * W;
    if (idx >= total) return;
    const int x = static_cast<int>(idx % W);
    const int y = static_cast<int>((idx / W) % H);
    const int n = static_cast<int>(idx / (static_cast<long long>(W) * H));
    const long long plane = static_cast<long long>(H) * W;
    const float* ip = in + static_cast<long long>(n) * 3 * plane + static_cast<long long>(y) * W + x;
    const int Hp = H + 6, Wp = W + 8;
    store_stem_px(out + ((static_cast<long long>(n) * Hp + y + 3) * Wp + x + 3) * 4, lo, ip[0], ip[plane], ip[2 * plane]);
}

int op_nchw_f32_to_stem(Engine* e, const float* in, int N, int H, int W, __half* out, long long lo) {
    const long long total = static_cast<long long>(N) * H * W;
    e->launch_begin("k_nchw_f32_to_stem", "pre", 0.0, total * (12.0 + (lo ? 16.0 : 8.0)));
    k_nchw_f32_to_stem<<<grid_for(total, 256), 256, 0, e->stream>>>(in, N, H, W, out, lo);
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

// uint8 HWC page -> normalised fp16 stem layout. Mirrors NormalizeImage
// (reference db_pp/image_operators.py:93-102): (x * scale - mean[c]) / std[c] in fp32, applied to the
// channel-flipped image (processor_ocr_db_pp.py:124) when flip != 0.
__global__ void k_u8_to_stem(const uint8_t* __restrict__ in, int N, int H, int W, float3 mean, float3 stdv,
                             float scale, int flip, __half* __restrict__ out, long long lo) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long total = static_cast<long long>(N) * H * W;
    if (idx >= total) return;
    const int x = static_cast<int>(idx % W);
    const int y = static_cast<int>((idx / W) % H);
    const int n = static_cast<int>(idx / (static_cast<long long>(W) * H));
    const uint8_t* ip = in + idx * 3;
    float c0 = static_cast<float>(ip[0]), c1 = static_cast<float>(ip[1]), c2 = static_cast<float>(ip[2]);
    if (flip) {
        const float t = c0;
        c0 = c2;
        c2 = t;
    }
    // no FMA contraction: numpy evaluates mul, sub, div as separate fp32 roundings
    const float v0 = __fdiv_rn(__fsub_rn(__fmul_rn(c0, scale), mean.x), stdv.x);
    const float v1 = __fdiv_rn(__fsub_rn(__fmul_rn(c1, scale), mean.y), stdv.y);
    const float v2 = __fdiv_rn(__fsub_rn(__fmul_rn(c2, scale), mean.z), stdv.z);
    const int Hp = H + 6, Wp = W + 8;
    store_stem_px(out + ((static_cast<long long>(n) * Hp + y + 3) * Wp + x + 3) * 4, lo, v0, v1, v2);
}

// PP-OCR recogniser pre-process after the host cv2.resize: PPOcrRecPreProcessor.resize_norm_img
// (ocr_rec_pp/processor_ocr_rec_pp.py:56-63): astype(float32) -> HWC->CHW -> / 255 -> -= 0.5 -> /= 0.5, all float32, then
// zero padding of the columns >= resized_w up to the batch width.  in: uint8 [B, H, W, 3] (each crop left-aligned, bytes
// beyond its width ignored), widths [B]; out: fp32 [B, 3, H, W].  One thread per pixel: 3 bytes read, 3 planes written.
__global__ void k_pp_rec_norm(const uint8_t* __restrict__ in, const int32_t* __restrict__ widths, int B, int H, int W,
                              float* __restrict__ out) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long plane = static_cast<long long>(H) * W;
    if (idx >= plane * B) return;
    const int b = static_cast<int>(idx / plane);
    const long long r = idx - b * plane;
    const int x = static_cast<int>(r % W);
    float v[3] = {0.f, 0.f, 0.f};
    if (x < __ldg(widths + b)) {
        const uint8_t* ip = in + idx * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c)
            v[c] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(ip[c]), 255.f), 0.5f), 0.5f);  // numpy's op order, no contraction
    }
    float* op = out + static_cast<long long>(b) * 3 * plane + r;
#pragma unroll
    for (int c = 0; c < 3; ++c) op[c * plane] = v[c];
}

int op_pp_rec_norm(Engine* e, const uint8_t* in, const int32_t* widths, int B, int H, int W, float* out) {
    const long long total = static_cast<long long>(B) * H * W;
    e->launch_begin("k_pp_rec_norm", "pp_rec_pre", 0.0, total * (3.0 + 12.0));
    k_pp_rec_norm<<<grid_for(total, 256), 256, 0, e->stream>>>(in, widths, B, H, W, out);
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

int op_u8_to_stem(Engine* e, const uint8_t* in, int N, int H, int W, const float* mean3, const float* std3,
                  float scale, int flip, __half* out, long long lo) {
    const long long total = static_cast<long long>(N) * H * W;
    e->launch_begin("k_u8_to_stem", "pre", 0.0, total * (3.0 + (lo ? 16.0 : 8.0)));
    k_u8_to_stem<<<grid_for(total, 256), 256, 0, e->stream>>>(
        in, N, H, W, make_float3(mean3[0], mean3[1], mean3[2]), make_float3(std3[0], std3[1], std3[2]), scale,
        flip, out, lo);
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

// 3x3 stride-2 pad-1 max-pool on NHWC fp16, 8 channels (16 B) per thread.
__global__ void k_maxpool3x3s2(const __half* __restrict__ in, int N, int H, int W, int C, int Ho, int Wo,
                               __half* __restrict__ out) {
    const int cv = C >> 3;
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long total = static_cast<long long>(N) * Ho * Wo * cv;
    if (idx >= total) return;
    unsigned t = static_cast<unsigned>(idx);  // 32-bit index decode (every launch covers < 2^32 work items): a 64-bit division is ~100 SASS instructions
    const int c8 = static_cast<int>(t % cv);
    t /= cv;
    const int ox = static_cast<int>(t % Wo);
    t /= Wo;
    const int oy = static_cast<int>(t % Ho);
    const int n = static_cast<int>(t / Ho);
    __half2 m[4];
    const __half2 ninf = __float2half2_rn(-65504.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) m[i] = ninf;
    for (int r = 0; r < 3; ++r) {
        const int iy = 2 * oy - 1 + r;
        if (iy < 0 || iy >= H) continue;
        for (int s = 0; s < 3; ++s) {
            const int ix = 2 * ox - 1 + s;
            if (ix < 0 || ix >= W) continue;
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(
                in + ((static_cast<long long>(n) * H + iy) * W + ix) * C + c8 * 8));
            const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
            for (int i = 0; i < 4; ++i) m[i] = __hmax2(m[i], h[i]);
        }
    }
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) oh[i] = m[i];
    *reinterpret_cast<uint4*>(out + ((static_cast<long long>(n) * Ho + oy) * Wo + ox) * C + c8 * 8) = o;
}

// fp32x variant: pixels are [hi(C) | lo(C)] pairs; the maximum is taken over hi + lo in fp32 and stored as a pair again
__global__ void k_maxpool3x3s2_split(const __half* __restrict__ in, int N, int H, int W, int C, int Ho, int Wo,
                                     __half* __restrict__ out) {
    const int cv = C >> 3;
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long total = static_cast<long long>(N) * Ho * Wo * cv;
    if (idx >= total) return;
    unsigned t = static_cast<unsigned>(idx);  // 32-bit index decode (every launch covers < 2^32 work items): a 64-bit division is ~100 SASS instructions
    const int c8 = static_cast<int>(t % cv);
    t /= cv;
    const int ox = static_cast<int>(t % Wo);
    t /= Wo;
    const int oy = static_cast<int>(t % Ho);
    const int n = static_cast<int>(t / Ho);
    float m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
    for (int r = 0; r < 3; ++r) {
        const int iy = 2 * oy - 1 + r;
        if (iy < 0 || iy >= H) continue;
        for (int s = 0; s < 3; ++s) {
            const int ix = 2 * ox - 1 + s;
            if (ix < 0 || ix >= W) continue;
            const __half* ip = in + ((static_cast<long long>(n) * H + iy) * W + ix) * 2 * C + c8 * 8;
            const uint4 uh = __ldg(reinterpret_cast<const uint4*>(ip));
            const uint4 ul = __ldg(reinterpret_cast<const uint4*>(ip + C));
            const __half2* hh = reinterpret_cast<const __half2*>(&uh);
            const __half2* hl = reinterpret_cast<const __half2*>(&ul);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 a = __half22float2(hh[i]), b = __half22float2(hl[i]);
                m[2 * i] = fmaxf(m[2 * i], a.x + b.x);
                m[2 * i + 1] = fmaxf(m[2 * i + 1], a.y + b.y);
            }
        }
    }
    uint4 oh, ol;
    __half2* ph = reinterpret_cast<__half2*>(&oh);
    __half2* pl = reinterpret_cast<__half2*>(&ol);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        ph[i] = __floats2half2_rn(m[2 * i], m[2 * i + 1]);
        const float2 f = __half22float2(ph[i]);
        pl[i] = __floats2half2_rn(m[2 * i] - f.x, m[2 * i + 1] - f.y);
    }
    __half* op = out + ((static_cast<long long>(n) * Ho + oy) * Wo + ox) * 2 * C + c8 * 8;
    *reinterpret_cast<uint4*>(op) = oh;
    *reinterpret_cast<uint4*>(op + C) = ol;
}

int op_maxpool3x3s2(Engine* e, const Tensor& in, Tensor& out) {
    if (in.C % 8) return set_err(e, DV_ERR_UNSUPPORTED, "maxpool: C %% 8 != 0");
    if (in.lo > 0) {
        if (in.lo != in.C || in.ldc() != 2 * in.C || out.lo != out.C || out.ldc() != 2 * out.C)
            return set_err(e, DV_ERR_UNSUPPORTED, "maxpool: split tensors must be dense [hi | lo] pixels");
        const long long total = static_cast<long long>(out.N) * out.H * out.W * (in.C / 8);
        e->launch_begin("k_maxpool3x3s2", "maxpool", 0.0, 4.0 * (double)in.elems() + 4.0 * (double)out.N * out.H * out.W * in.C);
        k_maxpool3x3s2_split<<<grid_for(total, 256), 256, 0, e->stream>>>(in.p, in.N, in.H, in.W, in.C, out.H, out.W, out.p);
        e->launch_end();
        DV_CUDA(e, cudaGetLastError());
        return 0;
    }
    const long long total = static_cast<long long>(out.N) * out.H * out.W * (in.C / 8);
    e->launch_begin("k_maxpool3x3s2", "maxpool", 0.0, 2.0 * (double)in.elems() + 2.0 * (double)out.N * out.H * out.W * in.C);
    k_maxpool3x3s2<<<grid_for(total, 256), 256, 0, e->stream>>>(in.p, in.N, in.H, in.W, in.C, out.H, out.W, out.p);
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

// ConvTranspose2d(64 -> 1, k=2, s=2) + bias + sigmoid: each input pixel produces a 2x2 block of the
// fp32 probability map (reference db_net/dbnet.py:539). One thread per input pixel; weights w[c][dy*2+dx].
// SPLIT (fp32x): input pixels are [hi(64) | lo(64)] pairs and the weights come as fp32 (w32).
template <bool SPLIT>
__global__ void k_deconv2x2_c1_sigmoid(const __half* __restrict__ in, long long npix, int H, int W,
                                       const __half* __restrict__ w, const float* __restrict__ w32, float bias, float* __restrict__ out) {
    __shared__ float sw[64 * 4];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sw[i] = SPLIT ? w32[i] : __half2float(w[i]);
    __syncthreads();
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= npix) return;
    const int x = static_cast<int>(idx % W);
    const int y = static_cast<int>((idx / W) % H);
    const long long n = idx / (static_cast<long long>(W) * H);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const uint4* ip = reinterpret_cast<const uint4*>(in + idx * (SPLIT ? 128 : 64));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint4 u = __ldg(ip + j);
        const __half2* h = reinterpret_cast<const __half2*>(&u);
        uint4 ul = make_uint4(0u, 0u, 0u, 0u);
        if constexpr (SPLIT) ul = __ldg(ip + 8 + j);
        const __half2* hl = reinterpret_cast<const __half2*>(&ul);
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
            float2 f = __half22float2(h[e2]);
            if constexpr (SPLIT) {
                const float2 g = __half22float2(hl[e2]);
                f.x += g.x;
                f.y += g.y;
            }
            const int c = j * 8 + e2 * 2;
            a0 = fmaf(f.x, sw[c * 4 + 0], a0);
            a1 = fmaf(f.x, sw[c * 4 + 1], a1);
            a2 = fmaf(f.x, sw[c * 4 + 2], a2);
            a3 = fmaf(f.x, sw[c * 4 + 3], a3);
            a0 = fmaf(f.y, sw[c * 4 + 4], a0);
            a1 = fmaf(f.y, sw[c * 4 + 5], a1);
            a2 = fmaf(f.y, sw[c * 4 + 6], a2);
            a3 = fmaf(f.y, sw[c * 4 + 7], a3);
        }
    }
    const int W2 = 2 * W;
    float* op = out + (n * 2 * H + 2 * y) * W2 + 2 * x;
    const float s0 = 1.f / (1.f + expf(-(a0 + bias)));
    const float s1 = 1.f / (1.f + expf(-(a1 + bias)));
    const float s2 = 1.f / (1.f + expf(-(a2 + bias)));
    const float s3 = 1.f / (1.f + expf(-(a3 + bias)));
    *reinterpret_cast<float2*>(op) = make_float2(s0, s1);
    *reinterpret_cast<float2*>(op + W2) = make_float2(s2, s3);
}

int op_deconv2x2_c1_sigmoid(Engine* e, const Tensor& in, const __half* w, const float* w32, float bias, float* out) {
    if (in.C != 64) return set_err(e, DV_ERR_UNSUPPORTED, "deconv2x2_c1: C != 64");
    if (in.lo > 0 && (in.lo != 64 || in.ldc() != 128 || !w32)) return set_err(e, DV_ERR_UNSUPPORTED, "deconv2x2_c1: split input needs [hi | lo] pixels and fp32 weights");
    const long long npix = static_cast<long long>(in.N) * in.H * in.W;
    e->launch_begin("k_deconv2x2_c1_sigmoid", "bin.deconv2", 2.0 * npix * 64 * 4, npix * ((in.lo ? 256.0 : 128.0) + 16.0));
    if (in.lo > 0) k_deconv2x2_c1_sigmoid<true><<<grid_for(npix, 128), 128, 0, e->stream>>>(in.p, npix, in.H, in.W, w, w32, bias, out);
    else k_deconv2x2_c1_sigmoid<false><<<grid_for(npix, 128), 128, 0, e->stream>>>(in.p, npix, in.H, in.W, w, w32, bias, out);
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

// ---- generic layout conversion for the operator-level ABI (dv_conv2d_nhwc_f16 tests)
__global__ void k_nchw_f32_to_nhwc_f16(const float* __restrict__ in, int N, int C, int H, int W,
                                       __half* __restrict__ out) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long total = static_cast<long long>(N) * C * H * W;
    if (idx >= total) return;
    unsigned t = static_cast<unsigned>(idx);  // 32-bit index decode (every launch covers < 2^32 work items): a 64-bit division is ~100 SASS instructions
    const int c = static_cast<int>(t % C);
    t /= C;
    const int x = static_cast<int>(t % W);
    t /= W;
    const int y = static_cast<int>(t % H);
    const int n = static_cast<int>(t / H);
    out[idx] = __float2half_rn(in[((static_cast<long long>(n) * C + c) * H + y) * W + x]);
}
__global__ void k_nhwc_f16_to_nchw_f32(const __half* __restrict__ in, int N, int C, int H, int W,
                                       float* __restrict__ out, int ld, int lo, int Hp, int Wp, int py, int px) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long total = static_cast<long long>(N) * C * H * W;
    if (idx >= total) return;
    unsigned t = static_cast<unsigned>(idx);  // 32-bit index decode (every launch covers < 2^32 work items): a 64-bit division is ~100 SASS instructions
    const int x = static_cast<int>(t % W);
    t /= W;
    const int y = static_cast<int>(t % H);
    t /= H;
    const int c = static_cast<int>(t % C);
    const int n = static_cast<int>(t / C);
    // Hp x Wp = the rows / pixels per row of the buffer in memory (a zero-bordered tensor: interior at (py, px))
    const __half* ip = in + ((static_cast<long long>(n) * Hp + y + py) * Wp + x + px) * ld + c;
    out[idx] = __half2float(ip[0]) + (lo ? __half2float(ip[lo]) : 0.f);
}
int op_nchw_f32_to_nhwc_f16(Engine* e, const float* in, int N, int C, int H, int W, __half* out) {
    const long long total = static_cast<long long>(N) * C * H * W;
    e->launch_begin("k_nchw_f32_to_nhwc_f16", "layout", 0.0, total * 6.0);
    k_nchw_f32_to_nhwc_f16<<<grid_for(total, 256), 256, 0, e->stream>>>(in, N, C, H, W, out);
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}
int op_nhwc_f16_to_nchw_f32(Engine* e, const __half* in, int N, int C, int H, int W, float* out, int ld, int lo, int Hp, int Wp, int py, int px) {
    const long long total = static_cast<long long>(N) * C * H * W;
    e->launch_begin("k_nhwc_f16_to_nchw_f32", "layout", 0.0, total * 6.0);
    k_nhwc_f16_to_nchw_f32<<<grid_for(total, 256), 256, 0, e->stream>>>(in, N, C, H, W, out, ld ? ld : C, lo, Hp ? Hp : H, Wp ? Wp : W, py, px);
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

}  // namespace dv
