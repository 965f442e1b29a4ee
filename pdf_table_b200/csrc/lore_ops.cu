// HBM-bound kernels of the Lore / CenterNet table-structure detector around the tensor-core convolution:
//   k_img_to_stem8        TableLorePreProcessor normalisation (lore/processer_lore.py:66-93) fused with the layout
//                         change into the zero-bordered 8-channel image the stride-1 7x7 stem reads
//   k_maxpool2x2          nn.MaxPool2d(2, 2) of Tree.downsample / the level roots (center_net/modeling_centernet.py:253)
//   k_dcn_im2col          modulated deformable sampling of DCN (lore/dcnv2.py:71-86 -> torchvision deform_conv2d):
//                         col[m][tap*C + c] = mask * bilinear(x, y + dy, x + dx); the 9C-wide GEMM runs on tcgen05
//   k_up_dw_add           IDAUp: depthwise ConvTranspose2d(k=2f, s=f, p=f/2) + the skip addition
//                         (lore/lore_dla_34.py:93-110)
//   k_sigmoid_cols        hm.sigmoid_() (lore/lineless_table_process.py:599) on the first columns of the packed maps
//   k_gather_patch3x3     3x3 zero-padded patches of the 64-channel feature map at the decoded cell / corner points,
//                         so the `ax` / `cr` heads are evaluated only where the reference gathers them
//   k_logi_combine        logi = ax(centre) + sum of cr(4 corners)
#include "engine.h"

namespace dv {

static inline int grid_for(long long n, int block) { return static_cast<int>((n + block - 1) / block); }

// fp32x (split-fp16) helpers: a value = hi + lo, both fp16, `lo` elements apart (engine.h Tensor::lo)
__device__ __forceinline__ void load8_split(const __half* p, long long lo, float* v) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(p + lo));
    const __half2 *ha = reinterpret_cast<const __half2*>(&a), *hb = reinterpret_cast<const __half2*>(&b);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 x = __half22float2(ha[i]), y = __half22float2(hb[i]);
        v[2 * i] = x.x + y.x;
        v[2 * i + 1] = x.y + y.y;
    }
}
__device__ __forceinline__ void store8_split(__half* p, long long lo, const float* v) {
    uint4 a, b;
    __half2 *ha = reinterpret_cast<__half2*>(&a), *hb = reinterpret_cast<__half2*>(&b);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        ha[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        const float2 f = __half22float2(ha[i]);
        hb[i] = __floats2half2_rn(v[2 * i] - f.x, v[2 * i + 1] - f.y);
    }
    *reinterpret_cast<uint4*>(p) = a;
    *reinterpret_cast<uint4*>(p + lo) = b;
}

// ---------------------------------------------------------------------------------------------- stem input
// in: uint8 HWC [N,H,W,3] or fp32 NCHW [N,3,H,W] (already normalised) -> fp16 [N, H+6, W+8, 8], interior at (+3,+3).
// The u8 path mirrors numpy: (x / 255. - mean) / std evaluated in float64, then .astype(float32).
constexpr int kStemPix = 8;  // pixels per thread: the per-CTA table of 768 float64 normalisations is amortised over 2048 pixels
__global__ void __launch_bounds__(256)
k_img_to_stem8(const uint8_t* __restrict__ u8, const float* __restrict__ f32, int N, int H, int W, double3 mean, double3 stdv,
               int flip, __half* __restrict__ out, long long lo, int cpp) {
    // u8 input: (x / 255. - mean) / std in float64 -> float32 for every (channel, byte) once per CTA (three double divisions
    // per pixel otherwise: the kernel ran at a quarter of its HBM bound)
    __shared__ float lut[3][256];
    if (u8) {
        for (int i = threadIdx.x; i < 768; i += blockDim.x) {
            const int ci = i >> 8;
            const double m = ci == 0 ? mean.x : ci == 1 ? mean.y : mean.z, sd = ci == 0 ? stdv.x : ci == 1 ? stdv.y : stdv.z;
            lut[ci][i & 255] = static_cast<float>(__ddiv_rn(__dsub_rn(__ddiv_rn(static_cast<double>(i & 255), 255.0), m), sd));
        }
        __syncthreads();
    }
    const long long total = static_cast<long long>(N) * H * W;
    const long long base = static_cast<long long>(blockIdx.x) * (256 * kStemPix) + threadIdx.x;
#pragma unroll 1
    for (int it = 0; it < kStemPix; ++it) {
    const long long idx = base + static_cast<long long>(it) * 256;
    if (idx >= total) return;
    unsigned t = static_cast<unsigned>(idx);  // 32-bit index decode (a launch covers < 2^32 pixels)
    const int x = static_cast<int>(t % W);
    t /= W;
    const int y = static_cast<int>(t % H), n = static_cast<int>(t / H);
    float v0, v1, v2;
    if (u8) {
        const uint8_t* ip = u8 + idx * 3;
        const int b0 = ip[0], b1 = ip[1], b2 = ip[2];
        v0 = lut[0][flip ? b2 : b0];
        v1 = lut[1][b1];
        v2 = lut[2][flip ? b0 : b2];
    } else {
        const long long plane = static_cast<long long>(H) * W;
        const float* ip = f32 + static_cast<long long>(n) * 3 * plane + static_cast<long long>(y) * W + x;
        v0 = ip[0];
        v1 = ip[plane];
        v2 = ip[2 * plane];
    }
    const __half2 a = __floats2half2_rn(v0, v1);
    const __half2 b = __floats2half2_rn(v2, 0.f);
    uint4 u;
    u.x = *reinterpret_cast<const uint32_t*>(&a);
    u.y = *reinterpret_cast<const uint32_t*>(&b);
    u.z = 0u;
    u.w = 0u;
    const int Hp = H + 6, Wp = W + 8;
    __half* op = out + ((static_cast<long long>(n) * Hp + y + 3) * Wp + x + 3) * cpp;
    // cpp = 8: the stride-1 stem of DLA-34 (8-channel pixels); cpp = 4: the stride-2 stem of the ResNet-18 detector (A_STEM view)
    if (cpp == 8) *reinterpret_cast<uint4*>(op) = u;
    else *reinterpret_cast<uint2*>(op) = make_uint2(u.x, u.y);
    if (lo > 0) {  // fp32x: the residual halves go to a second image batch `lo` elements further
        const float2 fa = __half22float2(a);
        const __half2 la = __floats2half2_rn(v0 - fa.x, v1 - fa.y);
        const __half2 lb = __floats2half2_rn(v2 - __low2float(b), 0.f);
        u.x = *reinterpret_cast<const uint32_t*>(&la);
        u.y = *reinterpret_cast<const uint32_t*>(&lb);
        if (cpp == 8) *reinterpret_cast<uint4*>(op + lo) = u;
        else *reinterpret_cast<uint2*>(op + lo) = make_uint2(u.x, u.y);
    }
    }
}

int op_img_to_stem8(Engine* e, const uint8_t* u8, const float* f32, int N, int H, int W, const float* mean3, const float* std3,
                    int flip, __half* out, long long lo, int cpp) {
    const long long total = static_cast<long long>(N) * H * W;
    double3 m = make_double3(0, 0, 0), s = make_double3(1, 1, 1);
    if (u8) {
        // the reference holds mean / std as float32 arrays; float32 -> float64 promotion is exact
        m = make_double3(mean3[0], mean3[1], mean3[2]);
        s = make_double3(std3[0], std3[1], std3[2]);
    }
    e->launch_begin("k_img_to_stem8", "pre", 0.0, total * ((u8 ? 3.0 : 12.0) + 2.0 * cpp));
    k_img_to_stem8<<<grid_for(total, 256 * kStemPix), 256, 0, e->stream>>>(u8, f32, N, H, W, m, s, flip, out, lo, cpp);
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------- 2x2 max-pool
__global__ void __launch_bounds__(256)
k_maxpool2x2(const __half* __restrict__ in, int N, int H, int W, int C, int ldi, __half* __restrict__ out, int ldo) {
    const int cv = C >> 3, Ho = H >> 1, Wo = W >> 1;
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
    const __half* ip = in + ((static_cast<long long>(n) * H + 2 * oy) * W + 2 * ox) * ldi + c8 * 8;
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(ip));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(ip + ldi));
    const uint4 c = __ldg(reinterpret_cast<const uint4*>(ip + static_cast<long long>(W) * ldi));
    const uint4 d = __ldg(reinterpret_cast<const uint4*>(ip + static_cast<long long>(W) * ldi + ldi));
    uint4 o;
    const __half2 *ha = reinterpret_cast<const __half2*>(&a), *hb = reinterpret_cast<const __half2*>(&b);
    const __half2 *hc = reinterpret_cast<const __half2*>(&c), *hd = reinterpret_cast<const __half2*>(&d);
    __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) ho[i] = __hmax2(__hmax2(ha[i], hb[i]), __hmax2(hc[i], hd[i]));
    *reinterpret_cast<uint4*>(out + ((static_cast<long long>(n) * Ho + oy) * Wo + ox) * ldo + c8 * 8) = o;
}

__global__ void __launch_bounds__(256)
k_maxpool2x2_split(const __half* __restrict__ in, int N, int H, int W, int C, int ldi, long long loi, __half* __restrict__ out, int ldo, long long loo) {
    const int cv = C >> 3, Ho = H >> 1, Wo = W >> 1;
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= static_cast<long long>(N) * Ho * Wo * cv) return;
    unsigned t = static_cast<unsigned>(idx);  // 32-bit index decode (every launch covers < 2^32 work items): a 64-bit division is ~100 SASS instructions
    const int c8 = static_cast<int>(t % cv);
    t /= cv;
    const int ox = static_cast<int>(t % Wo);
    t /= Wo;
    const int oy = static_cast<int>(t % Ho);
    const int n = static_cast<int>(t / Ho);
    const __half* ip = in + ((static_cast<long long>(n) * H + 2 * oy) * W + 2 * ox) * ldi + c8 * 8;
    float a[8], b[8];
    load8_split(ip, loi, a);
    load8_split(ip + ldi, loi, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fmaxf(a[i], b[i]);
    load8_split(ip + static_cast<long long>(W) * ldi, loi, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fmaxf(a[i], b[i]);
    load8_split(ip + static_cast<long long>(W) * ldi + ldi, loi, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fmaxf(a[i], b[i]);
    store8_split(out + ((static_cast<long long>(n) * Ho + oy) * Wo + ox) * ldo + c8 * 8, loo, a);
}

int op_maxpool2x2(Engine* e, const Tensor& in, const Tensor& out) {
    if ((in.C % 8) || (in.H & 1) || (in.W & 1) || out.C != in.C || out.H != in.H / 2 || out.W != in.W / 2 || (in.lo > 0) != (out.lo > 0))
        return set_err(e, DV_ERR_UNSUPPORTED, "maxpool2x2: bad shapes");
    const long long total = static_cast<long long>(out.N) * out.H * out.W * (in.C / 8);
    if (in.lo > 0) {
        e->launch_begin("k_maxpool2x2", "maxpool", 0.0, 4.0 * (double)in.elems() * 1.25);
        k_maxpool2x2_split<<<grid_for(total, 256), 256, 0, e->stream>>>(in.p, in.N, in.H, in.W, in.C, in.ldc(), in.lo, out.p, out.ldc(), out.lo);
        e->launch_end();
        DV_CUDA(e, cudaGetLastError());
        return 0;
    }
    e->launch_begin("k_maxpool2x2", "maxpool", 0.0, 2.0 * (double)in.elems() * 1.25);
    k_maxpool2x2<<<grid_for(total, 256), 256, 0, e->stream>>>(in.p, in.N, in.H, in.W, in.C, in.ldc(), out.p, out.ldc());
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------- DCN sampling
// om: fp32 [M, 32] = conv_offset_mask output; columns 2k, 2k+1 = (dy, dx) of tap k, 18+k = mask logit (dcnv2.py:73-76).
// One thread = (pixel, 8 channels), looping over the nine taps; grid = (pixel blocks, image rows).  Sampling rule =
// torchvision deform_conv2d_kernel bilinear_interpolate.
//
// History of this kernel, all from ncu (profiles/r1m, r1x): v1 was one thread per (pixel, tap, 8 ch) with flat 64-bit
// indices -- 450 instructions per thread, mostly index arithmetic.  v2 looped over the taps with 32-bit 2-D indices
// (5x fewer), but every one of the C/8 threads of a pixel still recomputed the same sigmoid, floor, clamps and corner
// weights, and each fp16 sample cost a convert plus an FFMA: 1861 instructions per thread at 78 % issue utilisation.
// v3 (this one): the (pixel, tap) jobs of a warp are computed ONCE, spread over its lanes, and handed to the channel
// threads by three shuffles (corner-0 offset with the two clamped steps in its top bits, and the four mask-scaled
// weights as two half2), and the blend uses the mixed-precision FHFMA (fp16 sample x fp16 weight + fp32 accumulator).
// The weights are therefore rounded to fp16 (2^-11 relative) before the blend; the output is fp16 as before.
__device__ __forceinline__ float fhfma(unsigned short a, unsigned short b, float c) {
    asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(c) : "h"(a), "h"(b));
    return c;
}

__global__ void __launch_bounds__(256)
k_dcn_im2col(const __half* __restrict__ in, int H, int W, int lcv, int ldi, const float* __restrict__ om, __half* __restrict__ col) {
    const int cv = 1 << lcv, C = cv << 3;
    const int t = blockIdx.x * 256 + threadIdx.x, lane = threadIdx.x & 31;
    const int x = t >> lcv, c8 = t & (cv - 1);
    const bool active = x < W;
    const int row = blockIdx.y;  // n * H + y
    const int y = row % H;
    const int lppw = lcv >= 5 ? 0 : 5 - lcv;  // log2(pixels per warp)
    const int xw = (t - lane) >> lcv;         // first pixel of this warp
    const int pw = lcv >= 5 ? 0 : lane >> lcv;
    const int njobs = 9 << lppw;

    // ---- the warp's (tap, pixel) jobs, job = tap * ppw + pixel-in-warp, lane `job & 31` of round `job >> 5`
    uint32_t jp[2] = {0u, 0u}, jw01[2] = {0u, 0u}, jw23[2] = {0u, 0u};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int job = r * 32 + lane;
        const int tap = job >> lppw, jx = xw + (job & ((1 << lppw) - 1));
        if (job < njobs && jx < W) {
            const float* o = om + static_cast<size_t>(row * W + jx) * 32;
            const float dy = __ldg(o + 2 * tap), dx = __ldg(o + 2 * tap + 1);
            const float mask = 1.f / (1.f + expf(-__ldg(o + 18 + tap)));
            const int ky = (tap * 11) >> 5, kx = tap - 3 * ky;
            const float py = static_cast<float>(y + ky - 1) + dy;
            const float px = static_cast<float>(jx + kx - 1) + dx;
            const bool inside = py > -1.f && py < static_cast<float>(H) && px > -1.f && px < static_cast<float>(W);
            const float fy = floorf(py), fx = floorf(px);
            const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
            const float ly = py - fy, lx = px - fx, hy = 1.f - ly, hx = 1.f - lx;
            const bool oky0 = inside && y0 >= 0 && y0 <= H - 1, oky1 = inside && y0 + 1 >= 0 && y0 + 1 <= H - 1;
            const bool okx0 = x0 >= 0 && x0 <= W - 1, okx1 = x0 + 1 >= 0 && x0 + 1 <= W - 1;
            const float w00 = (oky0 && okx0) ? hy * hx * mask : 0.f, w01 = (oky0 && okx1) ? hy * lx * mask : 0.f;
            const float w10 = (oky1 && okx0) ? ly * hx * mask : 0.f, w11 = (oky1 && okx1) ? ly * lx * mask : 0.f;
            // clamped addresses (zero weight where clamped)
            const int cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y0 + 1, 0), H - 1);
            const int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x0 + 1, 0), W - 1);
            jp[r] = static_cast<uint32_t>(cy0 * W + cx0) | (static_cast<uint32_t>(cx1 - cx0) << 30) |
                    (static_cast<uint32_t>(cy1 - cy0) << 31);
            const __half2 h01 = __floats2half2_rn(w00, w01), h23 = __floats2half2_rn(w10, w11);
            jw01[r] = *reinterpret_cast<const uint32_t*>(&h01);
            jw23[r] = *reinterpret_cast<const uint32_t*>(&h23);
        }
        if (njobs <= 32) break;
    }

    const __half* base = in + static_cast<size_t>(row - y) * W * ldi + c8 * 8;  // image origin + channel group
    __half* dst = col + static_cast<size_t>(row * W + x) * 9 * C + c8 * 8;
    const int rstep = W * ldi;
    // three taps per round: their twelve 16-byte corner loads are issued before any is consumed
#pragma unroll 1
    for (int t0 = 0; t0 < 9; t0 += 3) {
        uint4 u[3][4];
        uint32_t w01[3], w23[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const int job = ((t0 + q) << lppw) + pw;
            const bool r1 = (job >> 5) != 0;  // uniform over the warp
            const uint32_t p = __shfl_sync(0xffffffffu, r1 ? jp[1] : jp[0], job & 31);
            w01[q] = __shfl_sync(0xffffffffu, r1 ? jw01[1] : jw01[0], job & 31);
            w23[q] = __shfl_sync(0xffffffffu, r1 ? jw23[1] : jw23[0], job & 31);
            const __half* b00 = base + static_cast<size_t>(p & 0x3fffffffu) * ldi;
            const int sx = (p & 0x40000000u) ? ldi : 0, sy = (p & 0x80000000u) ? rstep : 0;
            if (active) {
                u[q][0] = __ldg(reinterpret_cast<const uint4*>(b00));
                u[q][1] = __ldg(reinterpret_cast<const uint4*>(b00 + sx));
                u[q][2] = __ldg(reinterpret_cast<const uint4*>(b00 + sy));
                u[q][3] = __ldg(reinterpret_cast<const uint4*>(b00 + sy + sx));
            }
        }
        if (!active) continue;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            float acc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t wp = (k < 2) ? w01[q] : w23[q];
                const unsigned short wk = static_cast<unsigned short>((k & 1) ? (wp >> 16) : (wp & 0xffffu));
                const uint32_t* h = reinterpret_cast<const uint32_t*>(&u[q][k]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc[2 * i] = fhfma(static_cast<unsigned short>(h[i] & 0xffffu), wk, acc[2 * i]);
                    acc[2 * i + 1] = fhfma(static_cast<unsigned short>(h[i] >> 16), wk, acc[2 * i + 1]);
                }
            }
            uint4 out;
            __half2* ho = reinterpret_cast<__half2*>(&out);
#pragma unroll
            for (int i = 0; i < 4; ++i) ho[i] = __floats2half2_rn(acc[2 * i], acc[2 * i + 1]);
            *reinterpret_cast<uint4*>(dst + (t0 + q) * C) = out;
        }
    }
}

// fp32x variant: one thread = (pixel, 8 channels), nine taps in a loop; samples are hi + lo in fp32, blend weights stay fp32
// (torchvision's bilinear_interpolate arithmetic), the column row is written as [hi(9C) | lo(9C)].
__global__ void __launch_bounds__(256)
k_dcn_im2col_split(const __half* __restrict__ in, int N, int H, int W, int C, int ldi, long long lo, const float* __restrict__ om,
                   __half* __restrict__ col) {
    const int cv = C >> 3;
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= static_cast<long long>(N) * H * W * cv) return;
    const int c8 = static_cast<int>(idx % cv);
    const long long pix = idx / cv;
    const int x = static_cast<int>(pix % W), y = static_cast<int>((pix / W) % H);
    const long long n = pix / (static_cast<long long>(W) * H);
    const float* o = om + pix * 32;
    const __half* base = in + n * H * W * ldi + c8 * 8;
    __half* dst = col + pix * 18 * C + c8 * 8;
    for (int tap = 0; tap < 9; ++tap) {
        const float dy = __ldg(o + 2 * tap), dx = __ldg(o + 2 * tap + 1);
        const float mask = 1.f / (1.f + expf(-__ldg(o + 18 + tap)));
        const int ky = tap / 3, kx = tap - 3 * ky;
        const float py = static_cast<float>(y + ky - 1) + dy, px = static_cast<float>(x + kx - 1) + dx;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (py > -1.f && py < static_cast<float>(H) && px > -1.f && px < static_cast<float>(W)) {
            const float fy = floorf(py), fx = floorf(px);
            const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
            const float ly = py - fy, lx = px - fx, hy = 1.f - ly, hx = 1.f - lx;
            const float wts[4] = {hy * hx, hy * lx, ly * hx, ly * lx};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int yy = y0 + (k >> 1), xx = x0 + (k & 1);
                if (yy < 0 || yy > H - 1 || xx < 0 || xx > W - 1) continue;
                float v[8];
                load8_split(base + (static_cast<long long>(yy) * W + xx) * ldi, lo, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = fmaf(wts[k], v[i], acc[i]);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] *= mask;
        }
        store8_split(dst + tap * C, 9LL * C, acc);
    }
}

int op_dcn_im2col(Engine* e, const Tensor& in, const float* om, __half* col, const char* layer) {
    if (in.lo > 0) {
        if (in.C % 8) return set_err(e, DV_ERR_UNSUPPORTED, "dcn_im2col: C %% 8 != 0");
        const long long total = static_cast<long long>(in.N) * in.H * in.W * (in.C / 8);
        const double px = static_cast<double>(in.N) * in.H * in.W;
        e->launch_begin("k_dcn_im2col", layer, px * 9 * in.C * 9.0, px * (in.C * 4.0 + 128.0 + 36.0 * in.C));
        k_dcn_im2col_split<<<grid_for(total, 256), 256, 0, e->stream>>>(in.p, in.N, in.H, in.W, in.C, in.ldc(), in.lo, om, col);
        e->launch_end();
        DV_CUDA(e, cudaGetLastError());
        return 0;
    }
    const int cv = in.C >> 3;
    int lcv = 0;
    while ((1 << lcv) < cv) ++lcv;
    if ((in.C % 8) || (1 << lcv) != cv) return set_err(e, DV_ERR_UNSUPPORTED, "dcn_im2col: C must be 8 * 2^k");
    if (static_cast<long long>(in.N) * in.H * in.W * 9 * in.C > 0x7fffffffLL * 4 || static_cast<long long>(in.N) * in.H > 65535)
        return set_err(e, DV_ERR_UNSUPPORTED, "dcn_im2col: tensor too large (N*H <= 65535)");
    const double px = static_cast<double>(in.N) * in.H * in.W;
    e->launch_begin("k_dcn_im2col", layer, px * 9 * in.C * 9.0, px * (in.C * 2.0 + 128.0 + 18.0 * in.C));
    k_dcn_im2col<<<dim3((in.W * cv + 255) / 256, in.N * in.H), 256, 0, e->stream>>>(in.p, in.H, in.W, lcv, in.ldc(), om, col);
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------- IDAUp upsample + add
// out[n, oy, ox, c] = skip[n, oy, ox, c] + sum_{iy, ix} in[n, iy, ix, c] * w[ky][kx][c],  ky = oy + f/2 - iy*f in [0, 2f)
// (ConvTranspose2d(o, o, 2f, stride=f, padding=f//2, groups=o)); w is fp32 [2f][2f][C].
__global__ void __launch_bounds__(256)
k_up_dw_add(const __half* __restrict__ in, int N, int h, int w, int C, int ldi, const float* __restrict__ wt, int f,
            const __half* __restrict__ skip, int lds, __half* __restrict__ out, int ldo) {
    const int cv = C >> 3, Ho = h * f, Wo = w * f, k2 = 2 * f, pad = f >> 1;
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
    float acc[8];
    if (skip != nullptr) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(skip + ((static_cast<long long>(n) * Ho + oy) * Wo + ox) * lds + c8 * 8));
        const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 v = __half22float2(hh[i]);
            acc[2 * i] = v.x;
            acc[2 * i + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    }
    const int iy1 = (oy + pad) / f, ix1 = (ox + pad) / f;
    for (int a = 0; a < 2; ++a) {
        const int iy = iy1 - a, ky = oy + pad - iy * f;
        if (iy < 0 || iy >= h || ky >= k2) continue;
        for (int b = 0; b < 2; ++b) {
            const int ix = ix1 - b, kx = ox + pad - ix * f;
            if (ix < 0 || ix >= w || kx >= k2) continue;
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(in + ((static_cast<long long>(n) * h + iy) * w + ix) * ldi + c8 * 8));
            const float4* wp = reinterpret_cast<const float4*>(wt + (static_cast<long long>(ky) * k2 + kx) * C + c8 * 8);
            const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
            const __half2* hh = reinterpret_cast<const __half2*>(&u);
            const float2 v0 = __half22float2(hh[0]), v1 = __half22float2(hh[1]), v2 = __half22float2(hh[2]), v3 = __half22float2(hh[3]);
            acc[0] = fmaf(v0.x, w0.x, acc[0]);
            acc[1] = fmaf(v0.y, w0.y, acc[1]);
            acc[2] = fmaf(v1.x, w0.z, acc[2]);
            acc[3] = fmaf(v1.y, w0.w, acc[3]);
            acc[4] = fmaf(v2.x, w1.x, acc[4]);
            acc[5] = fmaf(v2.y, w1.y, acc[5]);
            acc[6] = fmaf(v3.x, w1.z, acc[6]);
            acc[7] = fmaf(v3.y, w1.w, acc[7]);
        }
    }
    uint4 o;
    __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) ho[i] = __floats2half2_rn(acc[2 * i], acc[2 * i + 1]);
    *reinterpret_cast<uint4*>(out + ((static_cast<long long>(n) * Ho + oy) * Wo + ox) * ldo + c8 * 8) = o;
}

// Same operation, one (output-phase class, channel group) per thread: an output pixel (m f + py, x f + px) always uses the same 2 x 2
// filter taps of its class, so the class's 4 x 8 weights stay in registers while the thread walks kUpPix input-grid positions.
// The per-pixel kernel above spent 8 of its 14 memory instructions re-fetching those weights (2.0 TB/s of 6.5).
constexpr int kUpPix = 4;
__global__ void __launch_bounds__(256)
k_up_dw_add_cls(const __half* __restrict__ in, int N, int h, int w, int C, int ldi, const float* __restrict__ wt, int f,
                const __half* __restrict__ skip, int lds, __half* __restrict__ out, int ldo) {
    const int cv = C >> 3, lanes = 256 / cv, k2 = 2 * f, pad = f >> 1, Wo = w * f, Ho = h * f;
    const int c8 = threadIdx.x % cv, lane = threadIdx.x / cv;
    const int py = blockIdx.y / f, px = blockIdx.y - py * f;
    const int dy = py >= pad ? 1 : 0, dx = px >= pad ? 1 : 0;
    const int ky0 = (py + pad) - dy * f, kx0 = (px + pad) - dx * f;  // taps (ky0 + a f, kx0 + b f) meet inputs (m + dy - a, x + dx - b)
    float wr[2][2][8];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const float4* wp = reinterpret_cast<const float4*>(wt + (static_cast<long long>(ky0 + a * f) * k2 + kx0 + b * f) * C + c8 * 8);
            const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
            wr[a][b][0] = w0.x, wr[a][b][1] = w0.y, wr[a][b][2] = w0.z, wr[a][b][3] = w0.w;
            wr[a][b][4] = w1.x, wr[a][b][5] = w1.y, wr[a][b][6] = w1.z, wr[a][b][7] = w1.w;
        }
    const long long total = static_cast<long long>(N) * h * w;
    if (lane >= lanes) return;
#pragma unroll 2
    for (int it = 0; it < kUpPix; ++it) {
        const long long p = (static_cast<long long>(blockIdx.x) * kUpPix + it) * lanes + lane;
        if (p >= total) return;
        unsigned t = static_cast<unsigned>(p);  // 32-bit index decode
        const int x = static_cast<int>(t % w);
        t /= w;
        const int m = static_cast<int>(t % h), n = static_cast<int>(t / h);
        const long long opix = (static_cast<long long>(n) * Ho + m * f + py) * Wo + x * f + px;
        // all five loads are issued before the first use (ncu: with the bounds checks as branches every load was waited for on
        // its own); a tap outside the image reads a clamped address and gets a zero weight (fma(v, 0, acc) == acc)
        uint4 us = make_uint4(0u, 0u, 0u, 0u), ui[2][2];
        float ok[2][2];
        if (skip != nullptr) us = __ldg(reinterpret_cast<const uint4*>(skip + opix * lds + c8 * 8));
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int iy = m + dy - a, ix = x + dx - b;
                ok[a][b] = (iy >= 0 && iy < h && ix >= 0 && ix < w) ? 1.f : 0.f;
                const int cy = min(max(iy, 0), h - 1), cx = min(max(ix, 0), w - 1);
                ui[a][b] = __ldg(reinterpret_cast<const uint4*>(in + ((static_cast<long long>(n) * h + cy) * w + cx) * ldi + c8 * 8));
            }
        float acc[8];
        {
            const __half2* hh = reinterpret_cast<const __half2*>(&us);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 v = __half22float2(hh[i]);
                acc[2 * i] = v.x;
                acc[2 * i + 1] = v.y;
            }
        }
        // same accumulation order as k_up_dw_add: a = 0, 1 over b = 0, 1
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const __half2* hh = reinterpret_cast<const __half2*>(&ui[a][b]);
                const float g = ok[a][b];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 v = __half22float2(hh[i]);
                    acc[2 * i] = fmaf(v.x, wr[a][b][2 * i] * g, acc[2 * i]);
                    acc[2 * i + 1] = fmaf(v.y, wr[a][b][2 * i + 1] * g, acc[2 * i + 1]);
                }
            }
        uint4 o;
        __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int i = 0; i < 4; ++i) ho[i] = __floats2half2_rn(acc[2 * i], acc[2 * i + 1]);
        *reinterpret_cast<uint4*>(out + opix * ldo + c8 * 8) = o;
    }
}

__global__ void __launch_bounds__(256)
k_up_dw_add_split(const __half* __restrict__ in, int N, int h, int w, int C, int ldi, long long loi, const float* __restrict__ wt, int f,
                  const __half* __restrict__ skip, int lds, long long los, __half* __restrict__ out, int ldo, long long loo) {
    const int cv = C >> 3, Ho = h * f, Wo = w * f, k2 = 2 * f, pad = f >> 1;
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= static_cast<long long>(N) * Ho * Wo * cv) return;
    unsigned t = static_cast<unsigned>(idx);  // 32-bit index decode (every launch covers < 2^32 work items): a 64-bit division is ~100 SASS instructions
    const int c8 = static_cast<int>(t % cv);
    t /= cv;
    const int ox = static_cast<int>(t % Wo);
    t /= Wo;
    const int oy = static_cast<int>(t % Ho);
    const int n = static_cast<int>(t / Ho);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (skip != nullptr) load8_split(skip + ((static_cast<long long>(n) * Ho + oy) * Wo + ox) * lds + c8 * 8, los, acc);
    const int iy1 = (oy + pad) / f, ix1 = (ox + pad) / f;
    for (int a = 0; a < 2; ++a) {
        const int iy = iy1 - a, ky = oy + pad - iy * f;
        if (iy < 0 || iy >= h || ky >= k2) continue;
        for (int b = 0; b < 2; ++b) {
            const int ix = ix1 - b, kx = ox + pad - ix * f;
            if (ix < 0 || ix >= w || kx >= k2) continue;
            float v[8];
            load8_split(in + ((static_cast<long long>(n) * h + iy) * w + ix) * ldi + c8 * 8, loi, v);
            const float* wp = wt + (static_cast<long long>(ky) * k2 + kx) * C + c8 * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(v[i], __ldg(wp + i), acc[i]);
        }
    }
    store8_split(out + ((static_cast<long long>(n) * Ho + oy) * Wo + ox) * ldo + c8 * 8, loo, acc);
}

// skip.p == nullptr: plain up-sampling (CenterNet's IDAUp concatenates instead of adding)
int op_up_dw_add(Engine* e, const Tensor& in, const float* wt, int f, const Tensor& skip, const Tensor& out, const char* layer) {
    if ((in.C % 8) || out.C != in.C || out.H != in.H * f || out.W != in.W * f || (f != 2 && f != 4 && f != 8) ||
        (skip.p != nullptr && (skip.C != in.C || skip.H != out.H || skip.W != out.W)))
        return set_err(e, DV_ERR_UNSUPPORTED, "up_dw_add: bad shapes");
    const long long total = static_cast<long long>(out.N) * out.H * out.W * (in.C / 8);
    if (in.lo > 0) {
        if (out.lo <= 0 || (skip.p != nullptr && skip.lo <= 0)) return set_err(e, DV_ERR_UNSUPPORTED, "up_dw_add: mixed split / plain tensors");
        e->launch_begin("k_up_dw_add", layer, 8.0 * (double)out.elems(), 4.0 * ((double)in.elems() + 2.0 * (double)out.elems()));
        k_up_dw_add_split<<<grid_for(total, 256), 256, 0, e->stream>>>(in.p, in.N, in.H, in.W, in.C, in.ldc(), in.lo, wt, f, skip.p,
                                                                     skip.p ? skip.ldc() : 0, skip.lo, out.p, out.ldc(), out.lo);
        e->launch_end();
        DV_CUDA(e, cudaGetLastError());
        return 0;
    }
    e->launch_begin("k_up_dw_add", layer, 8.0 * (double)out.elems(), 2.0 * ((double)in.elems() + 2.0 * (double)out.elems()));
    static const bool cls_env = !(getenv("DV_UPCLS") && atoi(getenv("DV_UPCLS")) == 0);
    const int cv = in.C / 8;
    if (cls_env && cv <= 256 && (256 % cv) == 0) {
        const long long per_cta = static_cast<long long>(256 / cv) * kUpPix;
        const long long ctas = (static_cast<long long>(in.N) * in.H * in.W + per_cta - 1) / per_cta;
        k_up_dw_add_cls<<<dim3(static_cast<unsigned>(ctas), f * f), 256, 0, e->stream>>>(in.p, in.N, in.H, in.W, in.C, in.ldc(), wt, f, skip.p,
                                                                                      skip.p ? skip.ldc() : 0, out.p, out.ldc());
    } else {
        k_up_dw_add<<<grid_for(total, 256), 256, 0, e->stream>>>(in.p, in.N, in.H, in.W, in.C, in.ldc(), wt, f, skip.p, skip.p ? skip.ldc() : 0, out.p,
                                                               out.ldc());
    }
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------- slice copy
__global__ void __launch_bounds__(256)
k_copy_slice(const __half* __restrict__ in, long long pixels, int C, int ldi, __half* __restrict__ out, int ldo) {
    const int cv = C >> 3;
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= pixels * cv) return;
    const unsigned i32 = static_cast<unsigned>(idx);  // 32-bit index decode
    const long long px = i32 / cv;
    const int c8 = static_cast<int>(i32 % cv);
    *reinterpret_cast<uint4*>(out + px * ldo + c8 * 8) = __ldg(reinterpret_cast<const uint4*>(in + px * ldi + c8 * 8));
}

int op_copy_slice(Engine* e, const Tensor& in, const Tensor& out) {
    if ((in.C % 8) || out.C != in.C || out.H != in.H || out.W != in.W || out.N != in.N) return set_err(e, DV_ERR_UNSUPPORTED, "copy_slice: bad shapes");
    const long long pixels = static_cast<long long>(in.N) * in.H * in.W;
    if ((in.lo > 0) != (out.lo > 0)) return set_err(e, DV_ERR_UNSUPPORTED, "copy_slice: mixed split / plain tensors");
    e->launch_begin("k_copy_slice", "concat", 0.0, pixels * in.C * 4.0);
    k_copy_slice<<<grid_for(pixels * (in.C / 8), 256), 256, 0, e->stream>>>(in.p, pixels, in.C, in.ldc(), out.p, out.ldc());
    if (in.lo > 0) k_copy_slice<<<grid_for(pixels * (in.C / 8), 256), 256, 0, e->stream>>>(in.p + in.lo, pixels, in.C, in.ldc(), out.p + out.lo, out.ldc());
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------- sigmoid on leading columns
__global__ void __launch_bounds__(256) k_sigmoid_cols(float* __restrict__ maps, long long rows, int ld, int ncols) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= rows * ncols) return;
    float* p = maps + (idx / ncols) * ld + (idx % ncols);
    *p = 1.f / (1.f + expf(-*p));
}

int op_sigmoid_cols(Engine* e, float* maps, long long rows, int ld, int ncols) {
    e->launch_begin("k_sigmoid_cols", "hm.sigmoid", 0.0, rows * ncols * 8.0);
    k_sigmoid_cols<<<grid_for(rows * ncols, 256), 256, 0, e->stream>>>(maps, rows, ld, ncols);
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------- sparse `ax` / `cr` heads
// counts[n] selected cells per image -> offsets (exclusive prefix), total rows, 4 x total (the `cr` GEMM's row count)
__global__ void k_cell_offsets(const int32_t* __restrict__ counts, int N, int cap, int32_t* __restrict__ offsets,
                               int32_t* __restrict__ totals, int32_t* __restrict__ overflow) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int acc = 0;
    for (int n = 0; n < N; ++n) {
        offsets[n] = acc;
        acc += counts[n];
    }
    offsets[N] = acc;
    if (acc > cap) {
        *overflow = acc;
        acc = cap;
    }
    totals[0] = acc;
    totals[1] = 4 * acc;
}

// One CTA per (cell j, image n): rows of the two patch matrices.  feat: NHWC fp16 [N,H,W,C];
// col_ax[row][tap*C + c] (row = offsets[n] + j), col_cr[4*row + k][...]; out-of-image taps are zero (conv padding 1).
__global__ void __launch_bounds__(128)
k_gather_patch3x3(const __half* __restrict__ feat, int H, int W, int C, int K, int cap, const int32_t* __restrict__ counts,
                  const int32_t* __restrict__ offsets, const int32_t* __restrict__ ax_idx, const int32_t* __restrict__ cr_idx,
                  __half* __restrict__ col_ax, __half* __restrict__ col_cr, int ldf, long long lof) {
    const int j = blockIdx.x, n = blockIdx.y;
    if (j >= counts[n]) return;
    const int row = offsets[n] + j;
    if (row >= cap) return;
    const int cv = C >> 3;
    const size_t o = static_cast<size_t>(n) * K + j;
    for (int t = threadIdx.x; t < 5 * 9 * cv; t += blockDim.x) {
        const int c8 = t % cv;
        const int tap = (t / cv) % 9;
        const int pt = t / (9 * cv);
        const int pix = pt == 0 ? ax_idx[o] : cr_idx[o * 4 + pt - 1];
        const int y = pix / W + tap / 3 - 1, x = pix % W + tap % 3 - 1;
        uint4 u = make_uint4(0, 0, 0, 0), ul = make_uint4(0, 0, 0, 0);
        const bool in_img = y >= 0 && y < H && x >= 0 && x < W;
        const __half* src = feat + ((static_cast<long long>(n) * H + y) * W + x) * ldf + c8 * 8;
        if (in_img) u = __ldg(reinterpret_cast<const uint4*>(src));
        // fp32x: the feature map is a split pair and the patch rows are [hi(9C) | lo(9C)]
        const int rw = lof > 0 ? 18 * C : 9 * C;
        __half* dst = pt == 0 ? col_ax + static_cast<long long>(row) * rw : col_cr + (4LL * row + pt - 1) * rw;
        *reinterpret_cast<uint4*>(dst + tap * C + c8 * 8) = u;
        if (lof > 0) {
            if (in_img) ul = __ldg(reinterpret_cast<const uint4*>(src + lof));
            *reinterpret_cast<uint4*>(dst + 9 * C + tap * C + c8 * 8) = ul;
        }
    }
}

// The ResNet-18 detector's `ax` / `cr` heads end in a 1x1 conv over their own 64-channel hidden maps: one CTA per (cell j,
// image n) copies the hidden pixel of the centre (from fa) and of the four corners (from fc) into the GEMM rows.
__global__ void __launch_bounds__(64)
k_gather_pix(const __half* __restrict__ fa, const __half* __restrict__ fc, int H, int W, int C, int K, int cap,
             const int32_t* __restrict__ counts, const int32_t* __restrict__ offsets, const int32_t* __restrict__ ax_idx,
             const int32_t* __restrict__ cr_idx, __half* __restrict__ col_ax, __half* __restrict__ col_cr, int ldf, long long lof) {
    const int j = blockIdx.x, n = blockIdx.y;
    if (j >= counts[n]) return;
    const int row = offsets[n] + j;
    if (row >= cap) return;
    const int cv = C >> 3;
    const size_t o = static_cast<size_t>(n) * K + j;
    const int rw = lof > 0 ? 2 * C : C;  // fp32x rows are [hi(C) | lo(C)]
    for (int t = threadIdx.x; t < 5 * cv; t += blockDim.x) {
        const int c8 = t % cv, pt = t / cv;
        const int pix = pt == 0 ? ax_idx[o] : cr_idx[o * 4 + pt - 1];
        const __half* src = (pt == 0 ? fa : fc) + (static_cast<long long>(n) * H * W + pix) * ldf + c8 * 8;
        __half* dst = pt == 0 ? col_ax + static_cast<long long>(row) * rw : col_cr + (4LL * row + pt - 1) * rw;
        *reinterpret_cast<uint4*>(dst + c8 * 8) = __ldg(reinterpret_cast<const uint4*>(src));
        if (lof > 0) *reinterpret_cast<uint4*>(dst + C + c8 * 8) = __ldg(reinterpret_cast<const uint4*>(src + lof));
    }
}

__global__ void __launch_bounds__(256)
k_logi_combine(const float* __restrict__ ax, const float* __restrict__ cr, int C, const int32_t* __restrict__ totals,
               float* __restrict__ out) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long row = idx / C;
    if (row >= totals[0]) return;
    const int c = static_cast<int>(idx % C);
    const float* r = cr + 4 * row * C + c;
    float s = __fadd_rn(0.f, r[0]);
    s = __fadd_rn(s, r[C]);
    s = __fadd_rn(s, r[2 * C]);
    s = __fadd_rn(s, r[3 * C]);
    out[idx] = __fadd_rn(ax[idx], s);
}

int op_cell_offsets(Engine* e, const int32_t* counts, int N, int cap, int32_t* offsets, int32_t* totals, int32_t* overflow) {
    e->launch_begin("k_cell_offsets", "lore_feat", 0.0, N * 8.0);
    k_cell_offsets<<<1, 32, 0, e->stream>>>(counts, N, cap, offsets, totals, overflow);
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

int op_gather_patch3x3(Engine* e, const Tensor& feat, int K, int cap, const int32_t* counts, const int32_t* offsets,
                       const int32_t* ax_idx, const int32_t* cr_idx, __half* col_ax, __half* col_cr) {
    if ((feat.C % 8) || (feat.ld != 0 && feat.lo == 0)) return set_err(e, DV_ERR_UNSUPPORTED, "gather_patch3x3: dense C %% 8 == 0 input");
    e->launch_begin("k_gather_patch3x3", "lore_feat", 0.0, (double)cap * 5 * 9 * feat.C * (feat.lo ? 8.0 : 4.0));
    k_gather_patch3x3<<<dim3(K, feat.N), 128, 0, e->stream>>>(feat.p, feat.H, feat.W, feat.C, K, cap, counts, offsets, ax_idx, cr_idx,
                                                              col_ax, col_cr, feat.ldc(), feat.lo);
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

int op_gather_pix(Engine* e, const Tensor& fa, const Tensor& fc, int K, int cap, const int32_t* counts, const int32_t* offsets,
                  const int32_t* ax_idx, const int32_t* cr_idx, __half* col_ax, __half* col_cr) {
    if ((fa.C % 8) || fa.C != fc.C || fa.ldc() != fc.ldc() || fa.lo != fc.lo || fa.H != fc.H || fa.W != fc.W)
        return set_err(e, DV_ERR_UNSUPPORTED, "gather_pix: the two hidden maps must share one layout, C %% 8 == 0");
    e->launch_begin("k_gather_pix", "lore_feat", 0.0, (double)cap * 5 * fa.C * (fa.lo ? 8.0 : 4.0));
    k_gather_pix<<<dim3(K, fa.N), 64, 0, e->stream>>>(fa.p, fc.p, fa.H, fa.W, fa.C, K, cap, counts, offsets, ax_idx, cr_idx, col_ax,
                                                      col_cr, fa.ldc(), fa.lo);
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

int op_logi_combine(Engine* e, const float* ax, const float* cr, int C, int cap, const int32_t* totals, float* out) {
    e->launch_begin("k_logi_combine", "lore_feat", 0.0, (double)cap * C * 24.0);
    k_logi_combine<<<grid_for(static_cast<long long>(cap) * C, 256), 256, 0, e->stream>>>(ax, cr, C, totals, out);
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

}  // namespace dv
