// det -> rec glue on the GPU: the perspective crop of OcrCommonUtils.crop_image
// (utils/ocr/ocr_common_utils.py:214-262) = cv2.warpPerspective(img, T, (w, h)) with its defaults (INTER_LINEAR,
// BORDER_CONSTANT 0) on a uint8 HWC page, restated bit for bit from OpenCV's WarpPerspectiveInvoker + remapBilinear
// (imgproc/src/imgwarp.cpp; checked against cv2 4.13 on 2.5 M pixels of 300 random quads, tools/warp_restatement.py):
//   * the caller passes M = cv2.invert(T) (the host keeps getPerspectiveTransform / invert, 9 doubles per crop);
//   * destination columns are walked in blocks of bw0 = min(1024 / min(16, h), w) columns: X0 = M0*xb + M1*y + M2 at the
//     block start, then (X0 + M0*x1) * (32 / (W0 + M6*x1)) in double with NO contraction, clamped to the int range and
//     rounded half-to-even (cvRound) -> 1/32-pixel fixed point;  sx = X >> 5 (saturated to int16), ax = X & 31;
//   * weights (32-ax)(32-ay)*32 ... (the BilinearTab_i entries, exact integers summing to 1 << 15); neighbours outside
//     the page contribute the border value 0;  dst = (sum + (1 << 14)) >> 15.
// One thread per destination pixel (three channels); crops are packed back to back in one output buffer.
#include "engine.h"

namespace dv {

namespace {

__global__ void __launch_bounds__(256)
k_warp_perspective_u8(const uint8_t* __restrict__ img, int H, int W, const double* __restrict__ minv /*[n][9]*/,
                      const int32_t* __restrict__ sizes /*[n][2] = (w, h)*/, const long long* __restrict__ offsets /*[n]*/,
                      uint8_t* __restrict__ out) {
    const int crop = blockIdx.y;
    const int w = sizes[2 * crop], h = sizes[2 * crop + 1];
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= w * h) return;
    const int y = idx / w, x = idx - y * w;
    const double* M = minv + 9 * crop;
    const int bh0 = min(16, h);
    const int bw0 = min(1024 / bh0, w);
    const int xb = x / bw0 * bw0, x1 = x - xb;
    const double dxb = static_cast<double>(xb), dy = static_cast<double>(y), dx1 = static_cast<double>(x1);
    const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(M[0], dxb), __dmul_rn(M[1], dy)), M[2]);
    const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(M[3], dxb), __dmul_rn(M[4], dy)), M[5]);
    const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(M[6], dxb), __dmul_rn(M[7], dy)), M[8]);
    double Wd = __dadd_rn(W0, __dmul_rn(M[6], dx1));
    Wd = Wd != 0.0 ? __ddiv_rn(32.0, Wd) : 0.0;
    const double fX = fmax(-2147483648.0, fmin(2147483647.0, __dmul_rn(__dadd_rn(X0, __dmul_rn(M[0], dx1)), Wd)));
    const double fY = fmax(-2147483648.0, fmin(2147483647.0, __dmul_rn(__dadd_rn(Y0, __dmul_rn(M[3], dx1)), Wd)));
    const int X = __double2int_rn(fX), Y = __double2int_rn(fY);  // round half to even = cvRound
    const int sx = max(-32768, min(32767, X >> 5)), sy = max(-32768, min(32767, Y >> 5));
    const int ax = X & 31, ay = Y & 31;
    const int w00 = (32 - ax) * (32 - ay) * 32, w01 = ax * (32 - ay) * 32, w10 = (32 - ax) * ay * 32, w11 = ax * ay * 32;
    int acc[3] = {0, 0, 0};
    auto tap = [&](int yy, int xx, int wt) {
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
            const uint8_t* p = img + (static_cast<long long>(yy) * W + xx) * 3;
            acc[0] += p[0] * wt;
            acc[1] += p[1] * wt;
            acc[2] += p[2] * wt;
        }
    };
    tap(sy, sx, w00);
    tap(sy, sx + 1, w01);
    tap(sy + 1, sx, w10);
    tap(sy + 1, sx + 1, w11);
    uint8_t* o = out + offsets[crop] + static_cast<long long>(idx) * 3;
    o[0] = static_cast<uint8_t>((acc[0] + (1 << 14)) >> 15);
    o[1] = static_cast<uint8_t>((acc[1] + (1 << 14)) >> 15);
    o[2] = static_cast<uint8_t>((acc[2] + (1 << 14)) >> 15);
}

// cv2.resize(src, (dw, dh)) with its default INTER_LINEAR on uint8 HWC, bit for bit (OpenCV resize.cpp: resizeGeneric_ with
// HResizeLinear<uchar,int,short,2048> and the VResizeLinear<uchar,int,short,FixedPtCast<22>> specialisation; checked against
// cv2 4.13 on 26 M elements, tools/resize_restatement.py):
//   scale = 1.0 / (dst / (double) src);  f = (float)((d + 0.5) * scale - 0.5);  s = floor(f);  f -= s
//   columns: s < 0 -> (s, f) = (0, 0);  s >= sw - 1 -> (sw - 1, 0)          rows: f kept, the two row indices are clipped
//   a = cvRound((1 - f) * 2048), cvRound(f * 2048)  (round half to even)
//   H(row, dx) = p[s] * a0 + p[min(s + 1, sw - 1)] * a1
//   dst = (((b0 * (H(r0) >> 4)) >> 16) + ((b1 * (H(r1) >> 4)) >> 16) + 2) >> 2
// and the one special case cv2 makes for INTER_LINEAR: an exact 2x reduction in both directions is the 2x2 box average
// (a + b + c + d + 2) >> 2 (INTER_AREA fast path).  Crops come packed (the warp kernel's output); crop i is written to
// out[i] = [dst_h, dst_w_pad, 3] with columns >= dst_widths[i] zero (the recogniser's zero padding).
__global__ void __launch_bounds__(256)
k_resize_linear_u8(const uint8_t* __restrict__ src, const long long* __restrict__ src_off, const int32_t* __restrict__ src_sizes,
                   const int32_t* __restrict__ dst_widths, int dst_h, int dst_w_pad, uint8_t* __restrict__ out) {
    const int crop = blockIdx.y;
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= dst_h * dst_w_pad) return;
    const int dy = idx / dst_w_pad, dx = idx - dy * dst_w_pad;
    uint8_t* o = out + (static_cast<long long>(crop) * dst_h * dst_w_pad + idx) * 3;
    const int dw = dst_widths[crop];
    if (dx >= dw) {
        o[0] = o[1] = o[2] = 0;
        return;
    }
    const int sw = src_sizes[2 * crop], sh = src_sizes[2 * crop + 1];
    const uint8_t* s = src + src_off[crop];
    if (sw == 2 * dw && sh == 2 * dst_h) {
        const uint8_t* p = s + (static_cast<long long>(2 * dy) * sw + 2 * dx) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) o[c] = static_cast<uint8_t>((p[c] + p[3 + c] + p[sw * 3 + c] + p[sw * 3 + 3 + c] + 2) >> 2);
        return;
    }
    const double scale_x = __ddiv_rn(1.0, __ddiv_rn(static_cast<double>(dw), static_cast<double>(sw)));
    const double scale_y = __ddiv_rn(1.0, __ddiv_rn(static_cast<double>(dst_h), static_cast<double>(sh)));
    float fx = static_cast<float>(__dsub_rn(__dmul_rn(__dadd_rn(static_cast<double>(dx), 0.5), scale_x), 0.5));
    int sx = static_cast<int>(floorf(fx));
    fx = __fsub_rn(fx, static_cast<float>(sx));
    if (sx < 0) {
        sx = 0;
        fx = 0.f;
    }
    if (sx >= sw - 1) {
        sx = sw - 1;
        fx = 0.f;
    }
    float fy = static_cast<float>(__dsub_rn(__dmul_rn(__dadd_rn(static_cast<double>(dy), 0.5), scale_y), 0.5));
    const int sy = static_cast<int>(floorf(fy));
    fy = __fsub_rn(fy, static_cast<float>(sy));
    const int a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, fx), 2048.f)), a1 = __float2int_rn(__fmul_rn(fx, 2048.f));
    const int b0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, fy), 2048.f)), b1 = __float2int_rn(__fmul_rn(fy, 2048.f));
    const int sx1 = min(sx + 1, sw - 1);
    const int r0 = min(max(sy, 0), sh - 1), r1 = min(max(sy + 1, 0), sh - 1);
    const uint8_t* p0 = s + static_cast<long long>(r0) * sw * 3;
    const uint8_t* p1 = s + static_cast<long long>(r1) * sw * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int h0 = p0[sx * 3 + c] * a0 + p0[sx1 * 3 + c] * a1;
        const int h1 = p1[sx * 3 + c] * a0 + p1[sx1 * 3 + c] * a1;
        const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
        o[c] = static_cast<uint8_t>(min(max(v, 0), 255));
    }
}

}  // namespace

int op_resize_linear_u8(Engine* e, const uint8_t* src, const long long* src_off, const int32_t* src_sizes, const int32_t* dst_widths,
                        int n, int dst_h, int dst_w_pad, uint8_t* out) {
    if (n <= 0) return 0;
    if (n > 65535) return set_err(e, DV_ERR_UNSUPPORTED, "resize_linear: more than 65535 crops per call");
    const int px = dst_h * dst_w_pad;
    e->launch_begin("k_resize_linear_u8", "crop", 0.0, static_cast<double>(n) * px * 3.0 * 5.0);
    k_resize_linear_u8<<<dim3((px + 255) / 256, n), 256, 0, e->stream>>>(src, src_off, src_sizes, dst_widths, dst_h, dst_w_pad, out);
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

int op_warp_perspective_u8(Engine* e, const uint8_t* img, int H, int W, const double* minv, const int32_t* sizes,
                           const long long* offsets, int n, int max_pixels, uint8_t* out) {
    if (n <= 0) return 0;
    if (n > 65535) return set_err(e, DV_ERR_UNSUPPORTED, "warp_perspective: more than 65535 crops per call");
    e->launch_begin("k_warp_perspective_u8", "crop", 0.0, static_cast<double>(n) * max_pixels * 3.0 * 5.0);
    k_warp_perspective_u8<<<dim3((max_pixels + 255) / 256, n), 256, 0, e->stream>>>(img, H, W, minv, sizes, offsets, out);
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

}  // namespace dv
