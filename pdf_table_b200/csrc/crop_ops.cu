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

// ---------------------------------------------------------------------------------------------------------------------
// The whole det -> rec glue without the host: (1) k_quad_homography = the host part of OcrCommonUtils.crop_image
// (utils/ocr/ocr_common_utils.py:227-257) + cv2.getPerspectiveTransform + cv2.invert per quad, one thread each, in double
// with explicit _rn intrinsics (no contraction); (2) k_crop_resize_fused = warp + keep-ratio resize in one pass, the resize
// taps evaluating the warp on the fly, so no variable-size intermediate crop exists and nothing about the launch depends on
// data the host has not seen.
//   cv2.getPerspectiveTransform (OpenCV 4.13, imgwarp.cpp) restated: the 8x8 system with rows (x_i, y_i, 1, 0, 0, 0,
//   -x_i X_i, -y_i X_i | X_i) and (0, 0, 0, x_i, y_i, 1, -x_i Y_i, -y_i Y_i | Y_i), where the four products are formed in
//   FLOAT (Point2f arithmetic) before the conversion to double -- this is what made the textbook double system differ from
//   cv2 by 1e-9 --, solved by cv::hal::LU64f: partial pivoting on |a|, d = -1 / pivot, row_j += (a_ji d) row_i, then back
//   substitution s -= a_ik x_k, x_i = s / a_ii.  cv2.invert of the 3x3: the closed-form cofactor path.  Both checked bit for
//   bit against cv2 (tools/homography_restatement.py: 500 / 500 and 3000 / 3000 random quads).
__device__ __forceinline__ double dfms(double a, double b, double c, double d) {  // a*b - c*d, two roundings + one
    return __dsub_rn(__dmul_rn(a, b), __dmul_rn(c, d));
}

// quads: [n][4][2], or -- with box_counts -- the dv_db_boxes output [pages][box_stride][8] read as per_page slots per page
// (slot k of page p = box k if k < box_counts[p], else skipped)
// width_rule 0: OCRRecognitionPreprocessor.keepratio_resize (ConvNextViT): cur_w = dst_w_max if ratio > dst_w_max / dst_h else
//   int(dst_h * ratio).  width_rule 1: PPOcrRecPreProcessor.resize_norm_img for a crop that is its own batch, as the reference's
//   orchestrator calls it (ocr_rec_pp/processor_ocr_rec_pp.py:43-59, one crop per call -- ocr_system_task.py:309-312):
//   resized_w = min(imgW, max(ceil(dst_h * ratio), 16)) with imgW = clamp(int(dst_h * max(ratio, 320 / 48)), 16, 1280); the padded
//   width imgW of each crop is the host's to recompute from `sizes` (predictors.pp_rec_padded_width).
__global__ void k_quad_homography(const float* __restrict__ quads, const int32_t* __restrict__ box_counts, int box_stride, int per_page,
                                  int n, int dst_h, int dst_w_max, int width_rule, double* __restrict__ minv /*[n][9]*/,
                                  int32_t* __restrict__ sizes /*[n][2]*/, int32_t* __restrict__ dst_widths /*[n]*/) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const float* src = quads + static_cast<long long>(q) * 8;
    if (box_counts != nullptr) {
        const int pg = q / per_page, k = q - pg * per_page;
        if (k >= box_counts[pg]) {
            sizes[2 * q] = sizes[2 * q + 1] = 0;
            dst_widths[q] = 0;
            return;
        }
        src = quads + (static_cast<long long>(pg) * box_stride + k) * 8;
    }
    double px[4], py[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        px[i] = static_cast<double>(src[2 * i]);
        py[i] = static_cast<double>(src[2 * i + 1]);
    }
    // corner order: the reference's exchange sort on x, then the left and the right pair on y
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i + 1; j < 4; ++j)
            if (px[i] > px[j]) {
                double t = px[i]; px[i] = px[j]; px[j] = t;
                t = py[i]; py[i] = py[j]; py[j] = t;
            }
    if (py[0] > py[1]) {
        double t = px[0]; px[0] = px[1]; px[1] = t;
        t = py[0]; py[0] = py[1]; py[1] = t;
    }
    if (py[2] > py[3]) {
        double t = px[2]; px[2] = px[3]; px[3] = t;
        t = py[2]; py[2] = py[3]; py[3] = t;
    }
    const double x1 = px[0], y1 = py[0], x2 = px[2], y2 = py[2], x3 = px[3], y3 = py[3], x4 = px[1], y4 = py[1];
    auto dist = [](double xa, double ya, double xb, double yb) {
        const double dx = __dsub_rn(xa, xb), dy = __dsub_rn(ya, yb);
        return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    };
    const double img_w = dist(__ddiv_rn(__dadd_rn(x1, x4), 2.0), __ddiv_rn(__dadd_rn(y1, y4), 2.0), __ddiv_rn(__dadd_rn(x2, x3), 2.0),
                              __ddiv_rn(__dadd_rn(y2, y3), 2.0));
    const double img_h = dist(__ddiv_rn(__dadd_rn(x1, x2), 2.0), __ddiv_rn(__dadd_rn(y1, y2), 2.0), __ddiv_rn(__dadd_rn(x4, x3), 2.0),
                              __ddiv_rn(__dadd_rn(y4, y3), 2.0));
    const int w = static_cast<int>(img_w), h = static_cast<int>(img_h);
    sizes[2 * q] = w;
    sizes[2 * q + 1] = h;
    dst_widths[q] = 0;
    if (w <= 0 || h <= 0) return;
    const double ratio = __ddiv_rn(static_cast<double>(w), static_cast<double>(h));
    int cur_w;
    if (width_rule == 1) {
        const double max_wh = fmax(ratio, __ddiv_rn(320.0, 48.0));
        int img_w = static_cast<int>(__dmul_rn(static_cast<double>(dst_h), max_wh));
        img_w = max(min(img_w, 1280), 16);
        const int ratio_w = max(static_cast<int>(ceil(__dmul_rn(static_cast<double>(dst_h), ratio))), 16);
        cur_w = ratio_w > img_w ? img_w : ratio_w;
        if (cur_w > dst_w_max) cur_w = dst_w_max;  // never past the caller's buffer (dst_w_max >= 1280 keeps the rule exact)
    } else {
        cur_w = ratio > __ddiv_rn(static_cast<double>(dst_w_max), static_cast<double>(dst_h))
                    ? dst_w_max
                    : static_cast<int>(__dmul_rn(static_cast<double>(dst_h), ratio));
    }
    if (cur_w <= 0) return;
    // float32 corner arrays: src = (x1,y1), (x2,y2), (x4,y4), (x3,y3);  dst = (0,0), (W-1,0), (0,H-1), (W-1,H-1)
    const float sx[4] = {static_cast<float>(x1), static_cast<float>(x2), static_cast<float>(x4), static_cast<float>(x3)};
    const float sy[4] = {static_cast<float>(y1), static_cast<float>(y2), static_cast<float>(y4), static_cast<float>(y3)};
    const float tw = static_cast<float>(__dsub_rn(img_w, 1.0)), th = static_cast<float>(__dsub_rn(img_h, 1.0));
    const float dxs[4] = {0.f, tw, 0.f, tw}, dys[4] = {0.f, 0.f, th, th};
    double A[8][8], b[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) A[i][k] = A[i + 4][k] = 0.0;
        A[i][0] = A[i + 4][3] = static_cast<double>(sx[i]);
        A[i][1] = A[i + 4][4] = static_cast<double>(sy[i]);
        A[i][2] = A[i + 4][5] = 1.0;
        A[i][6] = static_cast<double>(__fmul_rn(-sx[i], dxs[i]));  // products in float, as Point2f arithmetic
        A[i][7] = static_cast<double>(__fmul_rn(-sy[i], dxs[i]));
        A[i + 4][6] = static_cast<double>(__fmul_rn(-sx[i], dys[i]));
        A[i + 4][7] = static_cast<double>(__fmul_rn(-sy[i], dys[i]));
        b[i] = static_cast<double>(dxs[i]);
        b[i + 4] = static_cast<double>(dys[i]);
    }
    for (int i = 0; i < 8; ++i) {
        int k = i;
        for (int j = i + 1; j < 8; ++j)
            if (fabs(A[j][i]) > fabs(A[k][i])) k = j;
        if (fabs(A[k][i]) < 2.220446049250313e-16 * 100) return;  // singular (a degenerate quad): cv2 gives no usable transform
        if (k != i) {
            for (int j = i; j < 8; ++j) {
                const double t = A[i][j]; A[i][j] = A[k][j]; A[k][j] = t;
            }
            const double t = b[i]; b[i] = b[k]; b[k] = t;
        }
        const double d = __ddiv_rn(-1.0, A[i][i]);
        for (int j = i + 1; j < 8; ++j) {
            const double alpha = __dmul_rn(A[j][i], d);
            for (int c = i + 1; c < 8; ++c) A[j][c] = __dadd_rn(A[j][c], __dmul_rn(alpha, A[i][c]));
            b[j] = __dadd_rn(b[j], __dmul_rn(alpha, b[i]));
        }
    }
    for (int i = 7; i >= 0; --i) {
        double s = b[i];
        for (int c = i + 1; c < 8; ++c) s = __dsub_rn(s, __dmul_rn(A[i][c], b[c]));
        b[i] = __ddiv_rn(s, A[i][i]);
    }
    // T = [b0 b1 b2; b3 b4 b5; b6 b7 1]; cv2.invert (3x3 closed form)
    const double S00 = b[0], S01 = b[1], S02 = b[2], S10 = b[3], S11 = b[4], S12 = b[5], S20 = b[6], S21 = b[7], S22 = 1.0;
    double det = __dadd_rn(__dsub_rn(__dmul_rn(S00, dfms(S11, S22, S12, S21)), __dmul_rn(S01, dfms(S10, S22, S12, S20))),
                           __dmul_rn(S02, dfms(S10, S21, S11, S20)));
    if (det == 0.0) return;
    det = __ddiv_rn(1.0, det);
    double* M = minv + 9 * q;
    M[0] = __dmul_rn(dfms(S11, S22, S12, S21), det);
    M[1] = __dmul_rn(dfms(S02, S21, S01, S22), det);
    M[2] = __dmul_rn(dfms(S01, S12, S02, S11), det);
    M[3] = __dmul_rn(dfms(S12, S20, S10, S22), det);
    M[4] = __dmul_rn(dfms(S00, S22, S02, S20), det);
    M[5] = __dmul_rn(dfms(S02, S10, S00, S12), det);
    M[6] = __dmul_rn(dfms(S10, S21, S11, S20), det);
    M[7] = __dmul_rn(dfms(S01, S20, S00, S21), det);
    M[8] = __dmul_rn(dfms(S00, S11, S01, S10), det);
    dst_widths[q] = cur_w;
}

// one pixel of cv2.warpPerspective(page, T, (w, h)) -- the body of k_warp_perspective_u8
__device__ __forceinline__ void warp_pixel(const uint8_t* __restrict__ img, int H, int W, const double* __restrict__ M, int w, int h,
                                           int x, int y, int (&v)[3]) {
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
    const int X = __double2int_rn(fX), Y = __double2int_rn(fY);
    const int sx = max(-32768, min(32767, X >> 5)), sy = max(-32768, min(32767, Y >> 5));
    const int ax = X & 31, ay = Y & 31;
    const int wt[4] = {(32 - ax) * (32 - ay) * 32, ax * (32 - ay) * 32, (32 - ax) * ay * 32, ax * ay * 32};
    int acc[3] = {0, 0, 0};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int yy = sy + (t >> 1), xx = sx + (t & 1);
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
            const uint8_t* p = img + (static_cast<long long>(yy) * W + xx) * 3;
            acc[0] += p[0] * wt[t];
            acc[1] += p[1] * wt[t];
            acc[2] += p[2] * wt[t];
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = (acc[c] + (1 << 14)) >> 15;
}

__global__ void __launch_bounds__(256)
k_crop_resize_fused(const uint8_t* __restrict__ pages, int H, int W, const int32_t* __restrict__ page_idx, int per_page,
                    const double* __restrict__ minv,
                    const int32_t* __restrict__ sizes, const int32_t* __restrict__ dst_widths, int dst_h, int dst_w_pad,
                    uint8_t* __restrict__ out) {
    const int crop = blockIdx.y;
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= dst_h * dst_w_pad) return;
    const int dy = idx / dst_w_pad, dx = idx - dy * dst_w_pad;
    uint8_t* o = out + (static_cast<long long>(crop) * dst_h * dst_w_pad + idx) * 3;
    const int dw = dst_widths[crop];
    if (dx >= dw) {  // zero padding; a skipped quad (dw == 0) is an all-zero row block
        o[0] = o[1] = o[2] = 0;
        return;
    }
    const int sw = sizes[2 * crop], sh = sizes[2 * crop + 1];
    const uint8_t* img = pages + static_cast<long long>(page_idx ? page_idx[crop] : (per_page > 0 ? crop / per_page : 0)) * H * W * 3;
    const double* M = minv + 9 * crop;
    int p00[3], p01[3], p10[3], p11[3];
    if (sw == 2 * dw && sh == 2 * dst_h) {
        warp_pixel(img, H, W, M, sw, sh, 2 * dx, 2 * dy, p00);
        warp_pixel(img, H, W, M, sw, sh, 2 * dx + 1, 2 * dy, p01);
        warp_pixel(img, H, W, M, sw, sh, 2 * dx, 2 * dy + 1, p10);
        warp_pixel(img, H, W, M, sw, sh, 2 * dx + 1, 2 * dy + 1, p11);
#pragma unroll
        for (int c = 0; c < 3; ++c) o[c] = static_cast<uint8_t>((p00[c] + p01[c] + p10[c] + p11[c] + 2) >> 2);
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
    warp_pixel(img, H, W, M, sw, sh, sx, r0, p00);
    warp_pixel(img, H, W, M, sw, sh, sx1, r0, p01);
    warp_pixel(img, H, W, M, sw, sh, sx, r1, p10);
    warp_pixel(img, H, W, M, sw, sh, sx1, r1, p11);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int h0 = p00[c] * a0 + p01[c] * a1;
        const int h1 = p10[c] * a0 + p11[c] * a1;
        const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
        o[c] = static_cast<uint8_t>(min(max(v, 0), 255));
    }
}

}  // namespace

// cv2.warpAffine(img, M, (w, h), flags=INTER_LINEAR) (border constant 0) on uint8 HWC, bit for bit (OpenCV imgwarp.cpp
// WarpAffineInvoker: 10-bit fixed-point coordinates X = (cvRound((m1 y + m2) 1024) + 16 + cvRound(m0 x 1024)) >> 5, then the same
// 1/32-pixel bilinear remap as the perspective warp; the numpy restatement oracle/crop_ref.py warp_affine is checked against cv2 4.13 in
// tests/test_crop_cpu.py::test_warp_affine_restatement_equals_cv2, this kernel against cv2 in tests/test_gpu_crop.py).  `m` is
// the INVERTED 2x3 matrix (the host inverts with cv2's own formula, predictors.invert_affine).  Replaces the warp of
// TableLorePreProcessor.process (lore/processer_lore.py:80-91).
struct Affine6 {
    double m[6];
};
__global__ void __launch_bounds__(256)
k_warp_affine_u8(const uint8_t* __restrict__ img, int H, int W, Affine6 a, int w, int h, uint8_t* __restrict__ out) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= w * h) return;
    const int y = idx / w, x = idx - y * w;
    const double dx = static_cast<double>(x), dy = static_cast<double>(y);
    const int adelta = __double2int_rn(__dmul_rn(__dmul_rn(a.m[0], dx), 1024.0));
    const int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(a.m[3], dx), 1024.0));
    const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(a.m[1], dy), a.m[2]), 1024.0)) + 16;
    const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(a.m[4], dy), a.m[5]), 1024.0)) + 16;
    const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
    const int sx = max(-32768, min(32767, X >> 5)), sy = max(-32768, min(32767, Y >> 5));
    const int ax = X & 31, ay = Y & 31;
    const int wt[4] = {(32 - ax) * (32 - ay) * 32, ax * (32 - ay) * 32, (32 - ax) * ay * 32, ax * ay * 32};
    int acc[3] = {0, 0, 0};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int yy = sy + (t >> 1), xx = sx + (t & 1);
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
            const uint8_t* p = img + (static_cast<long long>(yy) * W + xx) * 3;
            acc[0] += p[0] * wt[t];
            acc[1] += p[1] * wt[t];
            acc[2] += p[2] * wt[t];
        }
    }
    uint8_t* o = out + static_cast<long long>(idx) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = static_cast<uint8_t>((acc[c] + (1 << 14)) >> 15);
}

int op_warp_affine_u8(Engine* e, const uint8_t* img, int H, int W, const double* m_inv6, int w, int h, uint8_t* out) {
    Affine6 a;
    for (int i = 0; i < 6; ++i) a.m[i] = m_inv6[i];
    e->launch_begin("k_warp_affine_u8", "pre", 0.0, static_cast<double>(w) * h * 3.0 * 5.0);
    k_warp_affine_u8<<<(w * h + 255) / 256, 256, 0, e->stream>>>(img, H, W, a, w, h, out);
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

// Layout -> table-structure glue: the warp above for MANY axis-aligned sub-rectangles of resident pages in one launch
// (blockIdx.y = table).  Each table is the slice pages[page][y0:y0+ch, x0:x0+cw] (OcrCommonUtils.crop_image_by_box,
// utils/ocr/ocr_common_utils.py:269-284) warped by its own inverted matrix exactly as cv2.warpAffine would warp the cut-out
// crop: the slice is addressed in place through the page pitch, samples outside the slice are the zero border.  A rect that
// does not lie inside its page yields a zero image (the host helper predictors.table_crop_rect never produces one).
__global__ void __launch_bounds__(256)
k_warp_affine_rects_u8(const uint8_t* __restrict__ pages, int n_pages, int H, int W, const int32_t* __restrict__ rects,
                       const double* __restrict__ minv, int w, int h, uint8_t* __restrict__ out) {
    const int k = blockIdx.y;
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= w * h) return;
    const int32_t* r = rects + k * 5;
    const int pg = r[0], x0 = r[1], y0 = r[2], cw = r[3], ch = r[4];
    uint8_t* o = out + (static_cast<long long>(k) * w * h + idx) * 3;
    if (pg < 0 || pg >= n_pages || x0 < 0 || y0 < 0 || cw <= 0 || ch <= 0 || x0 > W - cw || y0 > H - ch) {
        o[0] = o[1] = o[2] = 0;
        return;
    }
    const double* m = minv + k * 6;
    const long long pitch = static_cast<long long>(W) * 3;
    const uint8_t* img = pages + (static_cast<long long>(pg) * H + y0) * pitch + static_cast<long long>(x0) * 3;
    const int y = idx / w, x = idx - y * w;
    const double dx = static_cast<double>(x), dy = static_cast<double>(y);
    const int adelta = __double2int_rn(__dmul_rn(__dmul_rn(m[0], dx), 1024.0));
    const int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(m[3], dx), 1024.0));
    const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[1], dy), m[2]), 1024.0)) + 16;
    const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[4], dy), m[5]), 1024.0)) + 16;
    const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
    const int sx = max(-32768, min(32767, X >> 5)), sy = max(-32768, min(32767, Y >> 5));
    const int ax = X & 31, ay = Y & 31;
    const int wt[4] = {(32 - ax) * (32 - ay) * 32, ax * (32 - ay) * 32, (32 - ax) * ay * 32, ax * ay * 32};
    int acc[3] = {0, 0, 0};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int yy = sy + (t >> 1), xx = sx + (t & 1);
        if (yy >= 0 && yy < ch && xx >= 0 && xx < cw) {
            const uint8_t* p = img + yy * pitch + xx * 3;
            acc[0] += p[0] * wt[t];
            acc[1] += p[1] * wt[t];
            acc[2] += p[2] * wt[t];
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = static_cast<uint8_t>((acc[c] + (1 << 14)) >> 15);
}

int op_warp_affine_rects_u8(Engine* e, const uint8_t* pages, int n_pages, int H, int W, const int32_t* rects, const double* minv,
                            int n, int w, int h, uint8_t* out) {
    if (n == 0) return 0;
    e->launch_begin("k_warp_affine_rects_u8", "pre", 0.0, static_cast<double>(n) * w * h * 3.0 * 5.0);
    k_warp_affine_rects_u8<<<dim3((w * h + 255) / 256, n), 256, 0, e->stream>>>(pages, n_pages, H, W, rects, minv, w, h, out);
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

int op_crop_quads_for_rec(Engine* e, const uint8_t* pages, int H, int W, const float* quads, const int32_t* page_idx,
                          const int32_t* box_counts, int box_stride, int per_page, int n, int dst_h, int dst_w_pad, uint8_t* out,
                          int32_t* dst_widths, double* minv_ws, int32_t* sizes_ws, int width_rule) {
    if (n <= 0) return 0;
    if (n > 65535) return set_err(e, DV_ERR_UNSUPPORTED, "crop_quads_for_rec: more than 65535 quads per call");
    e->launch_begin("k_quad_homography", "crop", 0.0, n * 120.0);
    k_quad_homography<<<(n + 63) / 64, 64, 0, e->stream>>>(quads, box_counts, box_stride, per_page, n, dst_h, dst_w_pad, width_rule, minv_ws,
                                                           sizes_ws, dst_widths);
    e->launch_end();
    const int px = dst_h * dst_w_pad;
    e->launch_begin("k_crop_resize_fused", "crop", 0.0, static_cast<double>(n) * px * 3.0 * 2.0);
    k_crop_resize_fused<<<dim3((px + 255) / 256, n), 256, 0, e->stream>>>(pages, H, W, page_idx, box_counts ? per_page : 0, minv_ws, sizes_ws,
                                                                          dst_widths, dst_h, dst_w_pad, out);
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

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
