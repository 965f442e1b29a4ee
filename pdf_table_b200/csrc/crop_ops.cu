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

}  // namespace

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
