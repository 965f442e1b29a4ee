"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement of the det -> rec crop extraction, SURVEY.md 8(f)-1:
  * crop_geometry / crop_image: OcrCommonUtils.crop_image (utils/ocr/ocr_common_utils.py:214-262) -- corner ordering,
    float32 corner arrays, crop size from the mid-line distances, cv2.getPerspectiveTransform, cv2.warpPerspective;
  * warp_perspective: what the CUDA kernel implements, a numpy restatement of cv2.warpPerspective's INTER_LINEAR /
    BORDER_CONSTANT path on uint8 (OpenCV imgwarp.cpp WarpPerspectiveInvoker + remapBilinear, fixed point 1/32 pixel,
    15-bit weights).
  * resize_linear: numpy restatement of cv2.resize's default INTER_LINEAR on uint8 (OpenCV resize.cpp, 11-bit coefficients,
    the 2x2 box average for an exact 2x reduction) -- what k_resize_linear_u8 implements; checked against cv2 itself.
Pinned: crop_image against the reference's own function (oracle/gen_golden_crop.py -> tests/golden/crop.npz), and
warp_perspective against cv2.warpPerspective itself at test time (tests/test_crop_cpu.py).
"""
from __future__ import annotations

import math

import cv2
import numpy as np


def crop_geometry(position):
    """position: [4,2] quad -> (corners float32 [4,2], corners_trans float32 [4,2], (w, h)); ocr_common_utils.py:227-257."""
    position = np.asarray(position).tolist()
    for i in range(4):
        for j in range(i + 1, 4):
            if position[i][0] > position[j][0]:
                position[i], position[j] = position[j], position[i]
    if position[0][1] > position[1][1]:
        position[0], position[1] = position[1], position[0]
    if position[2][1] > position[3][1]:
        position[2], position[3] = position[3], position[2]
    x1, y1 = position[0]
    x2, y2 = position[2]
    x3, y3 = position[3]
    x4, y4 = position[1]

    def distance(xa, ya, xb, yb):
        return math.sqrt(pow(xa - xb, 2) + pow(ya - yb, 2))

    corners = np.zeros((4, 2), np.float32)
    corners[0] = [x1, y1]
    corners[1] = [x2, y2]
    corners[2] = [x4, y4]
    corners[3] = [x3, y3]
    img_width = distance((x1 + x4) / 2, (y1 + y4) / 2, (x2 + x3) / 2, (y2 + y3) / 2)
    img_height = distance((x1 + x2) / 2, (y1 + y2) / 2, (x4 + x3) / 2, (y4 + y3) / 2)
    trans = np.zeros((4, 2), np.float32)
    trans[0] = [0, 0]
    trans[1] = [img_width - 1, 0]
    trans[2] = [0, img_height - 1]
    trans[3] = [img_width - 1, img_height - 1]
    return corners, trans, (int(img_width), int(img_height))


def crop_image(img, position):
    corners, trans, size = crop_geometry(position)
    return cv2.warpPerspective(img, cv2.getPerspectiveTransform(corners, trans), size)


def warp_perspective(img, transform, w, h):
    """== cv2.warpPerspective(img, transform, (w, h)) for uint8 HWC (defaults: INTER_LINEAR, BORDER_CONSTANT 0)."""
    m = cv2.invert(np.asarray(transform, np.float64))[1].ravel()
    sh, sw = img.shape[:2]
    bh0 = min(16, h)
    bw0 = min(1024 // bh0, w)
    ys = np.arange(h, dtype=np.float64)[:, None]
    xs = np.arange(w)
    xb = (xs // bw0 * bw0).astype(np.float64)[None, :]
    x1 = (xs % bw0).astype(np.float64)[None, :]
    x0 = m[0] * xb + m[1] * ys + m[2]
    y0 = m[3] * xb + m[4] * ys + m[5]
    w0 = m[6] * xb + m[7] * ys + m[8]
    wd = w0 + m[6] * x1
    with np.errstate(divide="ignore", invalid="ignore"):
        wi = np.where(wd != 0, 32.0 / wd, 0.0)
    fx = np.maximum(-2147483648.0, np.minimum(2147483647.0, (x0 + m[0] * x1) * wi))
    fy = np.maximum(-2147483648.0, np.minimum(2147483647.0, (y0 + m[3] * x1) * wi))
    xi = np.rint(fx).astype(np.int64)
    yi = np.rint(fy).astype(np.int64)
    sx = np.clip(xi >> 5, -32768, 32767)
    sy = np.clip(yi >> 5, -32768, 32767)
    ax, ay = xi & 31, yi & 31

    def fetch(yy, xx):
        ok = (yy >= 0) & (yy < sh) & (xx >= 0) & (xx < sw)
        v = img[np.clip(yy, 0, sh - 1), np.clip(xx, 0, sw - 1)].astype(np.int64)
        return np.where(ok[..., None], v, 0)

    acc = (fetch(sy, sx) * ((32 - ax) * (32 - ay) * 32)[..., None] + fetch(sy, sx + 1) * (ax * (32 - ay) * 32)[..., None]
           + fetch(sy + 1, sx) * ((32 - ax) * ay * 32)[..., None] + fetch(sy + 1, sx + 1) * (ax * ay * 32)[..., None])
    return ((acc + (1 << 14)) >> 15).astype(np.uint8)


def resize_linear(img, dw, dh):
    """== cv2.resize(img, (dw, dh)) for uint8 HWC (default INTER_LINEAR)."""
    sh, sw = img.shape[:2]
    src = img.astype(np.int64)
    if sw == 2 * dw and sh == 2 * dh:  # cv2 turns an exact 2x INTER_LINEAR reduction into the INTER_AREA fast path
        return ((src[0::2, 0::2] + src[0::2, 1::2] + src[1::2, 0::2] + src[1::2, 1::2] + 2) >> 2).astype(np.uint8)

    def coeffs(dn, sn, clamp):
        scale = 1.0 / (dn / sn)
        f = ((np.arange(dn) + 0.5) * scale - 0.5).astype(np.float32)
        s = np.floor(f).astype(np.int64)
        f = f - s.astype(np.float32)
        if clamp:  # columns: the fraction is dropped at the borders; rows keep it and clip the row indices instead
            lo, hi = s < 0, s >= sn - 1
            f = np.where(lo | hi, np.float32(0), f)
            s = np.where(lo, 0, np.where(hi, sn - 1, s))
        return s, np.rint((np.float32(1) - f) * np.float32(2048)).astype(np.int64), np.rint(f * np.float32(2048)).astype(np.int64)

    sx, a0, a1 = coeffs(dw, sw, True)
    sy, b0, b1 = coeffs(dh, sh, False)
    hrow = src[:, sx] * a0[None, :, None] + src[:, np.minimum(sx + 1, sw - 1)] * a1[None, :, None]
    s0, s1 = hrow[np.clip(sy, 0, sh - 1)], hrow[np.clip(sy + 1, 0, sh - 1)]
    out = (((b0[:, None, None] * (s0 >> 4)) >> 16) + ((b1[:, None, None] * (s1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def get_perspective_transform(src, dst):
    """== cv2.getPerspectiveTransform(src, dst) (OpenCV 4.13): the 8x8 system whose four -x*X products are formed in float32
    (Point2f arithmetic), solved by cv::hal::LU64f (partial pivoting, d = -1/pivot, back substitution) in double."""
    src = np.asarray(src, np.float32)
    dst = np.asarray(dst, np.float32)
    a = np.zeros((8, 8))
    b = np.zeros(8)
    for i in range(4):
        sx, sy, dx, dy = src[i, 0], src[i, 1], dst[i, 0], dst[i, 1]  # float32 scalars: the products below round to float32
        a[i, 0] = a[i + 4, 3] = sx
        a[i, 1] = a[i + 4, 4] = sy
        a[i, 2] = a[i + 4, 5] = 1
        a[i, 6], a[i, 7], a[i + 4, 6], a[i + 4, 7] = np.float32(-sx * dx), np.float32(-sy * dx), np.float32(-sx * dy), np.float32(-sy * dy)
        b[i], b[i + 4] = dx, dy
    for i in range(8):
        k = i
        for j in range(i + 1, 8):
            if abs(a[j, i]) > abs(a[k, i]):
                k = j
        if abs(a[k, i]) < np.finfo(np.float64).eps * 100:
            return None
        if k != i:
            a[[i, k], i:] = a[[k, i], i:]
            b[[i, k]] = b[[k, i]]
        d = -1 / a[i, i]
        for j in range(i + 1, 8):
            alpha = a[j, i] * d
            for c in range(i + 1, 8):
                a[j, c] += alpha * a[i, c]
            b[j] += alpha * b[i]
    for i in range(7, -1, -1):
        s = b[i]
        for c in range(i + 1, 8):
            s -= a[i, c] * b[c]
        b[i] = s / a[i, i]
    return np.append(b, 1.0).reshape(3, 3)


def invert3(t):
    """== cv2.invert(t)[1] for a 3x3 double matrix (the closed-form cofactor path)."""
    s = np.asarray(t, np.float64)
    d = s[0, 0] * (s[1, 1] * s[2, 2] - s[1, 2] * s[2, 1]) - s[0, 1] * (s[1, 0] * s[2, 2] - s[1, 2] * s[2, 0]) + s[0, 2] * (s[1, 0] * s[2, 1] - s[1, 1] * s[2, 0])
    if d == 0:
        return None
    d = 1.0 / d
    return np.array([(s[1, 1] * s[2, 2] - s[1, 2] * s[2, 1]) * d, (s[0, 2] * s[2, 1] - s[0, 1] * s[2, 2]) * d, (s[0, 1] * s[1, 2] - s[0, 2] * s[1, 1]) * d,
                     (s[1, 2] * s[2, 0] - s[1, 0] * s[2, 2]) * d, (s[0, 0] * s[2, 2] - s[0, 2] * s[2, 0]) * d, (s[0, 2] * s[1, 0] - s[0, 0] * s[1, 2]) * d,
                     (s[1, 0] * s[2, 1] - s[1, 1] * s[2, 0]) * d, (s[0, 1] * s[2, 0] - s[0, 0] * s[2, 1]) * d, (s[0, 0] * s[1, 1] - s[0, 1] * s[1, 0]) * d]).reshape(3, 3)


def invert_affine(m23):
    """The inversion cv2.warpAffine applies to its 2x3 matrix when WARP_INVERSE_MAP is not set (OpenCV imgwarp.cpp), same doubles."""
    m = np.asarray(m23, np.float64).copy().ravel()
    d = m[0] * m[4] - m[1] * m[3]
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[4] * d, m[0] * d
    m[0] = a11
    m[1] *= -d
    m[3] *= -d
    m[4] = a22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    return m.reshape(2, 3)


def warp_affine(img, m23, w, h):
    """== cv2.warpAffine(img, m23, (w, h), flags=cv2.INTER_LINEAR) for uint8 HWC (border constant 0): OpenCV's
    WarpAffineInvoker (matrix inverted first, 10-bit fixed-point coordinates, rounding offset 16) + the bilinear remap."""
    m = invert_affine(m23).ravel()
    sh, sw = img.shape[:2]
    xs = np.arange(w, dtype=np.float64)
    ys = np.arange(h, dtype=np.float64)
    adelta = np.rint(m[0] * xs * 1024).astype(np.int64)
    bdelta = np.rint(m[3] * xs * 1024).astype(np.int64)
    x0 = np.rint((m[1] * ys + m[2]) * 1024).astype(np.int64) + 16
    y0 = np.rint((m[4] * ys + m[5]) * 1024).astype(np.int64) + 16
    xi = (x0[:, None] + adelta[None, :]) >> 5
    yi = (y0[:, None] + bdelta[None, :]) >> 5
    sx = np.clip(xi >> 5, -32768, 32767)
    sy = np.clip(yi >> 5, -32768, 32767)
    ax, ay = xi & 31, yi & 31

    def fetch(yy, xx):
        ok = (yy >= 0) & (yy < sh) & (xx >= 0) & (xx < sw)
        v = img[np.clip(yy, 0, sh - 1), np.clip(xx, 0, sw - 1)].astype(np.int64)
        return np.where(ok[..., None], v, 0)

    acc = (fetch(sy, sx) * ((32 - ax) * (32 - ay) * 32)[..., None] + fetch(sy, sx + 1) * (ax * (32 - ay) * 32)[..., None]
           + fetch(sy + 1, sx) * ((32 - ax) * ay * 32)[..., None] + fetch(sy + 1, sx + 1) * (ax * ay * 32)[..., None])
    return ((acc + (1 << 14)) >> 15).astype(np.uint8)
