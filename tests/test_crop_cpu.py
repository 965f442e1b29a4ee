"""SURVEY.md 8(f)-1, det -> rec crop extraction.  CPU side: the oracle's crop_image against the golden crops produced by the
reference's own OcrCommonUtils.crop_image, the numpy restatement of cv2.warpPerspective (what the CUDA kernel implements)
against cv2 itself, and the host corner-ordering rule."""
import math
import os

import cv2
import numpy as np
import pytest

from oracle import crop_ref as ref
from oracle import gen_golden_crop as gen
from pdf_table_b200 import predictors

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "crop.npz")


def test_oracle_crop_image_matches_reference_golden():
    g = np.load(GOLDEN)
    img = gen.page()
    for k in range(int(g["n"])):
        assert np.array_equal(ref.crop_image(img, g[f"quad{k}"]), g[f"crop{k}"]), k


def test_warp_restatement_equals_cv2_on_golden_quads():
    g = np.load(GOLDEN)
    img = gen.page()
    for k in range(int(g["n"])):
        corners, trans, (w, h) = ref.crop_geometry(g[f"quad{k}"])
        t = cv2.getPerspectiveTransform(corners, trans)
        assert np.array_equal(ref.warp_perspective(img, t, w, h), g[f"crop{k}"]), k


def test_warp_restatement_equals_cv2_on_random_quads():
    rng = np.random.default_rng(5)
    for trial in range(60):
        hh, ww = int(rng.integers(60, 400)), int(rng.integers(60, 500))
        img = rng.integers(0, 256, (hh, ww, 3), dtype=np.uint8)
        cx, cy = rng.uniform(0, ww), rng.uniform(0, hh)  # centres anywhere: many quads hang over the border
        bw, bh, ang = rng.uniform(4, 300), rng.uniform(2, 70), rng.uniform(-0.8, 0.8)
        c, s = math.cos(ang), math.sin(ang)
        pts = np.array([[-bw / 2, -bh / 2], [bw / 2, -bh / 2], [bw / 2, bh / 2], [-bw / 2, bh / 2]]) @ np.array([[c, s], [-s, c]]) + [cx, cy]
        pts += rng.uniform(-2, 2, pts.shape)
        corners, trans, (w, h) = ref.crop_geometry(pts)
        if w < 1 or h < 1:
            continue
        t = cv2.getPerspectiveTransform(corners, trans)
        assert np.array_equal(ref.warp_perspective(img, t, w, h), cv2.warpPerspective(img, t, (w, h))), trial


def test_host_crop_geometry_matches_oracle():
    g = np.load(GOLDEN)
    for k in range(int(g["n"])):
        a, b = predictors.crop_geometry(g[f"quad{k}"]), ref.crop_geometry(g[f"quad{k}"])
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]
        assert a[2] == (g[f"crop{k}"].shape[1], g[f"crop{k}"].shape[0])


def test_resize_restatement_equals_cv2():
    rng = np.random.default_rng(12)
    for trial in range(80):
        hh, ww = int(rng.integers(1, 150)), int(rng.integers(1, 1200))
        if trial % 8 == 0:
            hh, ww = 64, 2 * int(rng.integers(10, 300))  # exact 2x reduction: cv2's box-average special case
        img = rng.integers(0, 256, (hh, ww, 3), dtype=np.uint8)
        ratio = ww / float(hh)
        dw = 804 if ratio > 804 / 32 else int(32 * ratio)  # keepratio_resize
        if dw < 1:
            continue
        assert np.array_equal(ref.resize_linear(img, dw, 32), cv2.resize(img, (dw, 32))), (trial, hh, ww)
    for trial in range(30):  # PP-OCR rec sizes (height 48, width 16 .. 1280)
        hh, ww = int(rng.integers(2, 120)), int(rng.integers(2, 900))
        img = rng.integers(0, 256, (hh, ww, 3), dtype=np.uint8)
        dw = int(rng.integers(16, 1281))
        assert np.array_equal(ref.resize_linear(img, dw, 48), cv2.resize(img, (dw, 48))), (trial, hh, ww, dw)


def test_homography_restatement_equals_cv2():
    """cv2.getPerspectiveTransform and cv2.invert restated (what k_quad_homography computes per quad) -- bit for bit."""
    rng = np.random.default_rng(4)
    for trial in range(150):
        cx, cy = rng.uniform(0, 960), rng.uniform(0, 960)
        bw, bh, ang = rng.uniform(4, 600), rng.uniform(3, 100), rng.uniform(-0.8, 0.8)
        c, s = math.cos(ang), math.sin(ang)
        pts = np.array([[-bw / 2, -bh / 2], [bw / 2, -bh / 2], [bw / 2, bh / 2], [-bw / 2, bh / 2]]) @ np.array([[c, s], [-s, c]]) + [cx, cy]
        if trial % 2:
            pts = np.rint(pts + rng.uniform(-2, 2, pts.shape))
        corners, trans, _ = ref.crop_geometry(pts.astype(np.float32))
        t = cv2.getPerspectiveTransform(corners, trans)
        assert np.array_equal(ref.get_perspective_transform(corners, trans), t), trial
        assert np.array_equal(ref.invert3(t), cv2.invert(t)[1]), trial


def test_warp_affine_restatement_equals_cv2():
    """cv2.warpAffine restated (what k_warp_affine_u8 implements), incl. the Lore pre-process matrices, and the host inversion."""
    rng = np.random.default_rng(6)
    for trial in range(25):
        hh, ww = int(rng.integers(50, 500)), int(rng.integers(50, 700))
        img = rng.integers(0, 256, (hh, ww, 3), dtype=np.uint8)
        if trial % 3 == 0:  # the centre-anchored similarity of TableLorePreProcessor
            m = predictors.lore_affine(np.array([ww / 2.0, hh / 2.0], np.float32), max(hh, ww) * 1.0, 256, 256)
            w, h = 256, 256
        else:
            s, th = rng.uniform(0.3, 3.0), rng.uniform(-0.4, 0.4)
            m = np.array([[s * math.cos(th), -s * math.sin(th), rng.uniform(-50, 50)], [s * math.sin(th), s * math.cos(th), rng.uniform(-50, 50)]])
            w, h = int(rng.integers(32, 300)), int(rng.integers(32, 300))
        assert np.array_equal(ref.warp_affine(img, m, w, h), cv2.warpAffine(img, m, (w, h), flags=cv2.INTER_LINEAR)), trial


def test_table_crop_rect_is_the_numpy_slice_of_crop_image_by_box():
    """predictors.table_crop_rect == the slice img[round(y1):round(y2), round(x1):round(x2)] of OcrCommonUtils.crop_image_by_box,
    incl. half-to-even rounding, clamping at the far borders and the empty cases (negative start wraps in numpy -> empty)."""
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, (70, 90, 3), dtype=np.uint8)
    cases = [[0.5, 1.5, 2.5, 3.5], [10.0, 5.0, 200.0, 300.0], [-3.0, 2.0, 50.0, 40.0], [30.0, 30.0, 30.4, 60.0], [89.6, 0.0, 95.0, 70.0]]
    cases += [list(rng.uniform(-5, 100, 4)) for _ in range(200)]
    empty = 0
    for x1, y1, x2, y2 in cases:
        want = img[round(y1):round(y2), round(x1):round(x2)]
        if want.size == 0:
            empty += 1
            with pytest.raises(ValueError):
                predictors.table_crop_rect([x1, y1, x2, y2], 70, 90)
            continue
        x0, y0, cw, ch = predictors.table_crop_rect([x1, y1, x2, y2], 70, 90)
        assert np.array_equal(img[y0:y0 + ch, x0:x0 + cw], want)
    assert 0 < empty < len(cases)


def test_warp_of_a_page_slice_equals_warp_of_the_crop():
    """What k_warp_affine_rects_u8 computes, restated: warping the slice in place (page pitch, zero border at the slice's edges)
    is cv2.warpAffine of the cut-out crop with the crop's own Lore matrix."""
    page = np.random.default_rng(12).integers(0, 256, (300, 400, 3), dtype=np.uint8)
    for bbox in ([20.2, 30.7, 380.5, 290.1], [0.0, 0.0, 400.0, 300.0], [350.0, 10.0, 400.0, 300.0]):
        x0, y0, cw, ch = predictors.table_crop_rect(bbox, 300, 400)
        crop = np.ascontiguousarray(page[y0:y0 + ch, x0:x0 + cw])
        m = predictors.lore_affine(np.array([cw / 2.0, ch / 2.0], np.float32), max(ch, cw) * 1.0, 256, 256)
        assert np.array_equal(ref.warp_affine(page[y0:y0 + ch, x0:x0 + cw], m, 256, 256), cv2.warpAffine(crop, m, (256, 256), flags=cv2.INTER_LINEAR))
        assert np.array_equal(predictors.invert_affine(m), ref.invert_affine(m))


def test_order_point_matches_reference_golden():
    """predictors.order_point == OcrCommonUtils.order_point on the 300 quads of tests/golden/glue.npz (oracle/gen_golden_glue.py)."""
    g = np.load(os.path.join(os.path.dirname(GOLDEN), "glue.npz"))
    for q, f32, want in zip(g["quads"], g["is_f32"], g["ordered"]):
        got = predictors.order_point(q.astype(np.float32) if f32 else q)
        assert got.dtype == np.float32 and np.array_equal(got, want)


def test_sort_det_boxes_is_the_orchestrator_sort():
    """Reading order: ascending 0.01 * mean x + mean y, ties keep the detector's order (python's stable sort)."""
    rng = np.random.default_rng(13)
    boxes = rng.integers(0, 960, (40, 8)).astype(np.float32)
    boxes[7] = boxes[3]  # an exact tie
    got = predictors.sort_det_boxes(boxes)
    key = [0.01 * sum(r[::2]) / 4 + sum(r[1::2]) / 4 for r in boxes.tolist()]
    order = sorted(range(40), key=lambda i: (key[i], i))
    assert np.array_equal(got, boxes[order].astype(np.float64))
    assert order.index(3) + 1 == order.index(7)
    assert predictors.sort_det_boxes(np.zeros((0, 8), np.float32)).shape[0] == 0


def test_resize_restatement_equals_cv2_on_page_shapes():
    """What dv_resize_linear_u8 implements, at the shapes of the page pre-processors: DetResizeForTest's
    down-scale to a multiple of 32, its slight re-scale of a page that is already below the limit, an up-scale, an exact 2x
    reduction (cv2's box-average special case) and PicoDet's non-uniform resize to 608 x 800."""
    rng = np.random.default_rng(14)
    for (sh, sw), dst in [((1400, 1100), None), ((700, 900), None), ((333, 517), (800, 608)), ((1920, 1216), (960, 608)),
                          ((1600, 1216), (800, 608)), ((200, 150), (800, 608))]:
        img = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        if dst is not None:
            dh, dw = dst
        else:
            dh, dw = predictors.det_resize_shape(sh, sw, 960, "max")
            assert dh % 32 == 0 and dw % 32 == 0 and max(dh, dw) <= 960
        assert np.array_equal(ref.resize_linear(img, dw, dh), cv2.resize(img, (dw, dh))), (sh, sw, dh, dw)
