"""SURVEY.md 8(f)-1 on the GPU: predictors.crop_images (host homography + dv_warp_perspective_u8) against the golden crops of
the reference's OcrCommonUtils.crop_image and against cv2.warpPerspective on random quads -- bit-exact."""
import math
import os

import cv2
import numpy as np
import pytest
import torch

from oracle import gen_golden_crop as gen
from pdf_table_b200 import predictors
from pdf_table_b200.engine import Engine

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "crop.npz")


@pytest.fixture(scope="module")
def eng():
    e = Engine("post")
    yield e
    e.close()


def test_crops_match_reference_golden(eng):
    g = np.load(GOLDEN)
    page = torch.from_numpy(gen.page()).cuda()
    quads = [g[f"quad{k}"] for k in range(int(g["n"]))]
    crops = predictors.crop_images(eng, page, quads)
    assert len(crops) == len(quads)
    for k, c in enumerate(crops):
        assert c.is_cuda and c.dtype == torch.uint8
        assert np.array_equal(c.cpu().numpy(), g[f"crop{k}"]), k


def test_crops_match_cv2_on_random_quads(eng):
    rng = np.random.default_rng(9)
    img = rng.integers(0, 256, (700, 900, 3), dtype=np.uint8)
    page = torch.from_numpy(img).cuda()
    quads = []
    for _ in range(200):
        cx, cy = rng.uniform(-20, 920), rng.uniform(-20, 720)
        bw, bh, ang = rng.uniform(4, 500), rng.uniform(2, 90), rng.uniform(-0.8, 0.8)
        c, s = math.cos(ang), math.sin(ang)
        pts = np.array([[-bw / 2, -bh / 2], [bw / 2, -bh / 2], [bw / 2, bh / 2], [-bw / 2, bh / 2]]) @ np.array([[c, s], [-s, c]]) + [cx, cy]
        quads.append(pts + rng.uniform(-2, 2, pts.shape))
    crops = predictors.crop_images(eng, page, quads)
    checked = 0
    for q, c in zip(quads, crops):
        corners, trans, (w, h) = predictors.crop_geometry(q)
        if w < 1 or h < 1:
            assert c is None
            continue
        want = cv2.warpPerspective(img, cv2.getPerspectiveTransform(corners, trans), (w, h))
        assert np.array_equal(c.cpu().numpy(), want)
        checked += 1
    assert checked > 150


def test_no_quads_and_bad_arguments(eng):
    page = torch.zeros((64, 64, 3), dtype=torch.uint8, device="cuda")
    assert predictors.crop_images(eng, page, []) == []
    with pytest.raises(ValueError):
        eng.warp_perspective_u8(page, np.eye(3)[None], np.array([[0, 5]], np.int32))


def test_crops_for_recognition_match_host_pipeline(eng):
    """crop_image -> keepratio_resize on the device == the same two cv2 calls on the host, for every quad; the padded batch
    is what convnextvit_forward_u8 reads."""
    rng = np.random.default_rng(21)
    img = rng.integers(0, 256, (600, 800, 3), dtype=np.uint8)
    page = torch.from_numpy(img).cuda()
    quads = []
    for k in range(120):
        cx, cy = rng.uniform(0, 800), rng.uniform(0, 600)
        bw, bh, ang = rng.uniform(3, 700), rng.uniform(2, 100), rng.uniform(-0.5, 0.5)
        if k % 10 == 0:
            bw, bh, ang = 2 * float(rng.integers(20, 200)) + 0.2, 64.2, 0.0  # exact 2x reductions (64 -> 32)
        c, s = math.cos(ang), math.sin(ang)
        quads.append(np.array([[-bw / 2, -bh / 2], [bw / 2, -bh / 2], [bw / 2, bh / 2], [-bw / 2, bh / 2]]) @ np.array([[c, s], [-s, c]]) + [cx, cy])
    batch, widths, keep = predictors.crops_for_recognition(eng, page, quads)
    assert batch.shape[0] == len(keep) == len(widths) and batch.shape[1] == 32 and batch.shape[2] == max(widths)
    got = batch.cpu().numpy()
    exact2x = 0
    for row, (k, w) in enumerate(zip(keep, widths)):
        corners, trans, size = predictors.crop_geometry(quads[k])
        crop = cv2.warpPerspective(img, cv2.getPerspectiveTransform(corners, trans), size)
        want = predictors.keepratio_resize(crop)
        assert want.shape == (32, w, 3)
        assert np.array_equal(got[row, :, :w], want), k
        assert not got[row, :, w:].any()
        exact2x += int(crop.shape[0] == 64 and crop.shape[1] == 2 * w)
    assert len(keep) > 100 and exact2x > 0
