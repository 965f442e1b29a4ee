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
