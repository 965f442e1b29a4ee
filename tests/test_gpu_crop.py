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


def test_device_only_glue_matches_cv2(eng):
    """dv_crop_quads_for_rec: geometry + homography + inverse + warp + keep-ratio resize with no host step, against the
    reference's host calls (crop_geometry rule, cv2.getPerspectiveTransform, cv2.invert, cv2.warpPerspective, cv2.resize)."""
    rng = np.random.default_rng(33)
    pages = rng.integers(0, 256, (2, 500, 700, 3), dtype=np.uint8)
    quads, pidx = [], []
    for k in range(160):
        cx, cy = rng.uniform(0, 700), rng.uniform(0, 500)
        bw, bh, ang = rng.uniform(3, 650), rng.uniform(2, 90), rng.uniform(-0.6, 0.6)
        if k % 9 == 0:
            bw, bh, ang = 2 * float(rng.integers(20, 150)) + 0.2, 64.2, 0.0
        c, s = math.cos(ang), math.sin(ang)
        p = np.array([[-bw / 2, -bh / 2], [bw / 2, -bh / 2], [bw / 2, bh / 2], [-bw / 2, bh / 2]]) @ np.array([[c, s], [-s, c]]) + [cx, cy]
        p = np.roll(p, k % 4, axis=0)
        quads.append((np.rint(p) if k % 3 == 0 else p).astype(np.float32))
        pidx.append(k % 2)
    quads.append(np.array([[10, 10], [10.4, 10], [10.4, 30], [10, 30]], np.float32))  # zero-width crop
    pidx.append(0)
    q = torch.from_numpy(np.stack(quads)).cuda()
    out, widths, sizes, minv = eng.crop_quads_for_rec(torch.from_numpy(pages).cuda(), q, torch.tensor(pidx, dtype=torch.int32, device="cuda"))
    out, widths, sizes, minv = out.cpu().numpy(), widths.cpu().numpy(), sizes.cpu().numpy(), minv.cpu().numpy()
    assert widths[-1] == 0 and not out[-1].any()
    checked = 0
    for k, quad in enumerate(quads[:-1]):
        corners, trans, (w, h) = predictors.crop_geometry(quad)
        assert (sizes[k, 0], sizes[k, 1]) == (w, h), k
        if w <= 0 or h <= 0:
            assert widths[k] == 0
            continue
        t = cv2.getPerspectiveTransform(corners, trans)
        assert np.array_equal(minv[k], cv2.invert(t)[1]), k
        want = predictors.keepratio_resize(cv2.warpPerspective(pages[pidx[k]], t, (w, h)))
        assert widths[k] == want.shape[1], k
        assert np.array_equal(out[k, :, : widths[k]], want), k
        assert not out[k, :, widths[k]:].any()
        checked += 1
    assert checked > 140


def test_lore_preprocess_on_device_equals_host(eng):
    """a10 with the warp on the device: lore_preprocess_device == lore_preprocess (cv2.warpAffine on the host), pixels and meta,
    for an up-scaled and a down-scaled page; and the raw kernel against cv2.warpAffine for rotated / sheared matrices."""
    from pdf_table_b200 import synth

    for seed, (hh, ww) in enumerate([(300, 520), (1400, 1100), (1024, 1024)]):
        img = synth.synthetic_page(60 + seed, hh, ww)
        want, meta = predictors.lore_preprocess(img)
        got, meta_d = predictors.lore_preprocess_device(eng, img)
        assert np.array_equal(meta, meta_d)
        assert np.array_equal(got.cpu().numpy(), want)
    rng = np.random.default_rng(8)
    img = rng.integers(0, 256, (333, 517, 3), dtype=np.uint8)
    dev = torch.from_numpy(img).cuda()
    for trial in range(12):
        s, th = rng.uniform(0.3, 3.0), rng.uniform(-0.4, 0.4)
        m = np.array([[s * math.cos(th), -s * math.sin(th) + 0.1, rng.uniform(-60, 60)], [s * math.sin(th), s * math.cos(th), rng.uniform(-60, 60)]])
        w, h = int(rng.integers(32, 400)), int(rng.integers(32, 400))
        got = eng.warp_affine_u8(dev, predictors.invert_affine(m), w, h).cpu().numpy()
        assert np.array_equal(got, cv2.warpAffine(img, m, (w, h), flags=cv2.INTER_LINEAR)), trial
