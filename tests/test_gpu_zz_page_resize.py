"""SURVEY.md 8(f)-2 on the GPU: the page pre-processors' cv2.resize for pages that are already on the device
(Engine.resize_pages_u8 over dv_resize_linear_u8, bit-exact against cv2) and the detection / layout predictors fed with cuda
tensors instead of ndarrays -- identical results, nothing but the raw page goes up."""
import numpy as np
import pytest
import torch

import cv2
from pdf_table_b200 import predictors, synth
from pdf_table_b200.engine import Engine

pytestmark = pytest.mark.gpu


def test_page_resize_equals_cv2():
    eng = Engine("post")
    rng = np.random.default_rng(15)
    for (sh, sw), dst in [((1400, 1100), None), ((700, 900), None), ((333, 517), (800, 608)), ((1920, 1216), (960, 608)), ((200, 150), (800, 608))]:
        pages = rng.integers(0, 256, (2, sh, sw, 3), dtype=np.uint8)
        dh, dw = dst if dst is not None else predictors.det_resize_shape(sh, sw, 960, "max")
        got = eng.resize_pages_u8(torch.from_numpy(pages).cuda(), dw, dh).cpu().numpy()
        assert got.shape == (2, dh, dw, 3)
        for k in range(2):
            assert np.array_equal(got[k], cv2.resize(pages[k], (dw, dh))), (sh, sw, dh, dw, k)
    page = synth.synthetic_page(90, 1400, 1100)
    want, ratios = predictors.det_resize_for_test(page, 960, "max")
    got, ratios_d = predictors.det_resize_for_test_device(eng, torch.from_numpy(page).cuda(), 960, "max")
    assert ratios == ratios_d and np.array_equal(got.cpu().numpy(), want)
    eng.close()


def test_detection_task_accepts_device_pages():
    task = predictors.OcrDetectionTask(model="db_pp", state_dict=synth.dbnet_r18_state_dict(0), thresh=0.2)
    pages = [synth.synthetic_page(91, 1400, 1100), synth.synthetic_page(92, 320, 480), synth.synthetic_page(93, 700, 900)]
    want = task(pages)
    got = task([torch.from_numpy(p).cuda() for p in pages])
    mixed = task([pages[0], torch.from_numpy(pages[1]).cuda(), pages[2]])
    for w, g, m in zip(want, got, mixed):
        np.testing.assert_array_equal(g, w)
        np.testing.assert_array_equal(m, w)
    with pytest.raises(TypeError):
        task([torch.zeros((3, 64, 64), dtype=torch.float32).cuda()])


def test_layout_task_accepts_device_pages():
    bb, nk, hd = synth.picodet_state_dicts(0, 5)
    hd = dict(hd)
    for lvl in range(4):  # random weights never score above 0.5: lift the class bias so that some anchors fire
        b = hd[f"head_cls{lvl}.bias"].copy()
        b[:5] += 1.6
        hd[f"head_cls{lvl}.bias"] = b
    task = predictors.OcrLayoutTask(model="picodet", task_type="en", state_dict=(bb, nk, hd), score_threshold=0.5)
    pages = [synth.synthetic_page(94, 1000, 760), synth.synthetic_page(95, 640, 900)]
    want = task(pages)
    got = task([torch.from_numpy(p).cuda() for p in pages])
    assert sum(len(r) for r in want) > 0
    for rw, rg in zip(want, got):
        assert len(rw) == len(rg)
        for a, b in zip(rw, rg):
            assert a["label"] == b["label"] and a["score"] == b["score"] and np.array_equal(a["bbox"], b["bbox"])
