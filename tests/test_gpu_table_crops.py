"""SURVEY.md 8(f)-2 on the GPU: the layout -> table-structure glue (dv_crop_tables_for_tsr) against cv2 -- every table slice of a
batch of resident pages warped exactly as cv2.warpAffine warps the cut-out crop -- and OcrTableStructureTask.recognize_tables
against the per-crop host flow of the reference's orchestrator (ocr_pdf/ocr_system_task.py:184-198)."""
import math

import cv2
import numpy as np
import pytest
import torch

from pdf_table_b200 import predictors, synth
from pdf_table_b200.engine import Engine

pytestmark = pytest.mark.gpu


def test_table_slices_warp_equals_cv2_on_the_cut_out_crop():
    eng = Engine("post")
    pages = np.stack([synth.synthetic_page(70 + k, 700, 900) for k in range(3)])
    dev = torch.from_numpy(pages).cuda()
    rng = np.random.default_rng(9)
    # (a) the Lore pre-process matrices at the network resolution, incl. a slice touching all four page borders
    boxes = [(0, [0.0, 0.0, 900.0, 700.0]), (1, [100.4, 50.5, 620.5, 333.49]), (2, [851.5, 3.2, 899.6, 699.7]), (1, [10.0, 640.0, 420.0, 700.0])]
    rects, minv, want = [], [], []
    for pg, bbox in boxes:
        x0, y0, cw, ch = predictors.table_crop_rect(bbox, 700, 900)
        crop = pages[pg, y0:y0 + ch, x0:x0 + cw]
        assert crop.shape[:2] == pages[pg][round(bbox[1]):round(bbox[3]), round(bbox[0]):round(bbox[2])].shape[:2]
        w, meta = predictors.lore_preprocess(np.ascontiguousarray(crop))
        m = predictors.lore_affine(np.array([cw / 2.0, ch / 2.0], np.float32), max(ch, cw) * 1.0, 1024, 1024)
        rects.append([pg, x0, y0, cw, ch])
        minv.append(predictors.invert_affine(m))
        want.append(w)
    got = eng.crop_tables_for_tsr(dev, np.array(rects, np.int32), np.stack(minv), 1024, 1024).cpu().numpy()
    for k in range(len(boxes)):
        assert np.array_equal(got[k], want[k]), k
    # (b) rotated / sheared matrices on random slices, a small output frame; the last rect is not inside its page -> zero image
    rects, minv, want = [], [], []
    for trial in range(16):
        pg = int(rng.integers(0, 3))
        cw, ch = int(rng.integers(8, 600)), int(rng.integers(8, 500))
        x0, y0 = int(rng.integers(0, 900 - cw + 1)), int(rng.integers(0, 700 - ch + 1))
        s, th = rng.uniform(0.3, 3.0), rng.uniform(-0.4, 0.4)
        m = np.array([[s * math.cos(th), -s * math.sin(th) + 0.1, rng.uniform(-60, 60)], [s * math.sin(th), s * math.cos(th), rng.uniform(-60, 60)]])
        rects.append([pg, x0, y0, cw, ch])
        minv.append(predictors.invert_affine(m))
        want.append(cv2.warpAffine(np.ascontiguousarray(pages[pg, y0:y0 + ch, x0:x0 + cw]), m, (200, 160), flags=cv2.INTER_LINEAR))
    rects.append([1, 500, 400, 450, 100])
    minv.append(minv[0])
    want.append(np.zeros((160, 200, 3), np.uint8))
    got = eng.crop_tables_for_tsr(dev, np.array(rects, np.int32), np.stack(minv), 200, 160).cpu().numpy()
    for k in range(len(rects)):
        assert np.array_equal(got[k], want[k]), k
    assert eng.crop_tables_for_tsr(dev, np.zeros((0, 5), np.int32), np.zeros((0, 2, 3)), 64, 64).shape == (0, 64, 64, 3)
    eng.close()


def test_recognize_tables_equals_per_crop_calls():
    """All tables of two resident pages in one call == the task called on the cut-out crops (ndarray inputs), cell for cell."""
    sd = synth.lore_dla34_state_dict(0)
    sd["hm.2.bias"] = np.array([-0.3, -3.5], np.float32)  # random weights: shift the heat maps so that cells / corners pass the gates
    task = predictors.OcrTableStructureTask(model="Lore", task_type="wtw", state_dict=(sd, synth.lore_processor_state_dict(0)),
                                            max_cells_per_image=3000)
    pages = np.stack([synth.synthetic_page(80 + k, 960, 960) for k in range(2)])
    tables = [{"bbox": [40.3, 60.7, 700.2, 500.5], "page": 0, "label": "table"}, {"bbox": [0.0, 300.0, 960.0, 960.0], "page": 1, "label": "table"}]
    out = task.recognize_tables(torch.from_numpy(pages).cuda(), tables)
    crops = []
    for tb in tables:
        x1, y1, x2, y2 = tb["bbox"]
        crops.append(np.ascontiguousarray(pages[tb["page"]][round(y1):round(y2), round(x1):round(x2)]))  # crop_image_by_box
    want = task(crops)
    assert len(out) == 2
    for (bbox, got), w, tb in zip(out, want, tables):
        assert bbox == tb["bbox"] and got["inputs"] == tb["bbox"]
        np.testing.assert_array_equal(got["polygons"], w["polygons"])
        np.testing.assert_array_equal(got["logi"], w["logi"])
    assert sum(len(r["polygons"]) for _, r in out) > 5
    assert task.recognize_tables(pages[0], []) == []
    with pytest.raises(ValueError):
        task.recognize_tables(pages[0], [{"bbox": [50.0, 50.0, 50.2, 400.0]}])
