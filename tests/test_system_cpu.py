"""Host logic of pdf_table_b200/system.py (the model-calling methods of the reference's OcrSystemTask, ocr_system_task.py:146-330)
with stub predictors: ordering, filtering, return conventions and error behaviour -- no GPU, no engine."""
import numpy as np
import pytest
import torch

from pdf_table_b200 import predictors, system


class _Det:
    def __init__(self, boxes):
        self.boxes = boxes

    def __call__(self, image):
        return [self.boxes]


class _Rec:
    device = 0

    def __init__(self):
        self.calls = []

    def recognize_page(self, page, positions):
        self.calls.append((page, positions))
        return [None if k == 1 else f"t{k}" for k in range(len(positions))]


class _Tsr:
    device = 0

    def __init__(self):
        self.calls = []

    def recognize_tables(self, page, tables):
        self.calls.append((page, tables))
        return [[t["bbox"], {"polygons": np.zeros((1, 8), np.float32), "logi": np.zeros((1, 4), np.float32), "inputs": t["bbox"]}] for t in tables]

    def __call__(self, image):
        return [{"polygons": np.ones((2, 8), np.float32), "inputs": image}]


def test_text_detection_sorts_into_reading_order():
    boxes = np.array([[500, 300, 600, 300, 600, 320, 500, 320], [10, 300, 90, 300, 90, 320, 10, 320], [10, 20, 90, 20, 90, 40, 10, 40]], np.float32)
    result, metric = system.OcrSystemTask(text_detector=_Det(boxes)).text_detection("page")
    assert np.array_equal(result, boxes[[2, 1, 0]]) and set(metric) == {"use_time"}
    with pytest.raises(RuntimeError):
        system.OcrSystemTask().text_detection("page")


def test_text_recognition_orders_corners_and_keeps_the_reference_output_records():
    rec = _Rec()
    page = torch.zeros((50, 60, 3), dtype=torch.uint8)  # a tensor is used in place (on a GPU box: the resident page)
    det = np.array([[40, 10, 40, 30, 5, 30, 5, 10], [5, 35, 40, 35, 40, 45, 5, 45], [1, 1, 9, 1, 9, 5, 1, 5]], np.float64)
    out, metric = system.OcrSystemTask(text_recognizer=rec).text_recognition(det, page)
    assert rec.calls[0][0] is page
    for k, (o, pts) in enumerate(zip(out, rec.calls[0][1])):
        want = predictors.order_point(det[k])
        assert np.array_equal(pts, want) and np.array_equal(o["bbox"], want) and o["index"] == k + 1
    assert [o["text"] for o in out] == ["t0", "", "t2"]  # an empty crop keeps "" like the reference's exception path
    assert metric["total"] == 3 and set(metric) == {"use_time", "avg_use_time", "total"}
    with pytest.raises(ZeroDivisionError):  # the reference's avg over zero boxes
        system.OcrSystemTask(text_recognizer=rec).text_recognition(np.zeros((0, 8)), page)


def test_table_structure_detection_filters_and_orders_the_layout_tables():
    tsr = _Tsr()
    page = torch.zeros((100, 100, 3), dtype=torch.uint8)
    layout = [{"bbox": [0, 60, 50, 90], "label": "Table", "score": 0.9}, {"bbox": [0, 0, 50, 50], "label": "figure", "score": 0.9},
              {"bbox": [0, 10, 50, 40], "label": "table", "score": 0.2}, {"bbox": [5, 5, 9, 9], "label": "table", "score": 0.19}]
    outputs, metric = system.OcrSystemTask(table_structure_recognizer=tsr).table_structure_detection("p.png", image_full=page, layout_result=layout)
    assert [o[0] for o in outputs] == [[0, 10, 50, 40], [0, 60, 50, 90]] and tsr.calls[0][0] is page
    assert all(set(o[1]) >= {"polygons", "logi"} for o in outputs) and set(metric) == {"use_time"}
    # no layout result: the recogniser sees the whole image, its first result comes back
    result, _ = system.OcrSystemTask(table_structure_recognizer=tsr).table_structure_detection("p.png")
    assert result["inputs"] == "p.png" and result["polygons"].shape == (2, 8)
    assert system.get_layout_by_type(layout, label="figure", score_threshold=0.8) == [layout[1]]


def test_layout_analysis_returns_first_result():
    class _Lay:
        def __call__(self, image):
            return [[{"bbox": [1, 2, 3, 4], "label": "text", "score": 0.7}], "second"]

    result, metric = system.OcrSystemTask(layout_detector=_Lay()).layout_analysis("p")
    assert result[0]["label"] == "text" and set(metric) == {"use_time"}


def test_order_points_batch_equals_order_point():
    """The vectorised corner ordering the batched orchestrator uses must give the per-quad reference function's result bit for bit."""
    import os

    rng = np.random.default_rng(5)
    quads = [rng.uniform(0, 900, (200, 4, 2)), rng.integers(0, 60, (200, 4, 2)).astype(np.float64),  # integer grids: ties in the angles
             np.load(os.path.join(os.path.dirname(__file__), "golden", "glue.npz"))["quads"]]
    for q in quads:
        got = predictors.order_points_batch(q)
        assert got.dtype == np.float32 and got.shape == (len(q), 4, 2)
        for k in range(len(q)):
            assert np.array_equal(got[k], predictors.order_point(q[k])), k
    assert predictors.order_points_batch(np.zeros((0, 8))).shape == (0, 4, 2)
