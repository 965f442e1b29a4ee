"""The batched page loop (system.OcrSystemTask.predict_pages: layout + detection + recognition + table structure for a batch of
pages with overlapped host steps) must return exactly what the reference-shaped per-page calls return on the same inputs, and
the in-tree DBNet back-end (model="db") must follow the reference's own pre / post-processing."""
import os

import numpy as np
import pytest
import torch

from oracle import db_post_ref, dbnet_ref
from pdf_table_b200 import predictors, synth, system

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TABLE = [20.4, 60.6, 600.5, 420.3]


@pytest.fixture(scope="module")
def cascade():
    vocab = [chr(0x4E00 + i) for i in range(2, 7644)]
    det = predictors.OcrDetectionTask(model="db_pp", state_dict=synth.dbnet_r18_state_dict(0))
    rec = predictors.OcrRecognitionTask(model="ConvNextViT", state_dict=synth.convnext_vit_state_dict(0), vocab=vocab)
    lay = predictors.OcrLayoutTask(model="picodet", task_type="en", state_dict=synth.picodet_state_dicts(0, 5))
    sd = synth.lore_dla34_state_dict(0)
    sd["hm.2.bias"] = np.array([-0.3, -3.5], np.float32)
    tsr = predictors.OcrTableStructureTask(model="Lore", task_type="wtw", state_dict=(sd, synth.lore_processor_state_dict(0)))
    return system.OcrSystemTask(text_detector=det, text_recognizer=rec, table_structure_recognizer=tsr, layout_detector=lay)


def test_predict_pages_equals_the_per_page_calls(cascade):
    pages = np.stack([synth.synthetic_page(20 + i, 480, 640) for i in range(3)])
    planted = torch.from_numpy(np.stack([synth.synthetic_prob_map(30 + i, 480, 640, 14) for i in range(3)])[:, None]).cuda()
    tables = [[TABLE], [], [TABLE, [100.0, 200.0, 500.0, 470.0]]]
    out = cascade.predict_pages(pages, layout_tables=tables, det_kwargs={"prob_override": lambda prob, idx: planted[idx]},
                                keep_device_record=True)
    assert len(out) == 3 and sum(len(p["det"]) for p in out) > 20
    rec = cascade.device_record
    assert rec["box_counts"].cpu().tolist() == [len(p["det"]) for p in out] and int(rec["ids"].shape[0]) == sum(len(p["det"]) for p in out)
    assert int(rec["cell_counts"].shape[0]) == 3 and tuple(rec["cell_logi"].shape[1:]) == (256, 4)
    for p in range(3):
        # text_detection / text_recognition of the reference's orchestrator on this page alone
        det = cascade.text_detector(pages[p], prob_override=lambda prob, idx, p=p: planted[p:p + 1])[0]
        det = predictors.sort_det_boxes(det) if len(det) else np.zeros((0, 8))
        np.testing.assert_array_equal(out[p]["det"], det)
        ocr, _ = cascade.text_recognition(det, pages[p])
        assert [o["text"] for o in ocr] == [o["text"] for o in out[p]["ocr"]]
        for a, b in zip(ocr, out[p]["ocr"]):
            assert np.array_equal(a["bbox"], b["bbox"]) and a["index"] == b["index"]
        # layout and the table loop
        lay, _ = cascade.layout_analysis(pages[p])
        assert len(lay) == len(out[p]["layout"])
        for a, b in zip(lay, out[p]["layout"]):
            assert a["label"] == b["label"] and a["score"] == b["score"] and np.array_equal(a["bbox"], b["bbox"])
        want, _ = cascade.table_structure_detection(pages[p], image_full=pages[p],
                                                    layout_result=[{"bbox": b, "label": "table", "score": 0.9} for b in tables[p]])
        want = {tuple(w[0]): w[1] for w in (want if tables[p] else [])}
        assert len(out[p]["tables"]) == len(tables[p])
        for bbox, res in out[p]["tables"]:
            np.testing.assert_array_equal(res["polygons"], want[tuple(bbox)]["polygons"])
            np.testing.assert_array_equal(res["logi"], want[tuple(bbox)]["logi"])


def test_predict_stream_equals_predict_pages(cascade):
    """predict_stream (batch i + 1 uploaded by a copy stream while batch i computes) yields exactly predict_pages' results, batch
    by batch, from views of one pinned buffer."""
    host = torch.empty((3, 2, 480, 640, 3), dtype=torch.uint8).pin_memory()
    for b in range(3):
        for i in range(2):
            host[b, i] = torch.from_numpy(synth.synthetic_page(40 + 2 * b + i, 480, 640))
    batches = [host[b].numpy() for b in range(3)]
    tables = [[TABLE], [[100.0, 200.0, 500.0, 470.0]]]
    want = [cascade.predict_pages(bt, layout_tables=tables) for bt in batches]
    got = list(cascade.predict_stream(iter(batches), layout_tables=tables))
    assert len(got) == 3
    for w, g in zip(want, got):
        for pw, pg in zip(w, g):
            np.testing.assert_array_equal(pw["det"], pg["det"])
            assert [o["text"] for o in pw["ocr"]] == [o["text"] for o in pg["ocr"]]
            assert len(pw["layout"]) == len(pg["layout"]) and len(pw["tables"]) == len(pg["tables"])
            for (_, a), (_, b) in zip(pw["tables"], pg["tables"]):
                np.testing.assert_array_equal(a["polygons"], b["polygons"])
                np.testing.assert_array_equal(a["logi"], b["logi"])
    assert list(cascade.predict_stream(iter([]))) == []


def test_detection_task_dbnet_backend():
    """OcrDetectionTask(model="db"): the fused pre-process equals OCRDetectionPreprocessor's tensor (golden from the reference
    class) within the fp16 rounding of the stem input, the network sees exactly that input, and the boxes of a planted map
    equal the reference post-processor's (oracle restatement pinned to the reference golden on CPU)."""
    sd = synth.dbnet_r18_state_dict(0)
    task = predictors.OcrDetectionTask(model="db", state_dict=sd, image_short_side=160)
    assert task.box_thresh == 0.3
    page = synth.synthetic_page(5, 100, 150)
    pre = task._preprocess(page)
    assert pre["pages"][0].shape[:2] == predictors.dbnet_resize_shape(100, 150, 160) == (160, 256)
    # network parity on the reference-normalised input
    x = (pre["pages"][0][:, :, ::-1].astype(np.float32) - np.array(task.DB_MEAN, np.float32)) / np.float32(255.0)
    want = dbnet_ref.dbnet_r18_forward(sd, torch.from_numpy(np.ascontiguousarray(x.transpose(2, 0, 1))[None])).numpy()
    mean, std, scale = task._norm()
    got = task.predictor.dbnet_forward_u8(torch.from_numpy(pre["pages"][0][None].copy()).cuda(), mean, std, scale, flip=True).cpu().numpy()
    assert np.abs(got - want).max() < 1e-2
    # boxes: planted map at the resized shape, scaled back to the 100 x 150 page
    prob = synth.synthetic_prob_map(3, 160, 256, 8)
    res = task(page, prob_override=lambda p, idx: torch.from_numpy(prob)[None, None].cuda())
    want_boxes = db_post_ref.dbnet_postprocess(prob, (100, 150))
    assert res[0].dtype == np.int64 and len(want_boxes) > 3
    np.testing.assert_array_equal(res[0], want_boxes)
    # a resident page gives the same result as the host page
    res_dev = task(torch.from_numpy(page).cuda(), prob_override=lambda p, idx: torch.from_numpy(prob)[None, None].cuda())
    np.testing.assert_array_equal(res_dev[0], res[0])
