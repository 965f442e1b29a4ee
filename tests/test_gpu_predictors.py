"""The predictor mirror end to end on the GPU: construct -> __call__ with the reference's return conventions."""
import os

import numpy as np
import pytest
import torch

from oracle import convnextvit_ref as ref
from oracle import lore_decode_ref, lore_processor_ref
from pdf_table_b200 import predictors, synth

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_detection_task_call_and_composition():
    sd = synth.dbnet_r18_state_dict(0)
    task = predictors.OcrDetectionTask(model="db_pp", state_dict=sd, thresh=0.2)
    pages = [synth.synthetic_page(0, 320, 480), synth.synthetic_page(1, 300, 500), synth.synthetic_page(2, 320, 480)]
    res = task(pages)
    assert isinstance(res, list) and len(res) == 3
    for r in res:
        assert r.dtype == np.float32 and r.ndim == 2 and r.shape[1] == 8
    # same result as the explicit composition of the C-ABI calls for one page
    eng = task.predictor
    page, (rh, rw) = predictors.det_resize_for_test(pages[1], 960, "max")
    prob = eng.dbnet_forward_u8(torch.from_numpy(page[None].copy()).cuda(), task.MEAN, task.STD, 1 / 255.0, True)
    boxes, counts = eng.db_boxes(prob, [(300, 500)], 0.2, 0.6, 1.5, 1000)
    np.testing.assert_array_equal(boxes.cpu().numpy()[0, : int(counts[0])], res[1])
    single = task(pages[1])
    np.testing.assert_array_equal(single[0], res[1])


def test_recognition_task_strings_match_reference_golden():
    g = np.load(os.path.join(GOLDEN, "convnextvit_seed0.npz"))
    n = int(g["n_crops"])
    vocab = [chr(0x4E00 + i) for i in range(2, 7644)]  # id i+2 -> chr(0x4E00 + i + 2), the mapping used by gen_golden
    task = predictors.OcrRecognitionTask(model="ConvNextViT", state_dict=synth.convnext_vit_state_dict(0), vocab=vocab)
    crops = [g[f"crop{i}"] for i in range(n)]
    res = task(crops)
    assert isinstance(res, list) and all(isinstance(s, str) for s in res)
    sd = synth.convnext_vit_state_dict(0)
    logits = ref.convnextvit_forward(sd, ref.preprocess(crops))
    top2 = torch.topk(logits, 2, dim=-1).values
    safe = bool(((top2[..., 0] - top2[..., 1]) > 0.05).all())
    for i in range(n):
        want = "".join(chr(0x4E00 + int(v)) for v in g[f"ids{i}"])
        if safe:
            assert res[i] == want
        else:  # a token whose fp32 top-2 margin is inside the fp16 error band may legitimately differ
            assert abs(len(res[i]) - len(want)) <= 3


def test_table_structure_task_lore():
    """OcrTableStructureTask(model="Lore") end to end: return convention, u8 path == fp32 path, the decode of the
    engine's own maps equals the oracle decode bit for bit, and the logical locations follow the processor oracle."""
    sd = synth.lore_dla34_state_dict(0)
    sd["hm.2.bias"] = np.array([-0.3, -3.5], np.float32)  # random weights: shift the heat maps so that ~100 cells / corners pass the gates
    psd = synth.lore_processor_state_dict(0)
    task = predictors.OcrTableStructureTask(model="Lore", task_type="wtw", state_dict=(sd, psd), max_cells_per_image=3000)
    pages = [synth.synthetic_page(7, 700, 900), synth.synthetic_page(8, 1024, 768)]
    res = task(pages)
    assert isinstance(res, list) and len(res) == 2
    for r in res:
        assert set(r) >= {"polygons", "logi", "inputs"}
        assert r["polygons"].dtype == np.float32 and r["polygons"].shape[1] == 8 and r["logi"].shape == (len(r["polygons"]), 4)
        assert np.array_equal(r["logi"], np.floor(r["logi"])) and (r["logi"] >= 0).all()
    assert sum(len(r["polygons"]) for r in res) > 5
    # explicit composition through the C ABI, fp32 input path
    eng, post, proc = task.predictor, task.post, task.processor
    mean = np.array(eng.LORE_MEAN, dtype=np.float32).reshape(1, 1, 3)
    std = np.array(eng.LORE_STD, dtype=np.float32).reshape(1, 1, 3)
    pre = [predictors.lore_preprocess(p) for p in pages]
    x = np.stack([((w / 255. - mean) / std).astype(np.float32).transpose(2, 0, 1) for w, _ in pre])
    maps = eng.lore_detect_forward(torch.from_numpy(x).cuda())
    maps_u8 = eng.lore_detect_forward_u8(torch.from_numpy(np.stack([w for w, _ in pre])).cuda())
    assert torch.equal(maps, maps_u8)  # the fused normalisation is bit-exact w.r.t. numpy's float64 expression
    m = maps.cpu().numpy()
    for i, (_, meta) in enumerate(pre):
        z = np.zeros((1, 256, 256), np.float32)
        want = lore_decode_ref.lore_decode(m[i, :, :, 0:2].transpose(2, 0, 1), m[i, :, :, 2:4].transpose(2, 0, 1),
                                           m[i, :, :, 4:12].transpose(2, 0, 1), m[i, :, :, 12:20].transpose(2, 0, 1), z, z, meta)
        np.testing.assert_array_equal(res[i]["polygons"], want["polygons"])
    # logical locations: processor oracle on the engine's own cell features
    dec = post.lore_decode(maps, None, None, None, np.stack([predictors.lore_affine([np.float32(mm[0]), np.float32(mm[1])], np.float32(mm[2]), 256, 256, True) for _, mm in pre]))
    feat, offsets = eng.lore_cell_features(dec, max_rows=6000, check_overflow=True)
    offs = offsets.cpu().numpy()
    for i in range(2):
        f = feat[offs[i]: offs[i + 1]].cpu()
        if len(f) == 0:
            continue
        _, stacked = lore_processor_ref.lore_processor_forward(psd, f)
        want_logi = lore_decode_ref.round_logic(stacked.numpy())
        safe = np.abs((stacked.numpy() - np.floor(stacked.numpy())) - 0.5) > 2e-3
        np.testing.assert_array_equal(res[i]["logi"][safe], want_logi[safe])
        assert safe.mean() > 0.98


def test_recognition_task_recognize_page_equals_per_crop_calls():
    """det -> rec on the device: recognize_page(page, quads) returns exactly what the reference's flow returns -- crop every
    quad with OcrCommonUtils.crop_image (cv2 on the host) and call the task on the crops."""
    import math

    import cv2

    vocab = [chr(0x4E00 + i) for i in range(2, 7644)]
    task = predictors.OcrRecognitionTask(model="ConvNextViT", state_dict=synth.convnext_vit_state_dict(0), vocab=vocab)
    page = synth.synthetic_page(5, 480, 640)
    rng = np.random.default_rng(3)
    quads = []
    for k in range(12):
        cx, cy, bw, bh, ang = rng.uniform(100, 540), rng.uniform(60, 420), rng.uniform(40, 300), rng.uniform(12, 40), rng.uniform(-0.3, 0.3)
        c, s = math.cos(ang), math.sin(ang)
        quads.append((np.array([[-bw / 2, -bh / 2], [bw / 2, -bh / 2], [bw / 2, bh / 2], [-bw / 2, bh / 2]]) @ np.array([[c, s], [-s, c]])
                      + [cx, cy]).astype(np.float32))
    quads.append(np.array([[10, 10], [10.4, 10], [10.4, 30], [10, 30]], np.float32))  # zero-width crop: the reference's cv2 call raises
    got = task.recognize_page(page, quads)
    assert len(got) == len(quads) and got[-1] is None
    host_crops = []
    for q in quads[:-1]:
        corners, trans, size = predictors.crop_geometry(q)
        host_crops.append(cv2.warpPerspective(page, cv2.getPerspectiveTransform(corners, trans), size))
    assert got[:-1] == task(host_crops)


def test_recognition_task_fp32x_strings_identical_to_reference_golden():
    """precision="fp32x": the strings equal the reference post-processor's on every golden crop, unconditionally."""
    g = np.load(os.path.join(GOLDEN, "convnextvit_seed0.npz"))
    n = int(g["n_crops"])
    vocab = [chr(0x4E00 + i) for i in range(2, 7644)]
    task = predictors.OcrRecognitionTask(model="ConvNextViT", state_dict=synth.convnext_vit_state_dict(0), vocab=vocab, precision="fp32x")
    res = task([g[f"crop{i}"] for i in range(n)])
    assert res == ["".join(chr(0x4E00 + int(v)) for v in g[f"ids{i}"]) for i in range(n)]
    with pytest.raises(RuntimeError):
        predictors.OcrLayoutTask(model="picodet", state_dict=synth.picodet_state_dicts(0, 5), precision="fp32x")
