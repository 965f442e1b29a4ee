"""Host-side logic of the predictor mirror (no GPU): resize rules vs the reference's golden numbers, error behaviour."""
import os

import numpy as np
import pytest

from pdf_table_b200 import predictors, synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_det_resize_matches_reference():
    g = np.load(os.path.join(GOLDEN, "det_pre.npz"))
    for h, w, rh, rw, ratio_h, ratio_w in g["table"]:
        h, w = int(h), int(w)
        img = (np.arange(h * w * 3, dtype=np.int64) % 251).astype(np.uint8).reshape(h, w, 3)
        out, (a, b) = predictors.det_resize_for_test(img, 960, "max")
        assert out.shape[:2] == (int(rh), int(rw))
        assert a == ratio_h and b == ratio_w
    # pixel values of the resized + normalised page (the GPU kernel's input/normalisation contract)
    page = synth.synthetic_page(5, 100, 150)
    res, _ = predictors.det_resize_for_test(page, 960, "max")
    mean = np.array([0.485, 0.456, 0.406], np.float32).reshape(1, 1, 3)
    std = np.array([0.229, 0.224, 0.225], np.float32).reshape(1, 1, 3)
    x = (res[:, :, ::-1].astype("float32") * np.float32(1.0 / 255.0) - mean) / std
    np.testing.assert_array_equal(x.transpose(2, 0, 1), g["page_chw"])


def test_keepratio_resize_rules():
    from oracle import convnextvit_ref

    for h, w in [(32, 320), (48, 700), (20, 900), (64, 64), (32, 804), (10, 400)]:
        crop = (np.arange(h * w * 3) % 255).astype(np.uint8).reshape(h, w, 3)
        got = predictors.keepratio_resize(crop)
        want = convnextvit_ref.keepratio_resize(crop)  # reference restatement incl. zero pad to 804
        assert got.shape[0] == 32 and got.shape[1] <= 804
        np.testing.assert_array_equal(want[:, : got.shape[1]], got)
        assert (want[:, got.shape[1]:] == 0).all()


def test_lore_preprocess_matches_reference():
    """lore_preprocess (uint8 warp + meta) followed by numpy's normalisation == TableLorePreProcessor's pixel_values."""
    g = np.load(os.path.join(GOLDEN, "lore_pre.npz"))
    mean = np.array([0.408, 0.447, 0.470], dtype=np.float32).reshape(1, 1, 3)
    std = np.array([0.289, 0.274, 0.278], dtype=np.float32).reshape(1, 1, 3)
    for i, (h, w) in enumerate(g["sizes"]):
        warped, meta = predictors.lore_preprocess(synth.synthetic_page(3, int(h), int(w)))
        np.testing.assert_array_equal(meta, g[f"meta{i}"])
        x = ((warped / 255. - mean) / std).astype(np.float32).transpose(2, 0, 1)
        np.testing.assert_array_equal(x[:, 480:544, 480:544], g[f"patch{i}"])
        np.testing.assert_allclose([x.astype(np.float64).sum(), np.abs(x.astype(np.float64)).sum()], g[f"sum{i}"], rtol=1e-12)


def test_lore_wireless_preprocess_matches_reference():
    """upper_left=True (LoreConfig wireless, processer_lore.py:74-79): the warp anchored at the upper-left corner, 768 x 768."""
    from oracle import lore_decode_ref

    g = np.load(os.path.join(GOLDEN, "lore_resnet18_seed0.npz"))
    mean = np.array([0.408, 0.447, 0.470], dtype=np.float32).reshape(1, 1, 3)
    std = np.array([0.289, 0.274, 0.278], dtype=np.float32).reshape(1, 1, 3)
    for i, (h, w) in enumerate(g["pre_sizes"]):
        warped, meta = predictors.lore_preprocess(synth.synthetic_page(3, int(h), int(w)), (768, 768), upper_left=True)
        np.testing.assert_array_equal(meta, g[f"pre_meta{i}"])
        x = ((warped / 255. - mean) / std).astype(np.float32).transpose(2, 0, 1)
        np.testing.assert_array_equal(x[:, 100:164, 200:264], g[f"pre_patch{i}"])
        np.testing.assert_allclose([x.astype(np.float64).sum(), np.abs(x.astype(np.float64)).sum()], g[f"pre_sum{i}"], rtol=1e-12)
    for c, sc in (((0, 0), 900.0), ((3, 7), 500.0), ((9, 2), 640.0)):
        for inv in (False, True):
            np.testing.assert_array_equal(predictors.lore_affine_upper_left(c, sc, 192, 192, inv),
                                          lore_decode_ref.upper_left_matrix(np.float32(c), np.float32(sc), 192, 192, inv))


def test_picodet_preprocess_matches_reference():
    """cv2.resize on the un-flipped page, then flip + (x * scale - mean) / std in fp32 == OCRPicodetPreProcessor's image."""
    import cv2

    g = np.load(os.path.join(GOLDEN, "picodet_net_seed0.npz"))
    page = synth.synthetic_page(9, 500, 380)
    res = cv2.resize(page, (608, 800))
    mean = np.array([0.485, 0.456, 0.406], np.float32).reshape(1, 1, 3)
    std = np.array([0.229, 0.224, 0.225], np.float32).reshape(1, 1, 3)
    x = ((res[:, :, ::-1].astype("float32") * np.float32(1.0 / 255.0) - mean) / std).transpose(2, 0, 1)
    np.testing.assert_array_equal(x[:, 300:332, 200:232], g["pre_patch"])
    np.testing.assert_allclose([x.astype(np.float64).sum(), np.abs(x.astype(np.float64)).sum()], g["pre_sum"], rtol=1e-12)
    np.testing.assert_array_equal(g["pre_meta"], [500, 380, 800 / 500, 608 / 380, 800, 608])


def test_error_behaviour_without_gpu():
    import torch

    with pytest.raises(RuntimeError):
        predictors.OcrDetectionTask(model="east", state_dict={})
    with pytest.raises(RuntimeError):
        predictors.OcrRecognitionTask(model="LightweightEdge", state_dict={})
    with pytest.raises(RuntimeError):  # CRNN has no fp32x mode
        predictors.OcrRecognitionTask(model="CRNN", precision="fp32x", state_dict={})
    with pytest.raises(RuntimeError):
        predictors.OcrTableStructureTask(model="CenterNet", state_dict=({}, {}))
    with pytest.raises(RuntimeError):
        predictors.OcrTableStructureTask(model="Lore", task_type="ptn", state_dict=({}, {}))
    with pytest.raises(RuntimeError):
        predictors.OcrLayoutTask(model="DocXLayout", state_dict=({}, {}, {}))
    with pytest.raises(TypeError):
        predictors._read_image(12345)
    if not torch.cuda.is_available():
        from pdf_table_b200._lib import DocVisionError

        with pytest.raises(DocVisionError):  # no silent CPU fallback
            predictors.OcrDetectionTask(model="db_pp", state_dict=synth.dbnet_r18_state_dict(0))


def test_dbnet_backend_preprocess_matches_reference():
    """model="db": OCRDetectionPreprocessor's resize rule and its (x - mean) / 255 normalisation of the BGR-flipped page
    (tests/golden/dbnet_proc.npz, generated from the reference class by oracle/gen_golden_more.py:gen_dbnet_proc).  The second
    half is the arithmetic contract of the fused kernel: (x * 1 - mean) / 255 in float32."""
    import cv2

    g = np.load(os.path.join(GOLDEN, "dbnet_proc.npz"))
    for h, w, rh, rw in g["table"]:
        assert predictors.dbnet_resize_shape(int(h), int(w), 736) == (int(rh), int(rw))
    page = synth.synthetic_page(5, 100, 150)
    assert list(g["page_org_shape"]) == [100, 150]
    nh, nw = predictors.dbnet_resize_shape(100, 150, 736)
    res = cv2.resize(page, (nw, nh))  # resize before the flip == flip before the resize (cv2.resize acts per channel)
    mean = np.array(predictors.OcrDetectionTask.DB_MEAN, np.float32)
    x = (res[:, :, ::-1].astype(np.float32) * np.float32(1.0) - mean) / np.float32(255.0)
    np.testing.assert_array_equal(x.transpose(2, 0, 1), g["page_chw"])


def test_ids_to_texts_table_lookup_equals_the_per_token_join():
    """OcrRecognitionTask._ids_to_texts: the one-look-up path (single-character dictionary, 'blank' word at id 0, rows padded with
    -1 beyond their length) returns exactly what CTCLabelDecode's per-token join returns -- also for a row that shows id 0 inside
    its length (never produced by the collapse, but the mapping must not depend on that) and for a multi-character dictionary."""
    class _T:
        pass

    f = predictors.OcrRecognitionTask._ids_to_texts
    rng = np.random.default_rng(0)
    for chars in ([chr(0x4E00 + i) for i in range(60)], ["ab", "c"] + [chr(0x4E00 + i) for i in range(58)]):
        t = _T()
        t.character = ["blank"] + chars + [" "]
        ids = rng.integers(1, len(t.character), (200, 25)).astype(np.int32)
        lens = rng.integers(0, 26, 200).astype(np.int32)
        for i in range(200):
            ids[i, lens[i]:] = -1
        ids[7, 0] = 0
        lens[7] = max(int(lens[7]), 3)
        ids[7, :3] = [0, 5, 0]
        want = ["".join(t.character[int(v)] for v in row[:k]) for row, k in zip(ids, lens)]
        assert f(t, ids, lens) == want
    t = _T()
    t.character = None
    assert f(t, np.array([[3, 4, -1]], np.int32), np.array([2], np.int32)) == ["3 4"]


def test_packed_narrow_pointwise_weights():
    """pp_rec_graph.pw_pack_factor / block_diagonal: a 16- (32-) channel 1x1 layer is packed 4 (2) pixels per GEMM row; the packed
    product over a [M / p, p * K] view equals the layer applied per pixel."""
    from pdf_table_b200 import pp_rec_graph as G

    class _B:
        precise = False

    assert G.pw_pack_factor(_B(), 16, 32) == 4 and G.pw_pack_factor(_B(), 32, 64) == 2 and G.pw_pack_factor(_B(), 64, 64) == 1
    assert G.pw_pack_factor(_B(), 16, 128) == 1  # 4 x 128 output columns would not fit one n-tile
    pb = _B()
    pb.precise = True
    assert G.pw_pack_factor(pb, 16, 32) == 1
    rng = np.random.default_rng(1)
    w, b = rng.standard_normal((32, 16)).astype(np.float32), rng.standard_normal(32).astype(np.float32)
    wd, bd = G.block_diagonal(w, b, 4)
    x = rng.standard_normal((40, 16)).astype(np.float32)
    np.testing.assert_allclose((x.reshape(10, 64) @ wd.T + bd).reshape(40, 32), x @ w.T + b, rtol=1e-5, atol=1e-5)
