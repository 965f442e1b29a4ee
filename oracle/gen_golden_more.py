"""More reference-generated fixtures (build container only; see gen_golden.py)."""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

from . import ref_import
from pdf_table_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _ref_rec_processors():
    """The reference's OCRRecognitionPreprocessor / PostProcessor with a stand-in config object (the real
    configuration_ocr_recognition.py cannot be imported under transformers>=5)."""
    ref_import.setup()
    name = "pdftable.model.ocr_recognition.configuration_ocr_recognition"
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.OCRRecognitionConfig = type("OCRRecognitionConfig", (), {})
        sys.modules[name] = m
    from pdftable.model.ocr_recognition.processor_ocr_recognition import (OCRRecognitionPostProcessor,
                                                                          OCRRecognitionPreprocessor)
    return OCRRecognitionPreprocessor, OCRRecognitionPostProcessor


def gen_convnextvit():
    """Reference ConvNextViT (convnext_vit/modeling_convnext_vit.py:20) + its pre/post-processors on three
    synthetic text-line crops (32x320, 32x200 and a 48x700 crop that exercises resize + all three chunks)."""
    import cv2

    Pre, Post = _ref_rec_processors()
    cfg = types.SimpleNamespace(do_chunking=True, img_height=32, img_width=804)
    pre = Pre(cfg)
    crops = [synth.synthetic_text_crop(0, 32, 320), synth.synthetic_text_crop(1, 32, 200),
             cv2.resize(np.tile(synth.synthetic_text_crop(2, 32, 320), (1, 2, 1)), (700, 48))]
    chunks = pre(crops)["image"]
    model = ref_import.convnext_vit()
    sd = synth.convnext_vit_state_dict(0)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    with torch.no_grad():
        feats = model.cnn_model(chunks[:, 0:1] * 0.2989 + chunks[:, 1:2] * 0.5870 + chunks[:, 2:3] * 0.1140).last_hidden_state
        logits = model(chunks).logits
    # post-processor semantics: emulate its loop without a vocab file (label_mapping is an id -> char dict)
    post = Post.__new__(Post)
    post.label_mapping = {i: chr(0x4E00 + i) for i in range(1, 7644)}
    preds = post(logits)["preds"]
    ids = [[ord(ch) - 0x4E00 for ch in s] for s in preds]
    out = {
        "n_crops": np.int32(len(crops)),
        "chunks_sum": chunks.double().sum(dim=(1, 2, 3)).numpy(),
        "chunk0": chunks[0].numpy(),
        "feats0": feats[0].numpy().astype(np.float32),                   # chunk 0 of [9,512,1,75]
        "feats_abs_sum": feats.abs().double().sum(dim=(1, 2, 3)).numpy(),
        "logits_sub": logits[:, :, ::32].numpy().astype(np.float32),     # every 32nd class
        "logits_max": logits.max(-1).values.numpy().astype(np.float32),
        "argmax": logits.argmax(-1).numpy().astype(np.int32),
    }
    for i, (c, row) in enumerate(zip(crops, ids)):
        out[f"crop{i}"] = c
        out[f"ids{i}"] = np.array(row, np.int32)
    np.savez_compressed(os.path.join(GOLDEN, "convnextvit_seed0.npz"), **out)
    print("convnextvit_seed0", logits.shape, [len(r) for r in ids], float(logits.abs().max()))


def _ref_db_post():
    """The reference's DBPostProcess / PPOcrDetectionPostProcessor (db_pp/processor_ocr_db_pp.py:148-386) with the two
    missing wheels replaced by the restatements in oracle/db_post_ref.py (pyclipper offset, shapely area/length)."""
    ref_import.setup()
    from . import db_post_ref

    sys.modules["pyclipper"] = db_post_ref.PyclipperStandIn
    sg = sys.modules.get("shapely.geometry") or types.ModuleType("shapely.geometry")
    sg.Polygon = db_post_ref.ShapelyPolygonStandIn
    sys.modules["shapely.geometry"] = sg
    name = "pdftable.model.db_pp.configuration_db_pp"
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.DbPPConfig = type("DbPPConfig", (), {})
        sys.modules[name] = m
    import pdftable.model.db_pp.processor_ocr_db_pp as mod

    mod.pyclipper = db_post_ref.PyclipperStandIn
    mod.Polygon = db_post_ref.ShapelyPolygonStandIn
    return mod


DB_POST_CASES = [
    # name, map index, map h, map w, n_lines, src_h, src_w
    ("page960", 0, 960, 960, 40, 960, 960),
    ("page960b", 1, 960, 960, 60, 960, 960),
    ("scaled", 2, 640, 960, 30, 1000, 1500),
    ("small", 3, 160, 224, 6, 160, 224),
    ("tall", 4, 960, 320, 12, 2875, 960),
]


def gen_db_post():
    """Reference DB post-process on planted probability maps (synth.synthetic_prob_map): det_polygons per case."""
    mod = _ref_db_post()
    cfg = types.SimpleNamespace(thresh=0.2, box_thresh=0.6, unclip_ratio=1.5, use_dilation=False, score_mode="fast",
                                max_candidates=1000)
    post = mod.PPOcrDetectionPostProcessor(cfg)
    out = {}
    for name, idx, h, w, n_lines, src_h, src_w in DB_POST_CASES:
        prob = synth.synthetic_prob_map(idx, h, w, n_lines)
        shape_list = np.array([src_h, src_w, h / float(src_h), w / float(src_w)])
        res = post({"results": torch.from_numpy(prob)[None, None], "org_shape": (src_h, src_w, 3),
                    "shape_list": shape_list, "inputs": None})
        out[name] = res["det_polygons"].astype(np.float32)
        print("db_post", name, out[name].shape)
    np.savez_compressed(os.path.join(GOLDEN, "db_post.npz"), **out)


def gen_det_pre():
    """Reference DetResizeForTest + NormalizeImage + ToCHWImage (db_pp/image_operators.py:78-118, 212-316) on odd page
    sizes: resized shapes / ratios, and the normalised tensor of one small page."""
    ref_import.setup()
    from pdftable.model.db_pp.image_operators import DetResizeForTest, NormalizeImage, ToCHWImage

    op = DetResizeForTest(limit_side_len=960, limit_type="max")
    shapes = [(960, 960), (1000, 1500), (2875, 960), (700, 500), (33, 47), (20, 30), (1111, 1111), (480, 1919)]
    rows = []
    for h, w in shapes:
        img = (np.arange(h * w * 3, dtype=np.int64) % 251).astype(np.uint8).reshape(h, w, 3)
        d = op({"image": img})
        rows.append([h, w, d["image"].shape[0], d["image"].shape[1], d["shape"][2], d["shape"][3]])
    page = synth.synthetic_page(5, 100, 150)
    d = op({"image": page[:, :, ::-1]})
    d = NormalizeImage(scale=1.0 / 255.0, mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225], order="hwc")(d)
    d = ToCHWImage()(d)
    np.savez_compressed(os.path.join(GOLDEN, "det_pre.npz"), table=np.array(rows, np.float64), page_chw=d["image"].astype(np.float32),
                        page_shape=np.array(d["shape"], np.float64))
    print("det_pre", np.array(rows)[:, 2:4].tolist(), d["image"].shape)


GENERATORS = {"convnextvit": gen_convnextvit, "db_post": gen_db_post, "det_pre": gen_det_pre}


def _ref_dbnet_proc():
    """The reference's in-tree DBNet processors (db_net/processor_ocr_dbnet.py, db_net/ocr_detection_utils.py) with a stand-in
    config module and the two missing wheels replaced by the restatements of oracle/db_post_ref.py."""
    ref_import.setup()
    from . import db_post_ref

    sys.modules["pyclipper"] = db_post_ref.PyclipperStandIn
    sg = sys.modules.get("shapely.geometry") or types.ModuleType("shapely.geometry")
    sg.Polygon = db_post_ref.ShapelyPolygonStandIn
    sys.modules["shapely.geometry"] = sg
    name = "pdftable.model.db_net.configuration_dbnet"
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.DbNetConfig = type("DbNetConfig", (), {})
        sys.modules[name] = m
    import pdftable.model.db_net.ocr_detection_utils as utils
    import pdftable.model.db_net.processor_ocr_dbnet as mod

    utils.pyclipper = db_post_ref.PyclipperStandIn
    utils.Polygon = db_post_ref.ShapelyPolygonStandIn
    return mod


DBNET_POST_CASES = [
    # name, map index, map h, map w, n_lines, org_h, org_w
    ("same", 0, 736, 992, 40, 736, 992),
    ("scaled", 2, 640, 960, 30, 1000, 1500),
    ("small", 3, 160, 224, 6, 160, 224),
    ("tall", 4, 960, 320, 12, 2875, 960),
]


def gen_dbnet_proc():
    """Reference OCRDetectionPreprocessor / OCRDetectionPostProcessor (the model="db" back-end): resize rule on odd page sizes,
    the normalised tensor of one small page, and det_polygons on planted probability maps."""
    mod = _ref_dbnet_proc()
    cfg = types.SimpleNamespace(img_width=736, thresh=0.2, return_polygon=False)
    pre, post = mod.OCRDetectionPreprocessor(cfg), mod.OCRDetectionPostProcessor(cfg)
    mod.logger.info = lambda *a, **k: None
    shapes = [(960, 960), (1000, 1500), (2875, 960), (700, 500), (480, 1919), (736, 736), (100, 150)]
    rows = []
    for h, w in shapes:
        img = (np.arange(h * w * 3, dtype=np.int64) % 251).astype(np.uint8).reshape(h, w, 3)
        r = pre.resize(img)
        rows.append([h, w, r.shape[0], r.shape[1]])
    page = synth.synthetic_page(5, 100, 150)
    d = pre(page)
    out = {"table": np.array(rows, np.int64), "page_chw": d["image"].numpy().astype(np.float32), "page_org_shape": np.array(d["org_shape"], np.int64)}
    for name, idx, h, w, n_lines, org_h, org_w in DBNET_POST_CASES:
        prob = synth.synthetic_prob_map(idx, h, w, n_lines)
        res = post({"results": torch.from_numpy(prob)[None, None], "org_shape": [org_h, org_w]})
        out["post_" + name] = np.asarray(res["det_polygons"]).reshape(-1, 8).astype(np.int64)
        print("dbnet_proc", name, out["post_" + name].shape)
    np.savez_compressed(os.path.join(GOLDEN, "dbnet_proc.npz"), **out)
    print("dbnet_proc", np.array(rows)[:, 2:].tolist(), d["image"].shape)


GENERATORS["dbnet_proc"] = gen_dbnet_proc
