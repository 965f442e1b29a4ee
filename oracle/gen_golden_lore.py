"""Generate the Lore golden fixtures by running the REFERENCE's own modules (build container only).

    python -m oracle.gen_golden_lore          # writes tests/golden/lore_*.npz

  lore_dla34_seed0.npz : get_dla_dcn(34, heads) (lore/lore_dla_34.py:193) with the seeded synthetic state_dict.
  lore_decode.npz      : process_detect_output (lore/lineless_table_process.py:592) on planted head maps.  The
                         function hard-codes `.cuda()` and needs shapely; for this run `.cuda()` is neutralised and
                         shapely's Point/Polygon are replaced by the strict point-in-polygon restatement of
                         oracle/lore_decode_ref.py (third-party predicate -> parity unpinned for it).
  lore_processor_seed0.npz : LoreProcessModel (lore/lore_processor.py:399) fp32 on seeded features.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import lore_decode_ref, ref_import
from pdf_table_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
HEADS = {"hm": 2, "st": 8, "wh": 8, "ax": 256, "cr": 256, "reg": 2}
# (name, synth index, map h, map w, (source h, source w))
DECODE_CASES = [("t0", 0, 128, 128, (600, 800)), ("t1", 1, 128, 128, (1024, 1024)), ("t2", 2, 96, 160, (333, 517)),
                ("t3", 3, 256, 256, (1500, 1100))]


def gen_network():
    ref_import.setup()
    from pdftable.model.lore.lore_dla_34 import get_dla_dcn

    m = get_dla_dcn(34, HEADS, head_conv=256).eval()
    sd = synth.lore_dla34_state_dict(0)
    r = m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    assert not r.unexpected_keys and all(k.startswith("base.fc") or "num_batches" in k for k in r.missing_keys), r
    rng = np.random.default_rng(5)
    x = rng.standard_normal((1, 3, 64, 96)).astype(np.float32)
    with torch.no_grad():
        out = m(torch.from_numpy(x))[0]
    np.savez_compressed(os.path.join(GOLDEN, "lore_dla34_seed0.npz"), x=x, **{k: v.numpy() for k, v in out.items()})
    print("lore_dla34_seed0", {k: tuple(v.shape) for k, v in out.items()})


class _Pt:
    def __init__(self, xy):
        self.xy = (float(xy[0]), float(xy[1]))

    def within(self, poly):
        return lore_decode_ref.point_strictly_in_polygon(self.xy[0], self.xy[1], poly.pts)


class _Poly:
    def __init__(self, pts):
        self.pts = np.asarray([[float(p[0]), float(p[1])] for p in pts], np.float64)


def reference_meta(src_h, src_w, inp=None):
    """TableLorePreProcessor.process / update_meta (lore/processer_lore.py:66-130), wtw (not upper_left)."""
    c = np.array([src_w / 2.0, src_h / 2.0], dtype=np.float32)
    s = max(src_h, src_w) * 1.0
    meta = [c[0], c[1], s, inp[0], inp[1], inp[0] // 4, inp[1] // 4]
    return torch.from_numpy(np.array(meta)).long().unsqueeze(0)


def gen_decode():
    ref_import.setup()
    import pdftable.model.lore.lineless_table_process as L

    L.Point, L.Polygon = _Pt, _Poly
    torch.Tensor.cuda = lambda self, *a, **k: self  # the reference hard-codes .cuda(); this container has no GPU
    out = {}
    for name, idx, h, w, (src_h, src_w) in DECODE_CASES:
        maps = synth.lore_planted_maps(idx, h, w)
        meta = reference_meta(src_h, src_w, (4 * h, 4 * w))
        t = {k: torch.from_numpy(v)[None].clone() for k, v in maps.items()}
        # the engine's contract is the post-sigmoid map: make sigmoid_() a no-op for the planted probabilities
        orig = torch.Tensor.sigmoid_
        torch.Tensor.sigmoid_ = lambda self: self
        try:
            logi_feat, dets_feat, results, corner_st = L.process_detect_output(t, meta, upper_left=False, wiz_rev=True, vis_thresh=0.2)
        finally:
            torch.Tensor.sigmoid_ = orig
        n = logi_feat.shape[1]
        out[name + "_meta"] = meta.numpy()[0]
        out[name + "_logi_feat"] = logi_feat.numpy()[0]
        out[name + "_dets_feat"] = dets_feat.numpy()[0]
        out[name + "_results"] = results[1][:, :9]
        print(name, "cells", n, "rows", results[1].shape, "score>0:", int((results[1][:, 8] > 0).sum()))
    np.savez_compressed(os.path.join(GOLDEN, "lore_decode.npz"), **out)


def gen_processor():
    ref_import.setup()
    from pdftable.model.lore.configuration_lore import LoreConfig
    from pdftable.model.lore.lore_processor import LoreProcessModel

    m = LoreProcessModel(LoreConfig(task_type="wtw")).eval()
    sd = synth.lore_processor_state_dict(0)
    r = m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    assert not r.unexpected_keys and all(k.endswith("pe.pe") for k in r.missing_keys), r
    out = {}
    for n in (1, 7, 64, 200):
        rng = np.random.default_rng(90 + n)
        feat = rng.standard_normal((1, n, 256)).astype(np.float32)
        with torch.no_grad():
            logic, stacked = m(torch.from_numpy(feat))
        out[f"n{n}_feat"] = feat[0]
        out[f"n{n}_logic"] = logic.numpy()[0]
        out[f"n{n}_stacked"] = stacked.numpy()[0]
        print("processor n", n, float(stacked.min()), float(stacked.max()))
    np.savez_compressed(os.path.join(GOLDEN, "lore_processor_seed0.npz"), **out)


def gen_preprocess():
    """TableLorePreProcessor.__call__ (lore/processer_lore.py:132-160) on synthetic pages of several aspect ratios."""
    import sys
    import types

    ref_import.setup()
    m = types.ModuleType("pdftable.utils.ocr")  # drawing helpers only (save_result); not on the tensor path
    m.OcrCommonUtils = type("OcrCommonUtils", (), {})
    sys.modules["pdftable.utils.ocr"] = m
    from pdftable.model.lore.configuration_lore import LoreConfig
    from pdftable.model.lore.processer_lore import TableLorePreProcessor

    pre = TableLorePreProcessor(LoreConfig(task_type="wtw"))
    out = {"sizes": np.array([(600, 800), (1500, 1100), (333, 517), (1024, 1024)])}
    for i, (h, w) in enumerate(out["sizes"]):
        item = pre(synth.synthetic_page(3, int(h), int(w)))[0]
        px = item["pixel_values"].numpy()[0]
        out[f"meta{i}"] = item["meta"].numpy()[0]
        out[f"patch{i}"] = px[:, 480:544, 480:544].copy()
        out[f"sum{i}"] = np.array([px.astype(np.float64).sum(), np.abs(px.astype(np.float64)).sum()])
    np.savez_compressed(os.path.join(GOLDEN, "lore_pre.npz"), **out)
    print("lore_pre", [out[f"meta{i}"].tolist() for i in range(4)])


if __name__ == "__main__":
    gen_preprocess()
    gen_network()
    gen_processor()
    gen_decode()
