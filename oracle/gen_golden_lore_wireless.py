"""Generate tests/golden/lore_resnet18_seed0.npz by running the REFERENCE's own LoreDetectModel (build container only).

    python -m oracle.gen_golden_lore_wireless

The module (lore/lore_detector.py:148-389) is loaded with the seeded synthetic state_dict of
pdf_table_b200.synth.lore_resnet18_state_dict and run on one seeded 64 x 128 input; all six head outputs are stored.
Also in the file, for the configuration's upper-left-anchored frame: TableLorePreProcessor (lore/processer_lore.py:66-160)
with LoreConfig(task_type="wireless") on synthetic pages (meta, a pixel patch, sums), and process_detect_output
(lore/lineless_table_process.py:592) with upper_left=True, wiz_rev=False on planted maps (the same `.cuda()` / shapely
substitutions as oracle/gen_golden_lore.py).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ref_import
from pdf_table_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    ref_import.setup()
    from pdftable.model.lore.lore_detector import LoreDetectModel

    m = LoreDetectModel().eval()
    sd = synth.lore_resnet18_state_dict(0)
    r = m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    assert not r.unexpected_keys and all("num_batches" in k for k in r.missing_keys), r
    rng = np.random.default_rng(11)
    x = rng.standard_normal((1, 3, 64, 128)).astype(np.float32)
    with torch.no_grad():
        out = m(torch.from_numpy(x))[0]
    extra = {}
    # ---- pre-processor, wireless configuration
    import sys
    import types

    mod = types.ModuleType("pdftable.utils.ocr")  # drawing helpers only (save_result); not on the tensor path
    mod.OcrCommonUtils = type("OcrCommonUtils", (), {})
    sys.modules["pdftable.utils.ocr"] = mod
    from pdftable.model.lore.configuration_lore import LoreConfig
    from pdftable.model.lore.processer_lore import TableLorePreProcessor

    pre = TableLorePreProcessor(LoreConfig(task_type="wireless"))
    extra["pre_sizes"] = np.array([(600, 800), (1500, 1100), (333, 517)])
    for i, (h, w) in enumerate(extra["pre_sizes"]):
        item = pre(synth.synthetic_page(3, int(h), int(w)))[0]
        px = item["pixel_values"].numpy()[0]
        extra[f"pre_meta{i}"] = item["meta"].numpy()[0]
        extra[f"pre_patch{i}"] = px[:, 100:164, 200:264].copy()
        extra[f"pre_sum{i}"] = np.array([px.astype(np.float64).sum(), np.abs(px.astype(np.float64)).sum()])
    # ---- decode in the upper-left frame
    from .gen_golden_lore import _Poly, _Pt
    import pdftable.model.lore.lineless_table_process as L

    L.Point, L.Polygon = _Pt, _Poly
    torch.Tensor.cuda = lambda self, *a, **k: self
    maps = synth.lore_planted_maps(4, 192, 192)
    meta = torch.from_numpy(np.array([0, 0, 900.0, 768, 768, 192, 192])).long().unsqueeze(0)
    t = {k: torch.from_numpy(v)[None].clone() for k, v in maps.items()}
    orig = torch.Tensor.sigmoid_
    torch.Tensor.sigmoid_ = lambda self: self
    try:
        logi_feat, dets_feat, results, _ = L.process_detect_output(t, meta, upper_left=True, wiz_rev=False, vis_thresh=0.2)
    finally:
        torch.Tensor.sigmoid_ = orig
    extra["dec_meta"] = meta.numpy()[0]
    extra["dec_logi_feat"] = logi_feat.numpy()[0]
    extra["dec_dets_feat"] = dets_feat.numpy()[0]
    extra["dec_results"] = results[1][:, :9]
    print("decode cells", logi_feat.shape[1])
    np.savez_compressed(os.path.join(GOLDEN, "lore_resnet18_seed0.npz"), x=x, **{k: v.numpy() for k, v in out.items()}, **extra)
    print("lore_resnet18_seed0", {k: (tuple(v.shape), float(v.abs().max())) for k, v in out.items()})


if __name__ == "__main__":
    main()
