"""Generate tests/golden/picodet_post.npz by running the REFERENCE's OCRPicodetPostProcessor (build container only).

    python -m oracle.gen_golden_picodet

The processor module imports configuration_picodet, whose PicodetConfig trips transformers>=5's dataclass check
(SURVEY.md section 10); a dummy module is registered in its place and a plain object carrying the same attributes
(strides, thresholds, top-k, id2label) is handed to the processor.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

from . import ref_import
from pdf_table_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
# (name, synth index, classes, (org_h, org_w), objects)
CASES = [("en", 0, 5, (1100, 850), 12), ("ch", 1, 10, (1600, 1200), 20), ("table", 2, 1, (700, 1000), 4), ("empty", 3, 5, (800, 608), 0),
         ("dense", 4, 5, (2000, 1500), 60)]


def reference_post(num_classes):
    ref_import.setup()
    dummy = types.ModuleType("pdftable.model.picodet.configuration_picodet")
    dummy.PicodetConfig = object
    sys.modules["pdftable.model.picodet.configuration_picodet"] = dummy
    from pdftable.model.picodet.processor_picodet import OCRPicodetPostProcessor

    cfg = types.SimpleNamespace(strides=[8, 16, 32, 64], score_threshold=0.5, nms_threshold=0.5, nms_top_k=1000, keep_top_k=100,
                                id2label={i: f"c{i}" for i in range(num_classes)})
    return OCRPicodetPostProcessor(cfg)


def main():
    out = {}
    for name, idx, c, (oh, ow), nobj in CASES:
        scores, boxes = synth.picodet_planted_outputs(idx, c, n_objects=nobj)
        post = reference_post(c)
        sf = [800.0 / oh, 608.0 / ow]
        res = post({"boxes": scores, "boxes_num": boxes, "org_shape": [oh, ow], "scale_factor": sf, "target_shape": [800, 608]})
        rows = np.array([[r["category_id"], r["score"], *r["bbox"]] for r in res["bboxs"]], np.float64).reshape(-1, 6)
        out[name] = rows
        print(name, rows.shape, res["boxes_num"])
    np.savez_compressed(os.path.join(GOLDEN, "picodet_post.npz"), **out)


def gen_network():
    """LCNet + CSPPAN + PicoHead (picodet/lcnet.py:159, csp_pan.py:233, pico_head.py:972) with the seeded synthetic weights on a
    small pre-processed input, and OCRPicodetPreProcessor (processor_picodet.py:72-113) on a synthetic page."""
    import warnings

    import torch

    warnings.filterwarnings("ignore")
    ref_import.setup()
    dummy = types.ModuleType("pdftable.model.picodet.configuration_picodet")
    dummy.PicodetConfig = object
    sys.modules["pdftable.model.picodet.configuration_picodet"] = dummy
    from pdftable.model.picodet import pico_head
    from pdftable.model.picodet.csp_pan import CSPPAN
    from pdftable.model.picodet.lcnet import LCNet
    from pdftable.model.picodet.processor_picodet import OCRPicodetPreProcessor

    bbsd, nksd, hdsd = synth.picodet_state_dicts(0, 5)
    bb = LCNet(scale=1.0, feature_maps=[3, 4, 5]).eval()
    neck = CSPPAN(in_channels=[128, 256, 512], out_channels=128, kernel_size=5, num_features=4, num_csp_blocks=1, use_depthwise=True,
                  act="hard_swish", spatial_scales=[0.125, 0.0625, 0.03125]).eval()
    head = pico_head.PicoHead(conv_feat=dict(feat_in=128, feat_out=128, num_fpn_stride=4, num_convs=4, norm_type="bn", share_cls_reg=True,
                                             act="hard_swish", use_se=True), num_classes=5, fpn_stride=[8, 16, 32, 64], reg_max=7,
                              feat_in_chan=128, loss_class=dict(), nms=dict(), loss_dfl=None, loss_bbox=None, assigner=None).eval()
    for m, sd in ((bb, bbsd), (neck, nksd), (head, hdsd)):
        r = m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
        assert not r.unexpected_keys and all("scale_reg" in k or "project" in k or "num_batches" in k for k in r.missing_keys), r
    rng = np.random.default_rng(3)
    x = rng.standard_normal((1, 3, 192, 128)).astype(np.float32)
    with torch.no_grad():
        s, b = head(neck(bb(image=torch.from_numpy(x))), export_post_process=False)
    out = {"x": x}
    for lvl in range(4):
        out[f"scores{lvl}"], out[f"dfl{lvl}"] = s[lvl].numpy(), b[lvl].numpy()
    cfg = types.SimpleNamespace(order="hwc", norm_mean=[0.485, 0.456, 0.406], norm_std=[0.229, 0.224, 0.225], scale=1.0 / 255.0,
                                img_height=800, img_width=608)
    page = synth.synthetic_page(9, 500, 380)
    item = OCRPicodetPreProcessor(cfg)(page)
    px = item["image"].numpy()
    out["pre_patch"] = px[:, 300:332, 200:232].copy()
    out["pre_sum"] = np.array([px.astype(np.float64).sum(), np.abs(px.astype(np.float64)).sum()])
    out["pre_meta"] = np.array([*item["org_shape"], *item["scale_factor"], *item["target_shape"]], np.float64)
    np.savez_compressed(os.path.join(GOLDEN, "picodet_net_seed0.npz"), **out)
    print("picodet_net_seed0", [tuple(t.shape) for t in s], out["pre_meta"])


if __name__ == "__main__":
    main()
    gen_network()
