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


if __name__ == "__main__":
    main()
