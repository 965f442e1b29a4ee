"""Generate tests/golden/pp_rec_pre.npz by running the REFERENCE's PPOcrRecPreProcessor (build container only).

    python -m oracle.gen_golden_pp_rec_pre

Thirteen seeded synthetic text-line crops (heights 12 .. 70, aspect ratios 0.27 .. 30: below the minimum width, above the
maximum width, equal ratios, one grey-scale crop) -> the reference's batches (image tensors, sort indices, batch starts).
The module imports configuration_ocr_recognition_pp, whose config class trips transformers>=5's dataclass check
(SURVEY.md section 10); a plain namespace with the same four attributes is handed to the processor instead.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

from . import ref_import

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SHAPES = [(32, 320), (48, 100), (20, 15), (30, 8), (24, 200), (17, 150), (48, 48), (70, 2100), (12, 96), (33, 66), (33, 66), (40, 401),
          (25, 77)]


def crops():
    rng = np.random.default_rng(20240905)
    out = []
    for k, (h, w) in enumerate(SHAPES):
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        if k == 7:  # the one crop wider than limited_max_width: a smooth ramp with steps, so the 1280-wide golden batch compresses
            ramp = (np.arange(w)[None, :, None] // 7 * 3 + np.arange(h)[:, None, None] * 2 + np.arange(3)[None, None, :] * 40) % 256
            img = ramp.astype(np.uint8)
        out.append(img[:, :, 0].copy() if k == 6 else img)  # one 2-D grey crop: the reference converts it with GRAY2RGB
    return out


def main():
    ref_import.setup()
    dummy = types.ModuleType("pdftable.model.ocr_rec_pp.configuration_ocr_recognition_pp")
    dummy.PPOcrRecognitionConfig = object
    sys.modules["pdftable.model.ocr_rec_pp.configuration_ocr_recognition_pp"] = dummy
    from pdftable.model.ocr_rec_pp.processor_ocr_rec_pp import PPOcrRecPreProcessor

    cfg = types.SimpleNamespace(rec_image_shape=[3, 48, 320], rec_batch_num=6, limited_max_width=1280, limited_min_width=16)
    batches = PPOcrRecPreProcessor(cfg)(crops())
    out = {"n_batches": np.int64(len(batches)), "indices": batches[0]["indices"]}
    for k, b in enumerate(batches):
        out[f"image{k}"] = b["image"]
        out[f"beg{k}"] = np.int64(b["batch_beg_img_no"])
        print(k, b["image"].shape, b["batch_beg_img_no"])
    np.savez_compressed(os.path.join(GOLDEN, "pp_rec_pre.npz"), **out)


if __name__ == "__main__":
    main()
