"""Generate tests/golden/crop.npz by running the REFERENCE's OcrCommonUtils.crop_image (build container only).

    python -m oracle.gen_golden_crop

A synthetic 480x640 page and fourteen quads: axis-aligned, rotated up to 35 degrees, sheared (non-rectangular),
corner orders permuted, one hanging over the page border (border value 0), one 3-pixel-high sliver and one wider than
64 pixels on a short side (the remap's column blocks).  utils/ocr/ocr_common_utils.py imports fitz-free helpers only
through `pdftable.utils` (stubbed by ref_import) -- the module file itself is loaded unmodified.
"""
from __future__ import annotations

import math
import os

import numpy as np

from . import ref_import
from pdf_table_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def page():
    return synth.synthetic_page(77, 480, 640)


def quads():
    rng = np.random.default_rng(20240906)
    out = []
    specs = [(320, 100, 200, 30, 0.0), (100, 300, 150, 24, 0.1), (400, 350, 260, 40, -0.3), (200, 200, 90, 18, 0.6), (500, 80, 120, 60, -0.6),
             (60, 60, 100, 30, 0.2), (600, 440, 120, 40, 0.15), (320, 240, 300, 3.4, 0.05), (320, 400, 40, 120, 0.0), (150, 120, 64, 16, 0.0),
             (300, 300, 65, 15, 0.02), (450, 200, 33.3, 12.7, 1.2), (250, 60, 180, 28, -0.12), (320, 240, 500, 300, 0.01)]
    for k, (cx, cy, bw, bh, ang) in enumerate(specs):
        c, s = math.cos(ang), math.sin(ang)
        p = np.array([[-bw / 2, -bh / 2], [bw / 2, -bh / 2], [bw / 2, bh / 2], [-bw / 2, bh / 2]]) @ np.array([[c, s], [-s, c]]) + [cx, cy]
        if k % 3 == 1:
            p = p + rng.uniform(-2.5, 2.5, p.shape)  # not a rectangle
        p = np.roll(p, k, axis=0) if k % 2 else p[::-1]  # the corner order the detector emits is not canonical
        out.append(p.astype(np.float32) if k % 4 else np.rint(p).astype(np.int32).astype(np.float32))
    return out


def main():
    ref_import.setup()
    import sys
    import types

    u = sys.modules["pdftable.utils"]
    u.BaseUtil = type("BaseUtil", (), {})
    ocr_pkg = types.ModuleType("pdftable.utils.ocr")
    ocr_pkg.__path__ = [os.path.join(ref_import.R, "utils", "ocr")]
    sys.modules["pdftable.utils.ocr"] = ocr_pkg
    from pdftable.utils.ocr.ocr_common_utils import OcrCommonUtils

    img = page()
    out = {"n": np.int64(len(quads()))}
    for k, q in enumerate(quads()):
        crop = OcrCommonUtils.crop_image(img, q)
        out[f"quad{k}"] = q
        out[f"crop{k}"] = crop
        print(k, crop.shape)
    np.savez_compressed(os.path.join(GOLDEN, "crop.npz"), **out)


if __name__ == "__main__":
    main()
