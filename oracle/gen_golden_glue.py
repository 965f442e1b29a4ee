"""Generate tests/golden/glue.npz by running the REFERENCE's OcrCommonUtils.order_point (build container only).

    python -m oracle.gen_golden_glue

300 quads as the orchestrator hands them to order_point between detection and recognition (ocr_pdf/ocr_system_task.py:300-303):
DB-style integer boxes in clockwise order from the top-left corner, rotated and sheared float quads in permuted corner orders,
slivers, and quads with coincident x (the arctan2 ties of an axis-aligned box).  Stored: the inputs and the reference's outputs.
"""
from __future__ import annotations

import math
import os

import numpy as np

from . import ref_import

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def quads():
    rng = np.random.default_rng(20240907)
    out = []
    for k in range(300):
        cx, cy = rng.uniform(50, 900, 2)
        bw, bh = rng.uniform(8, 400), rng.uniform(3, 60)
        ang = 0.0 if k % 3 == 0 else rng.uniform(-1.3, 1.3)
        c, s = math.cos(ang), math.sin(ang)
        p = np.array([[-bw / 2, -bh / 2], [bw / 2, -bh / 2], [bw / 2, bh / 2], [-bw / 2, bh / 2]]) @ np.array([[c, s], [-s, c]]) + [cx, cy]
        if k % 5 == 1:
            p = p + rng.uniform(-3, 3, p.shape)
        if k % 3 == 0:
            p = np.rint(p)  # what DBPostProcess emits: integer corners
        p = np.roll(p, k % 4, axis=0)
        if k % 7 == 3:
            p = p[::-1]
        out.append(p.reshape(8).astype(np.float32 if k % 2 else np.float64))
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

    qs = quads()
    got = np.stack([OcrCommonUtils.order_point(q) for q in qs])
    assert got.dtype == np.float32 and got.shape == (300, 4, 2)
    np.savez_compressed(os.path.join(GOLDEN, "glue.npz"), quads=np.stack([q.astype(np.float64) for q in qs]),
                        is_f32=np.array([q.dtype == np.float32 for q in qs]), ordered=got)
    print("glue.npz:", got.shape)


if __name__ == "__main__":
    main()
