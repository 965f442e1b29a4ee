"""Generate the CenterNet golden fixtures by running the REFERENCE's own modules (build container only).

    python -m oracle.gen_golden_centernet

  centernet_dla34_seed0.npz : DLASeg() (center_net/modeling_centernet.py:601) with the seeded synthetic state_dict.
  centernet_decode.npz      : OCRTableCenterNetPostProcessor (center_net/processer_centernet.py:170) on planted head maps
                              (sigmoid_() neutralised: the engine's contract is the post-sigmoid map).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

from . import ref_import
from pdf_table_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
DECODE_CASES = [("c0", 20, 128, 128, (600, 800)), ("c1", 21, 128, 160, (1024, 1280)), ("c2", 22, 256, 256, (1500, 1100))]


def main():
    ref_import.setup()
    from pdftable.model.center_net.modeling_centernet import DLASeg

    m = DLASeg().eval()
    sd = synth.centernet_dla34_state_dict(0)
    r = m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    assert not r.unexpected_keys and all(k.startswith("base.fc") or "num_batches" in k for k in r.missing_keys), r
    rng = np.random.default_rng(6)
    x = rng.standard_normal((1, 3, 64, 96)).astype(np.float32)
    with torch.no_grad():
        out = m(torch.from_numpy(x))[0]
    np.savez_compressed(os.path.join(GOLDEN, "centernet_dla34_seed0.npz"), x=x, **{k: v.numpy() for k, v in out.items()})
    print("centernet_dla34_seed0", {k: tuple(v.shape) for k, v in out.items()})

    stub = types.ModuleType("pdftable.utils.ocr")
    stub.OcrCommonUtils = type("OcrCommonUtils", (), {})
    sys.modules.setdefault("pdftable.utils.ocr", stub)
    from pdftable.model.center_net.processer_centernet import OCRTableCenterNetPostProcessor

    post = OCRTableCenterNetPostProcessor()
    res = {}
    for name, idx, h, w, (src_h, src_w) in DECODE_CASES:
        maps = synth.lore_planted_maps(idx, h, w, with_feat=False)
        t = {"hm": torch.from_numpy(maps["hm"])[None].clone(), "reg": torch.from_numpy(maps["reg"])[None], "c2v": torch.from_numpy(maps["wh"])[None],
             "v2c": torch.from_numpy(maps["st"])[None]}
        meta = {"c": np.array([src_w / 2.0, src_h / 2.0], dtype=np.float32), "s": max(src_h, src_w) * 1.0, "out_height": h, "out_width": w}
        orig = torch.Tensor.sigmoid_
        torch.Tensor.sigmoid_ = lambda self: self
        try:
            out = post({"results": [t], "meta": meta})
        finally:
            torch.Tensor.sigmoid_ = orig
        res[name] = np.asarray(out["polygons"], np.float32).reshape(-1, 8)
        print(name, res[name].shape)
    np.savez_compressed(os.path.join(GOLDEN, "centernet_decode.npz"), **res)


if __name__ == "__main__":
    main()
