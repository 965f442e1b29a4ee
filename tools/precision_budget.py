#!/usr/bin/env python
"""Where the recogniser's distance from the fp32 oracle comes from, measured on CPU (no GPU needed): the oracle forward of
ConvNextViT (oracle/convnextvit_ref.py) with the operands of its GEMM-type layers rounded to fp16 -- the weights only, the
activations only, both -- against the unrounded forward.  The engine (fp16 operands, fp32 accumulation / residual stream / LN /
softmax) measures 8.5e-3 max |dlogit| on the GPU; "both" below is its CPU model.  Test infrastructure / analysis only.

    python tools/precision_budget.py [n_crops]
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import convnextvit_ref as ref  # noqa: E402
from pdf_table_b200 import synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    sd = synth.convnext_vit_state_dict(0)
    chunks = ref.preprocess([synth.synthetic_text_crop(700 + i, 32, 320) for i in range(n)])
    want = ref.convnextvit_forward(sd, chunks)
    lin0, conv0 = F.linear, F.conv2d
    r16 = lambda x: x.half().float()

    def run(act16: bool, w16: bool, split: bool = False):
        def lin(x, w, b=None):
            if split:  # split-fp16 (the Lore processor's mode): A_hi W_hi + A_hi W_lo + A_lo W_hi, fp32 accumulation
                xh, wh = r16(x), r16(w)
                xl, wl = r16(x - xh), r16(w - wh)
                return lin0(xh, wh, b) + lin0(xh, wl) + lin0(xl, wh)
            return lin0(r16(x) if act16 else x, r16(w) if w16 else w, b)

        def conv(x, w, b=None, *a, **k):
            if k.get("groups", 1) == 1 and w.shape[1] > 1:  # the GEMM-type convs; depthwise / patchify run in fp32 on CUDA cores
                if split:
                    xh, wh = r16(x), r16(w)
                    xl, wl = r16(x - xh), r16(w - wh)
                    return conv0(xh, wh, b, *a, **k) + conv0(xh, wl, None, *a, **k) + conv0(xl, wh, None, *a, **k)
                return conv0(r16(x) if act16 else x, r16(w) if w16 else w, b, *a, **k)
            return conv0(x, w, b, *a, **k)

        F.linear, F.conv2d = lin, conv
        try:
            return ref.convnextvit_forward(sd, chunks)
        finally:
            F.linear, F.conv2d = lin0, conv0

    print(f"ConvNextViT, {n} crops ({chunks.shape[0]} chunks), logit sigma {float(want.std()):.2f}")
    modes = {"fp16 weights only": (False, True, False), "fp16 activations only": (True, False, False),
             "both (the engine's operand precision)": (True, True, False), "split-fp16 on both operands (3 MMAs)": (True, True, True)}
    for name, (a, w, sp) in modes.items():
        got = run(a, w, sp)
        d = got - want
        print(f"  {name:40s} max|dlogit| = {float(d.abs().max()):.2e}   rms = {float(d.pow(2).mean().sqrt()):.2e}   "
              f"arg-max equal = {float((got.argmax(-1) == want.argmax(-1)).float().mean()):.5f}")


if __name__ == "__main__":
    main()
