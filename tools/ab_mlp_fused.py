#!/usr/bin/env python
"""A/B of the fused ConvNeXt / ViT MLP kernel against the two-GEMM path on the same weights and crops: run as
   DV_MLP_FUSED=0 python tools/ab_mlp_fused.py save /tmp/ref.npy ;  python tools/ab_mlp_fused.py cmp /tmp/ref.npy"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdf_table_b200 import synth, weights  # noqa: E402
from pdf_table_b200.engine import Engine  # noqa: E402

mode, path = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 8
sd = synth.convnext_vit_state_dict(0)
eng = Engine("convnext_vit", weights.pack_convnext_vit(sd))
rng = np.random.default_rng(3)
chunks = torch.from_numpy(rng.random((n * 3, 3, 32, 300), dtype=np.float32)).cuda()
logits = eng.convnextvit_forward(chunks, return_logits=True)[1].float().cpu().numpy()
if mode == "save":
    np.save(path, logits)
    print("saved", logits.shape, float(np.abs(logits).max()))
else:
    ref = np.load(path)
    d = np.abs(ref - logits)
    print("max |fused - unfused| =", float(d.max()), "mean", float(d.mean()), "logit sigma", float(ref.std()),
          "argmax equal", float((ref.argmax(-1) == logits.argmax(-1)).mean()))
