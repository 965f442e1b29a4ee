#!/usr/bin/env python
"""One Lore (DLA-34) detector forward on N synthetic 1024 x 1024 images -- a target for `ncu -k regex:...` captures."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from pdf_table_b200 import synth, weights  # noqa: E402
from pdf_table_b200.engine import Engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
eng = Engine("lore_dla34", weights.pack_lore_dla34(synth.lore_dla34_state_dict(0)))
img = torch.randint(0, 255, (n, 1024, 1024, 3), dtype=torch.uint8, device="cuda")
for _ in range(reps):
    eng.lore_detect_forward_u8(img)
torch.cuda.synchronize()
