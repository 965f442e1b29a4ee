#!/usr/bin/env python
"""BASELINE configs[2]: Lore (DLA-34 + DCNv2, wtw) table structure, batch of 16 synthetic 1024x1024 table crops on one
B200: uint8 crops -> detector -> decode -> sparse cell features -> processor.  Prints one JSON line (images/s, the
per-kernel device times from CUDA events, the conv kernel's achieved TFLOP/s).  Not the driver's bench (bench.py is)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdf_table_b200 import predictors, synth, weights  # noqa: E402
from pdf_table_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--profile", type=int, default=1)
args = ap.parse_args()

sd = synth.lore_dla34_state_dict(0)
sd["hm.2.bias"] = np.array([-0.3, -3.5], np.float32)  # ~100 cells / corners per image with the seeded weights
det = Engine("lore_dla34", weights.pack_lore_dla34(sd))
proc = Engine("lore_processor", weights.pack_lore_processor(synth.lore_processor_state_dict(0)))
post = Engine("post")
pre = [predictors.lore_preprocess(synth.synthetic_page(40 + i, 1024, 1024)) for i in range(4)]
imgs = torch.from_numpy(np.stack([pre[i % 4][0] for i in range(args.batch)])).cuda()
inv = np.stack([predictors.lore_affine([np.float32(pre[i % 4][1][0]), np.float32(pre[i % 4][1][1])], np.float32(pre[i % 4][1][2]), 256, 256, True)
                for i in range(args.batch)])
maps = torch.empty((args.batch, 256, 256, 24), dtype=torch.float32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def step():
    det.lore_detect_forward_u8(imgs, out=maps)
    dec = post.lore_decode(maps, None, None, None, inv)
    feat, offsets = det.lore_cell_features(dec, max_rows=args.batch * 1024)
    return dec, proc.lore_process_forward(feat, offsets)


dec, _ = step()
for _ in range(args.warmup):
    dec, _ = step()
torch.cuda.synchronize()
cells = int(dec["counts"].sum())
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
for a, b in evs:
    flush.fill_(1)
    a.record()
    step()
    b.record()
torch.cuda.synchronize()
ms = sum(a.elapsed_time(b) for a, b in evs) / args.steps
out = {"workload": f"Lore DLA-34+DCNv2 wtw, {args.batch} x 1024x1024 uint8 crops, detect+decode+features+processor",
       "images_per_sec": args.batch / (ms / 1e3), "ms_per_step": ms, "cells_per_step": cells,
       "model_gflop_per_image": det.model_flops / args.batch / 1e9}
if args.profile:
    for e in (det, proc, post):
        e.profile_begin()
    for _ in range(args.steps):
        flush.fill_(1)
        step()
    agg = {}
    for e in (det, proc, post):
        for r in e.profile_report():
            k = agg.setdefault(r["kernel"], {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "n": 0})
            k["ms"] += r["ms"]; k["flops"] += r["flops"]; k["bytes"] += r["bytes"]; k["n"] += 1
    out["kernels"] = {k: {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["n"] / args.steps,
                          "tflops": v["flops"] / (v["ms"] / 1e3) / 1e12 if v["flops"] else None,
                          "gbs": v["bytes"] / (v["ms"] / 1e3) / 1e9} for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}
print(json.dumps(out))
