#!/usr/bin/env python
"""PP-OCRv4 recogniser alone: device-resident crops/s for a few pass sizes (DV_REC_PASS is read at engine creation)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdf_table_b200 import pp_rec_graph, synth  # noqa: E402
from pdf_table_b200.engine import Engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
blob = pp_rec_graph.pack_pp_rec(synth.pp_ocrv4_rec_state_dict(0, 97))
crops = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (n, 48, 320, 3), dtype=np.uint8)).cuda()
widths = torch.full((n,), 320, dtype=torch.int32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {}
for ps in (int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "4096,1024,512,256,128,64").split(",")):
    os.environ["DV_REC_PASS"] = str(ps)
    eng, post = Engine("pp_rec", blob), Engine("post")
    for _ in range(3):
        ids, maxp = eng.rec_forward_u8(crops, widths)
    torch.cuda.synchronize()
    evs = []
    for _ in range(5):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ids, maxp = eng.rec_forward_u8(crops, widths)
        post.ctc_collapse(ids, maxp)
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs) / 5
    eng.profile_begin()
    eng.rec_forward_u8(crops, widths)
    agg = {}
    recs = eng.profile_report()
    if os.environ.get("LAYERS"):
        for r in recs:
            print("   %-22s %-8s %8.3f ms  %7.1f TF/s %7.0f GB/s" % (r["kernel"], r["layer"], r["ms"], r["flops"] / r["ms"] / 1e9, r["bytes"] / r["ms"] / 1e6))
    for r in recs:
        k = agg.setdefault(r["kernel"], [0.0, 0])
        k[0] += r["ms"]
        k[1] += 1
    out[ps] = {"ms": ms, "crops_per_sec": n / ms * 1e3, "kernels": {k: [round(v[0], 3), v[1]] for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])}}
    print(ps, round(ms, 2), "ms", round(n / ms * 1e3), "crops/s", out[ps]["kernels"], flush=True)
    eng.close()
    post.close()
json.dump(out, open("gpurun_out/pp_rec_pass_sweep.json", "w"))
