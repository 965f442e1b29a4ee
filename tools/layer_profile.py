"""Per-layer device times (CUDA events around every launch, dv_profile_*) of one forward of a model.
usage: python tools/layer_profile.py {dbnet|rec|ppdet|lore} [batch]   -- tuning aid, not the bench."""
import collections
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from pdf_table_b200 import synth, weights  # noqa: E402
from pdf_table_b200.engine import Engine  # noqa: E402

model = sys.argv[1]
if model == "dbnet":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    eng = Engine("dbnet_r18", weights.pack_dbnet_r18(synth.dbnet_r18_state_dict(0)))
    x = torch.randn(n, 3, 960, 960, device="cuda")
    out = torch.empty(n, 1, 960, 960, device="cuda")
    run = lambda: eng.dbnet_forward(x, out)
elif model == "rec":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 96
    eng = Engine("convnext_vit", weights.pack_convnext_vit(synth.convnext_vit_state_dict(0)))
    x = torch.rand(3 * n, 3, 32, 300, device="cuda")
    run = lambda: eng.convnextvit_forward(x)
elif model == "ppdet":
    from pdf_table_b200 import pp_det_graph

    n = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    eng = Engine("pp_det", pp_det_graph.pack_pp_det(synth.pp_ocrv4_det_state_dict(0)))
    x = torch.randint(0, 255, (n, 960, 960, 3), dtype=torch.uint8, device="cuda")
    out = torch.empty(n, 1, 960, 960, device="cuda")
    run = lambda: eng.dbnet_forward_u8(x, (0.485, 0.456, 0.406), (0.229, 0.224, 0.225), 1.0 / 255.0, True, out=out)
else:
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    eng = Engine("lore_dla34", weights.pack_lore_dla34(synth.lore_dla34_state_dict(0)))
    x = torch.randint(0, 255, (n, 1024, 1024, 3), dtype=torch.uint8, device="cuda")
    run = lambda: eng.lore_detect_forward_u8(x)
for _ in range(3):
    run()
torch.cuda.synchronize()
reps = 3
eng.profile_begin()
for _ in range(reps):
    run()
recs = eng.profile_report()
agg = collections.OrderedDict()
for r in recs:
    a = agg.setdefault((r["kernel"], r["layer"]), [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += r["ms"]
    a[2] += r["flops"]
    a[3] += r["bytes"]
tot = sum(a[1] for a in agg.values())
print(f"{model} batch {n}: {tot / reps:.3f} ms per forward (sum of kernels), {sum(a[2] for a in agg.values()) / tot / 1e9:.1f} TFLOP/s overall")
for k, a in agg.items():
    print(f"{k[0]:22s} {k[1]:28s} n={a[0] // reps:3d} {a[1] / reps:8.3f} ms {a[1] / tot * 100:5.1f}%  {a[2] / a[1] / 1e9 if a[1] else 0:7.1f} TF/s {a[3] / a[1] / 1e6 if a[1] else 0:8.1f} GB/s")
