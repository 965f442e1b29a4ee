"""Quick perf probe: ConvNextViT forward on random chunks (not the bench; used while tuning)."""
import sys, json, collections
import numpy as np, torch
sys.path.insert(0, ".")
from pdf_table_b200 import synth, weights
from pdf_table_b200.engine import Engine

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pc = int(sys.argv[3]) if len(sys.argv) > 3 else 96
eng = Engine("convnext_vit", weights.pack_convnext_vit(synth.convnext_vit_state_dict(0)))
eng.set_pass_crops(pc)
x = torch.rand(3 * N, 3, 32, 300, device="cuda")
for _ in range(2):
    eng.convnextvit_forward(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    eng.convnextvit_forward(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
fl = eng.model_flops
print(f"N={N} crops pass={pc}: {ms:.3f} ms/step, {N / ms * 1e3:.1f} crops/s, {fl / ms / 1e9:.1f} TFLOP/s (model flops {fl / 1e9 / N:.2f} G/crop)")
if len(sys.argv) > 4:
    eng.profile_begin()
    eng.convnextvit_forward(x)
    recs = eng.profile_report()
    agg = collections.OrderedDict()
    for r in recs:
        key = (r["kernel"], r["layer"].split(".")[-1] if r["kernel"].startswith("conv") else "")
        a = agg.setdefault(key, [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += r["ms"]; a[2] += r["flops"]; a[3] += r["bytes"]
    tot = sum(a[1] for a in agg.values())
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[0]:22s} {k[1]:8s} n={a[0]:4d} {a[1]:9.3f} ms {a[1]/tot*100:5.1f}%  {a[2]/a[1]/1e9 if a[1] else 0:8.1f} TF/s {a[3]/a[1]/1e6 if a[1] else 0:8.1f} GB/s")
