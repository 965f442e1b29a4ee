"""Quick perf probe: DBNet-R18 forward on synthetic pages (not the bench; used while tuning)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from pdf_table_b200 import synth
from pdf_table_b200 import weights
from pdf_table_b200.engine import Engine

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
H = W = int(sys.argv[2]) if len(sys.argv) > 2 else 960
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
sd = synth.dbnet_r18_state_dict(0)
eng = Engine("dbnet_r18", weights.pack_dbnet_r18(sd))
x = torch.randn(N, 3, H, W, device="cuda")
out = torch.empty(N, 1, H, W, device="cuda")
for _ in range(2):
    eng.dbnet_forward(x, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    eng.dbnet_forward(x, out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
fl = eng.model_flops
print(f"N={N} {H}x{W}: {ms:.3f} ms/step, {N / ms * 1e3:.1f} pages/s, {fl / ms / 1e9:.1f} TFLOP/s (model flops {fl / 1e9:.1f} G)")
