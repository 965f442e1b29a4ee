#!/usr/bin/env python
"""One conv_igemm_tcgen05 shape through dv_conv2d_nhwc_f16, timed with CUDA events (and meant to be wrapped in ncu):

    python tools/probe_conv.py N H W Cin Cout [k] [reps]

Tuning aid for the halo / patch modes (DV_HALO=0|1|2), not part of the bench."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from pdf_table_b200 import weights  # noqa: E402
from pdf_table_b200.engine import Engine  # noqa: E402


def main():
    n, h, w, cin, cout = (int(v) for v in sys.argv[1:6])
    k = int(sys.argv[6]) if len(sys.argv) > 6 else 3
    reps = int(sys.argv[7]) if len(sys.argv) > 7 else 10
    rng = np.random.default_rng(0)
    wt = (rng.standard_normal((cout, cin, k, k)) / np.sqrt(cin * k * k)).astype(np.float32)
    wp, bp = weights.pack_conv(wt, np.zeros(cout, np.float32))
    eng = Engine("post")
    x = torch.randn((n, h, w, cin), device="cuda").half()
    wd, bd = torch.from_numpy(wp).cuda(), torch.from_numpy(bp).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        eng.conv2d_nhwc(x, wd, bd, cout, k, 1, k // 2, act=1)
    ms = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.conv2d_nhwc(x, wd, bd, cout, k, 1, k // 2, act=1)
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    t = float(np.median(ms))
    flops = 2.0 * n * h * w * cin * k * k * cout
    byts = 2.0 * n * h * w * (cin + cout)
    print(f"conv {n}x{h}x{w} {cin}->{cout} k{k} DV_HALO={os.environ.get('DV_HALO', '')}: {t:.3f} ms  {flops / t / 1e9:.1f} TFLOP/s  {byts / t / 1e6:.0f} GB/s "
          f"(bounds: {flops / 1353.7e9:.3f} ms tensor, {byts / 6536e6:.3f} ms hbm)")


if __name__ == "__main__":
    main()
