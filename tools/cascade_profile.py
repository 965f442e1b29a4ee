#!/usr/bin/env python
"""Per-(engine, kernel, layer) device times of ONE step of bench.py's cascade (CUDA events around every launch, dv_profile_*),
each with its roofline bound max(FLOPs / tensor peak, bytes / HBM peak) from MEASURED_PEAKS.json and the time above it --
the work list of the headline number in measured milliseconds.  Tuning aid, not the bench.

    python tools/cascade_profile.py [--det ppocrv4|dbnet_r18] [--top 60]
"""
import argparse
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--det", default="ppocrv4")
    ap.add_argument("--top", type=int, default=60)
    args = ap.parse_args()
    bench.DET = args.det
    bench.TWO_STREAMS = False  # per-launch events are only meaningful when the branches do not overlap
    peaks, _ = bench.load_peaks() if hasattr(bench, "load_peaks") else ({"hbm_gbs": 6536.0, "bf16_tflops_sustained": 1353.7}, "")
    tf_peak = float(peaks.get("bf16_tflops_sustained", 1353.7)) * 1e12
    bw_peak = float(peaks.get("hbm_gbs", 6536.0)) * 1e9
    wl = bench.Cascade(0, 0, True)
    for _ in range(3):
        wl.step_device()
    torch.cuda.synchronize()
    names = {id(e): n for n, e in (("det", wl.det), ("rec", wl.rec), ("post", wl.post), ("layout", wl.layout), ("lore", wl.lore), ("lore_proc", wl.lore_proc))}
    for e in wl.engines:
        e.profile_begin()
    reps = 3
    for _ in range(reps):
        wl.flush_l2()
        wl.step_device()
    torch.cuda.synchronize()
    agg = collections.OrderedDict()
    for e in wl.engines:
        for r in e.profile_report():
            a = agg.setdefault((names.get(id(e), "post*"), r["kernel"], r["layer"]), [0, 0.0, 0.0, 0.0])
            a[0] += 1
            a[1] += r["ms"]
            a[2] += r["flops"]
            a[3] += r["bytes"]
    rows = []
    for (eng, k, layer), (n, ms, fl, by) in agg.items():
        ms, fl, by, n = ms / reps, fl / reps, by / reps, n // reps
        bound = max(fl / tf_peak, by / bw_peak) * 1e3
        rows.append((ms - bound, eng, k, layer, n, ms, bound, fl / ms / 1e9 if ms else 0, by / ms / 1e6 if ms else 0))
    tot = sum(r[5] for r in rows)
    print(f"cascade step ({args.det}): {tot:.2f} ms summed over kernels; peaks {tf_peak / 1e12:.0f} TFLOP/s, {bw_peak / 1e9:.0f} GB/s")
    per_eng = collections.Counter()
    for r in rows:
        per_eng[r[1]] += r[5]
    print("per engine: " + ", ".join(f"{k} {v:.2f}" for k, v in per_eng.most_common()))
    print(f"{'engine':9s} {'kernel':22s} {'layer':30s} {'n':>3s} {'ms':>8s} {'bound':>8s} {'above':>8s} {'TF/s':>7s} {'GB/s':>7s}")
    for above, eng, k, layer, n, ms, bound, tf, gb in sorted(rows, reverse=True)[: args.top]:
        print(f"{eng:9s} {k:22s} {layer[:30]:30s} {n:3d} {ms:8.3f} {bound:8.3f} {above:8.3f} {tf:7.1f} {gb:7.0f}")


if __name__ == "__main__":
    main()
