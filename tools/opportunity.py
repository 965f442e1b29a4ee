#!/usr/bin/env python
"""Rank the launches of a per-layer profile (tools/layer_profile.py output, profiles/*_layers_*.txt) by what is left above their
roofline: for each line, algorithmic FLOPs and bytes are recovered from the printed rate x time, the bound is
max(FLOPs / tensor peak, bytes / HBM peak) with the measured peaks of MEASURED_PEAKS.json (fallback: B200_PROFILING.md's), and
`headroom` = time - bound.  Grouped per kernel and listed per layer: the next round's work list, in measured milliseconds.

    python tools/opportunity.py profiles/r1t_layers_dbnet.txt profiles/r2j_layers_rec.txt profiles/r1x_layers_lore.txt
"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINE = re.compile(r"^(\S+)\s+(\S.*?)\s+n=\s*(\d+)\s+([\d.]+) ms\s+([\d.]+)%\s+([\d.]+) TF/s\s+([\d.]+) GB/s")


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p.get("bf16_tflops_sustained", p.get("bf16_tflops", 1400.0))), float(p.get("hbm_gbs", 6650.0))
    except Exception:
        return 1400.0, 6650.0


def main():
    tf_peak, bw_peak = peaks()
    print(f"peaks: {tf_peak:.0f} TFLOP/s (sustained bf16), {bw_peak:.0f} GB/s\n")
    for path in sys.argv[1:]:
        rows = []
        head = open(path).readline().strip()
        for ln in open(path):
            m = LINE.match(ln)
            if not m:
                continue
            kern, layer, n, ms, _, tf, gb = m.group(1), m.group(2).strip(), int(m.group(3)), float(m.group(4)), m.group(5), float(m.group(6)), float(m.group(7))
            t_flop = tf * ms / tf_peak  # (TF/s x ms) / peak TF/s = ms at the tensor peak
            t_byte = gb * ms / bw_peak
            bound = max(t_flop, t_byte)
            rows.append((kern, layer, ms, bound, "tensor" if t_flop >= t_byte else "hbm"))
        total = sum(r[2] for r in rows)
        print(f"## {os.path.basename(path)} — {head}")
        print(f"sum of launches {total:.3f} ms, sum of roofline bounds {sum(r[3] for r in rows):.3f} ms\n")
        by_k = {}
        for kern, _, ms, bound, _ in rows:
            a = by_k.setdefault(kern, [0.0, 0.0, 0])
            a[0] += ms
            a[1] += bound
            a[2] += 1
        print("| kernel | launches | ms | at roofline | headroom ms | share of total headroom |")
        print("|---|---:|---:|---:|---:|---:|")
        head_total = sum(v[0] - v[1] for v in by_k.values())
        for kern, (ms, bound, n) in sorted(by_k.items(), key=lambda kv: -(kv[1][0] - kv[1][1])):
            print(f"| `{kern}` | {n} | {ms:.3f} | {bound:.3f} | {ms - bound:.3f} | {100 * (ms - bound) / head_total:.1f}% |")
        print("\ntop layers by headroom:\n")
        print("| kernel | layer | ms | bound | roofline ms | x over roofline |")
        print("|---|---|---:|---|---:|---:|")
        for kern, layer, ms, bound, which in sorted(rows, key=lambda r: -(r[2] - r[3]))[:12]:
            print(f"| `{kern}` | {layer} | {ms:.3f} | {which} | {bound:.3f} | {ms / max(bound, 1e-6):.1f} |")
        print()


if __name__ == "__main__":
    main()
