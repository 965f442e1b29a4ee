#!/bin/bash
# ncu evidence for the non-GEMM stages north_star lists (DB threshold + box seed, CTC greedy decode, CenterNet / Lore peak NMS +
# top-K + gathers, PicoDet anchor decode, crop glue): dram bytes + duration per launch over one cascade step at BASELINE sizes
# (and the CenterNet test for k_cn_*).  usage: gpurun -- bash tools/ncu_nongemm.sh TAG      -> gpurun_out/TAG_nongemm_ncu.{csv,md}
TAG=${1:-r4}
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct
K='regex:k_db_|k_ctc|k_collapse|k_lore_|k_pico_|k_gather_patch|k_logi|k_cell_off|k_crop_|k_quad_|k_warp_affine|k_resize_linear|k_softmax_rows|k_sigmoid_cols|k_pp_rec_norm'
timeout 900 ncu --metrics $M --clock-control none -k "$K" -s 120 -c 200 --csv --log-file gpurun_out/${TAG}_nongemm_ncu.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-blocks > /dev/null 2>&1
timeout 600 ncu --metrics $M --clock-control none -k 'regex:k_cn_|k_ctc' --csv --log-file gpurun_out/${TAG}_nongemm_ncu_cn.csv python -m pytest tests/test_gpu_centernet.py tests/test_gpu_ctc.py -q -x > /dev/null 2>&1
python - "$TAG" <<'PY'
import csv, sys, collections, json, os
tag = sys.argv[1]
peak = 6536.0
try:
    peak = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"])
except Exception:
    pass
out = [f"# non-GEMM stages under ncu ({tag}): per-launch duration, DRAM bytes (read + write), achieved DRAM GB/s vs {peak:.0f} GB/s measured peak", "",
       "Cascade step at BASELINE sizes (32 pages 960x960, 1280 crops, 32 tables) and the CenterNet / CTC tests; `--clock-control none`.", "",
       "| kernel | launches | avg us | DRAM MB / launch | GB/s | frac of peak | warps active % | L1 hit % | L2 hit % |", "|---|---|---|---|---|---|---|---|---|"]
for f in (f"gpurun_out/{tag}_nongemm_ncu.csv", f"gpurun_out/{tag}_nongemm_ncu_cn.csv"):
    if not os.path.exists(f):
        continue
    rows = [r for r in csv.reader(l for l in open(f) if l.startswith('"'))]
    if not rows:
        continue
    hdr = rows[0]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    ii = hdr.index("ID")
    per = collections.OrderedDict()
    for r in rows[1:]:
        d = per.setdefault((r[ii], r[ki].split("(")[0]), {})
        d[r[mi]] = float(r[vi].replace(",", ""))
    agg = collections.OrderedDict()
    for (_, k), d in per.items():
        a = agg.setdefault(k, [0, 0.0, 0.0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        a[3] += d.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0.0)
        a[4] += d.get("l1tex__t_sector_hit_rate.pct", 0.0)
        a[5] += d.get("lts__t_sector_hit_rate.pct", 0.0)
    units = {r[mi]: r[hdr.index("Metric Unit")] for r in rows[1:]}
    tu = {"us": 1.0, "ms": 1e3, "ns": 1e-3}.get(units.get("gpu__time_duration.sum", "us"), 1.0)
    bu = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units.get("dram__bytes_read.sum", "byte"), 1.0)
    for k, a in agg.items():
        us, by = a[1] * tu / a[0], a[2] * bu / a[0]
        gbs = by / (us * 1e-6) / 1e9 if us else 0
        out.append(f"| `{k}` | {a[0]} | {us:.1f} | {by / 1e6:.2f} | {gbs:.0f} | {gbs / peak:.3f} | {a[3] / a[0]:.0f} | {a[4] / a[0]:.0f} | {a[5] / a[0]:.0f} |")
open(f"gpurun_out/{tag}_nongemm_ncu.md", "w").write("\n".join(out) + "\n")
print("\n".join(out))
PY
