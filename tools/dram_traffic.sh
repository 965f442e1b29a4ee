#!/bin/bash
# roofline.traffic source: dram__bytes_read.sum + dram__bytes_write.sum per launch of every kernel, from an ncu pass over the
# default bench workload (one-stream step) and over the three blocks.  usage: gpurun -- bash tools/dram_traffic.sh TAG
#   -> gpurun_out/TAG_dram_traffic_{full,pp_rec,rec_sweep,lore}.json (copy to profiles/ as TAG_{workload}_dram_traffic.json: bench.py globs *_dram_traffic.json)
TAG=${1:-r4}
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
run() {  # name, command...
  local name=$1; shift
  timeout 900 ncu --metrics $M --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_traffic_${name}.csv "$@" > /dev/null 2>&1
  python - "$TAG" "$name" "$*" <<'PY'
import csv, sys, json, collections, re
tag, name, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(l for l in open(f"gpurun_out/{tag}_traffic_{name}.csv") if l.startswith('"'))]
hdr = rows[0]
ki, mi, vi, ui, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    if r[mi].startswith("dram"):
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    else:
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
    base = re.sub(r"^void ", "", r[ki]).replace("<unnamed>::", "").replace("(anonymous namespace)::", "").replace("unnamed>::", "")
    base = re.sub(r"[<(].*$", "", base).split("::")[-1].strip()
    base = {"k_dwconv_row": "k_dwconv", "k_dwconv_c2": "k_dwconv", "k_up_dw_add_cls": "k_up_dw_add"}.get(base, base)  # bench.py's launch names
    per.setdefault((r[ii], base), {})[r[mi]] = v
agg = collections.OrderedDict()
for (_, k), d in per.items():
    a = agg.setdefault(k, [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get("dram__bytes_read.sum", 0.0)
    a[2] += d.get("dram__bytes_write.sum", 0.0)
    a[3] += d.get("gpu__time_duration.sum", 0.0)
out = {"source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none over every launch of `{cmd}` "
                 f"(tools/dram_traffic.sh, build {tag}); bytes per launch, averaged over the launches of each kernel",
       "workload": name,
       "kernels": {k: {"launches": a[0], "dram_read_bytes_per_launch": a[1] / a[0], "dram_write_bytes_per_launch": a[2] / a[0], "avg_us": a[3] / a[0]}
                   for k, a in agg.items()}}
json.dump(out, open(f"gpurun_out/{tag}_dram_traffic_{name}.json", "w"), indent=1)
top = sorted(out["kernels"].items(), key=lambda kv: -kv[1]["avg_us"] * kv[1]["launches"])[:6]
print(name, [(k, v["launches"], round((v["dram_read_bytes_per_launch"] + v["dram_write_bytes_per_launch"]) / 1e6, 1), round(v["avg_us"], 1)) for k, v in top])
PY
}
ONLY=${2:-all}
[ "$ONLY" = all -o "$ONLY" = full ] && DV_BENCH_STREAMS=1 run full python bench.py --steps 1 --warmup 3 --no-blocks --no-cpu-baseline
[ "$ONLY" = all -o "$ONLY" = pp_rec ] && run pp_rec python tools/bench_pp_rec.py
[ "$ONLY" = all -o "$ONLY" = lore ] && run lore python tools/bench_lore.py --steps 1
[ "$ONLY" = all -o "$ONLY" = rec_sweep ] && run rec_sweep python tools/bench_rec.py --steps 1
true
