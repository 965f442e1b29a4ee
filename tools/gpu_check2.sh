#!/bin/bash
# One gpurun call (round 2): GPU parity tests, smoke, the default bench line (full cascade + blocks), the reference arm.
# usage: gpurun --timeout 1500 -- bash tools/gpu_check2.sh TAG [pytest args]
TAG=${1:-r3}
shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
timeout 1200 python -m pytest tests -m gpu -q "$@" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -40 gpurun_out/${TAG}_pytest.log
fi
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("MAIN", round(d["value"], 1), d["unit"], "e2e", round(d["e2e"]["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e ms", round(d["e2e"]["ms_per_step"], 2),
          "h2d", d["e2e"]["h2d_bytes_per_step"], "d2h", d["e2e"]["d2h_bytes_per_step"], "boxes", d["e2e"].get("boxes_per_step"), "cells", d["e2e"].get("cells_per_step"))
    print("ROOF", {k: d["roofline"][k] for k in ("kernel", "achieved", "frac", "share_of_step")}, "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"], 3))
    for n, b in d["blocks"].items():
        print("BLOCK", n, round(b["value"], 1), b["unit"], "e2e", round(b["e2e"]["value"], 1), "roof", b["roofline"]["kernel"], round(b["roofline"]["frac"], 3), "cpu", b.get("cpu_baseline", {}).get("value"))
    for k, v in list(d["kernels"].items())[:14]:
        print("  %-28s %8.3f ms %5.0f x  tflops %s gbs %.0f" % (k, v["ms_per_step"], v["launches_per_step"], v["tflops"] and round(v["tflops"]), v["gbs"]))
except Exception as ex:
    print("bench line unreadable:", ex)
PY
if [ -n "$REF" ]; then
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"; tail -c 700 gpurun_out/${TAG}_bench_ref.json
fi
du -sm gpurun_out | tail -1
