#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (both arms), ncu launch list + one full capture of the conv kernel.
# usage: gpurun --timeout 1500 -- bash tools/gpu_check.sh [tag]      (env: FULL=1 adds the full-cascade lines, SKIP_NCU=1 drops the ncu passes)
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/${TAG}_bench_ref.json
if [ -n "$FULL" ]; then  # FULL=1: BASELINE configs[4] per GPU, table crops uploaded (planted) vs cut + warped on the device
timeout 600 python bench.py --cascade full --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_full_cascade.json 2> gpurun_out/${TAG}_bench_full.err; echo "full rc=$?"; tail -c 400 gpurun_out/${TAG}_bench_full.err
timeout 600 python bench.py --cascade full --tables device --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_full_cascade_devtables.json 2> gpurun_out/${TAG}_bench_full_dev.err; echo "full(device tables) rc=$?"; tail -c 400 gpurun_out/${TAG}_bench_full_dev.err
timeout 600 python tools/bench_rec.py --steps 5 > gpurun_out/${TAG}_rec_bench.json 2> gpurun_out/${TAG}_rec_bench.err; echo "rec bench (configs[3]) rc=$?"; tail -c 300 gpurun_out/${TAG}_rec_bench.err; head -c 700 gpurun_out/${TAG}_rec_bench.json; echo
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/*_bench_full_cascade*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), "pages/s device,", round(d["e2e"]["value"], 1), "e2e, h2d", d["e2e"]["h2d_bytes_per_step"])
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
fi
if [ -z "$SKIP_NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
# full captures: summarised ON THE BOX (raw metric CSV + headline / wait-site text), the .ncu-rep files stay there -- gpurun
# merges at most 64 MiB back and 30 full-set kernels with sources are more than that
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 54 -c 27 -o /tmp/${TAG}_conv -f python tools/probe_dbnet.py 32 960 1 > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/${TAG}_conv.ncu-rep --page raw --csv > gpurun_out/${TAG}_conv_full.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_fused -s 0 -c 7 -o /tmp/${TAG}_mlp -f python tools/layer_profile.py rec > gpurun_out/${TAG}_ncu_mlp.log 2>&1; echo "ncu mlp rc=$?"
ncu -i /tmp/${TAG}_mlp.ncu-rep --page raw --csv > gpurun_out/${TAG}_mlp_full.csv 2>/dev/null
{ python tools/ncu_keys.py /tmp/${TAG}_mlp.ncu-rep 4; python tools/ncu_sync_sites.py /tmp/${TAG}_mlp.ncu-rep 4 60; } > gpurun_out/${TAG}_mlp_ncu.txt 2>&1
fi
du -sm gpurun_out | tail -1
