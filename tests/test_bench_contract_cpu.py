"""bench.py's driver contract, the part that runs without a GPU: the reference arm prints ONE JSON line with the agreed keys, and
the B200 arm refuses to run (no CPU fallback) instead of printing a number."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600,
                          env={**os.environ, **(env or {})})


def test_reference_arm_json_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "pages_per_sec" and d["unit"] == "pages/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["workload"].startswith("BASELINE configs[4]") and "model" not in d["config"]  # the default line = the full cascade
    assert d["blocks"]["rec_sweep"]["unit"] == "crops/s" and d["blocks"]["rec_sweep"]["value"] > 0 and d["blocks"]["lore"]["unit"] == "images/s"
    assert d["blocks"]["pp_rec"]["unit"] == "crops/s" and d["blocks"]["pp_rec"]["value"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "pages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_ocr_cascade():
    r = _run("--impl", "reference", "--cascade", "ocr", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.strip()][0])
    assert d["config"]["workload"].startswith("BASELINE configs[1]") and d["value"] > 0 and "lore" not in d["blocks"]


def test_reference_arm_other_ranks_exit_quietly():
    r = _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        return  # on a GPU box the arm runs; the driver measures it
    r = _run("--steps", "1", "--warmup", "1")
    assert r.returncode != 0 and r.stdout.strip() == "" and "no CPU fallback" in (r.stderr + r.stdout)
