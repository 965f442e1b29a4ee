"""Print a fixed set of headline metrics (and the top stalled instructions) from an .ncu-rep: tuning aid.
usage: python tools/ncu_keys.py file.ncu-rep [kernel-index]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "inst_executed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__t_bytes.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_lg_throttle",
        "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_mio_throttle",
        "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_selected",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_wait",
        "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_branch_resolving",
        "smsp__pcsamp_warps_issue_stalled_no_instructions", "smsp__pcsamp_warps_issue_stalled_tex_throttle",
        "smsp__pcsamp_warps_issue_stalled_dispatch_stall"]
for w in WANT:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:70s}", [r[i][:60] for r in rows[2:]])
kid = sys.argv[2] if len(sys.argv) > 2 else "1"
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


h = next((r for r in rows if "Source" in r and "# Samples" in r), None)
if h is not None:
    isrc, ismp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    st = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    data = [r for r in rows[rows.index(h) + 1:] if len(r) == len(h)]  # ncu emits ragged separator rows
    print("total samples", sum(num(r[ismp]) for r in data))
    for r in sorted(data, key=lambda r: -num(r[ismp]))[:22]:
        top = sorted(((h[i], num(r[i])) for i in st if num(r[i]) > 0), key=lambda kv: -kv[1])[:3]
        print(r[ismp].rjust(6), r[iex].rjust(9), r[isrc][:64].ljust(64), top)
