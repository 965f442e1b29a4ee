#!/usr/bin/env python
"""Where a warp-specialised kernel waits: every sampled SASS instruction with more than N samples, in address order, with
the nearest preceding 'landmark' (TMA / MMA / TMEM / barrier instruction) so that an inlined mbarrier wait can be told
apart by its call site.  usage: python tools/ncu_sync_sites.py file.ncu-rep kernel-id [min-samples]"""
import csv
import subprocess
import sys

rep, kid = sys.argv[1], sys.argv[2]
thr = int(sys.argv[3]) if len(sys.argv) > 3 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = next(r for r in rows if "Source" in r and "# Samples" in r)
isrc, ismp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
data = [r for r in rows[rows.index(h) + 1:] if len(r) == len(h)]
data = data[:len(data) // 2] if len(data) > 2000 and data[0][isrc] == data[len(data) // 2][isrc] else data  # ncu lists the function twice


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


print("total samples", sum(num(r[ismp]) for r in data))
MARK = ("UTMALDG", "UTCHMMA", "UTCBAR", "LDTM", "BAR.SYNC", "STS.128", "LDG.E.128", "STG.E.128", "FENCE", "SYNCS.ARRIVE", "MUFU")
last = ""
for k, r in enumerate(data):
    t = r[isrc].strip()
    if any(m in t for m in MARK):
        last = f"{k}:{t[:40]}"
    if num(r[ismp]) >= thr:
        print(f"{k:5d} {r[ismp]:>6s} {r[iex]:>9s}  {t[:70]:70s} after [{last}]")
