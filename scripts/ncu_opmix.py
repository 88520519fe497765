"""Executed-instruction mix and stall reasons of one kernel launch from an ncu report (source page):

    python scripts/ncu_opmix.py report.ncu-rep [launch_skip] [kernel_regex]
"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
skip = sys.argv[2] if len(sys.argv) > 2 else "0"
kre = sys.argv[3] if len(sys.argv) > 3 else "k_step"
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
print(rows[0][1])
hdr = rows[1]
ix = {k: i for i, k in enumerate(hdr)}
ops, stalls = collections.Counter(), collections.Counter()
first = None
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    if r[ix["Address"]] == first:      # the page repeats when several launches match
        break
    first = first or r[ix["Address"]]
    try:
        ex = int(r[ix["Instructions Executed"]])
    except ValueError:
        continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]].strip())
    ops[m.group(2).split(".")[0] if m else "?"] += ex
    for k in hdr:
        if k.startswith("stall_") and "Not Issued" not in k:
            stalls[k] += int(r[ix[k]] or 0)
warps = max(int(r[ix["Instructions Executed"]]) for r in rows[2:4])
tot = sum(ops.values())
print(f"warps {warps}, warp-instructions {tot}, per warp {tot / warps:.0f}")
for op, c in ops.most_common(24):
    print(f"  {op:12s} {c / warps:8.1f} per thread")
ts = sum(stalls.values())
for k, c in stalls.most_common(8):
    print(f"  {k:24s} {100 * c / ts:5.1f} %")
