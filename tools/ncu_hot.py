#!/usr/bin/env python
"""Top stall-sample SASS lines of one kernel in an ncu report (needs --import-source on / -lineinfo).
   python tools/ncu_hot.py report.ncu-rep kernel_regex [top_n] [launch_skip]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
print(rows[0][:2])
S, N, I = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = [(int(r[N] or 0), int(r[I] or 0), k, r[S]) for k, r in enumerate(rows[hi + 1:]) if len(r) == len(hdr)]
tot = sum(d[0] for d in data)
print("total samples", tot, "instr lines", len(data), "warp-instr executed", sum(d[1] for d in data))
for s, i, k, src in sorted(data, reverse=True)[:top]:
    print(f"{100*s/max(tot,1):5.1f}%  samples={s:6d} exec={i:9d}  line#{k:4d}  {src.strip()[:110]}")
