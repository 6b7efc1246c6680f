#!/usr/bin/env python
"""Prints the kernels of the last forward in an ncu launch list (gpu__time_duration.sum)."""
import csv, sys
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
names = [r["Kernel Name"] for r in rows]
starts = [i for i, n in enumerate(names) if "k_pack_rows" in n]
start = starts[-1]
if not any("k_smpl_finalize" in n for n in names[start:]) and len(starts) > 1:      # the capture stopped inside the last forward
    start = starts[-2]
end = len(rows)
tot = 0.0
for r in rows[start:end]:
    if "k_pack_rows" in r["Kernel Name"] and r is not rows[start]:
        break
    d = float(r["Metric Value"]) / 1000
    tot += d
    print(f'{r["Kernel Name"][:58]:58s} grid={r["Grid Size"]:>14s} blk={r["Block Size"]:>12s} {d:8.1f} us')
print(f"total {tot:.1f} us")
