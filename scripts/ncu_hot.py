#!/usr/bin/env python
"""Summarise an ncu report's SASS source page: top instructions by stall samples with the
dominant stall reasons.   usage: ncu_hot.py report.ncu-rep [topN]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for k, r in enumerate(rows[2:]):
    try:
        s = int(r[ix["# Samples"]])
    except Exception:
        continue
    data.append((s, k, r))
total = sum(d[0] for d in data)
print(f"total samples {total}")
agg = {h: 0 for h in stalls}
for s, k, r in data:
    for h in stalls:
        agg[h] += int(r[ix[h]] or 0)
print("stall mix:", ", ".join(f"{h[6:]}={100*v/max(total,1):.1f}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for s, k, r in sorted(data, reverse=True)[:top]:
    why = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
    print(f"{100*s/max(total,1):5.1f}%  #{k:5d} {r[ix['Source']].strip()[:70]:70s} {why}")
