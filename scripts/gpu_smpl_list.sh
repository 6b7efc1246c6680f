#!/bin/bash
mkdir -p gpurun_out
NCU=$(command -v ncu || echo /usr/local/cuda/bin/ncu)
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/smpl_launches.csv python scripts/smpl_standalone.py 16384 bf16 1 > gpurun_out/smpl_list.log 2>&1; echo "exit=$?"
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/smpl_launches.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    k = r[ki][:40]; v = float(r[vi].replace(",", ""))
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
for k, (c, t) in agg.items(): print(f"{k:42s} n={c:5d} total={t/1e6:9.3f} ms avg={t/c/1e3:8.1f} us")
PY
