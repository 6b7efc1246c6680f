#!/bin/bash
# per-kernel device time of one training step (ncu launch list; cold-cache, serialised: shares, not absolutes)
mkdir -p gpurun_out
NCU=$(command -v ncu || echo /usr/local/cuda/bin/ncu)
timeout 900 $NCU --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/train_launches.csv \
   python bench.py --config train --steps 2 --warmup 3 --cpu-budget 0 ${TRAIN_ARGS} > gpurun_out/train_launches.log 2>&1; echo "ncu exit=$?"
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/train_launches.csv", errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
data = [(r[kn], float(r[mv].replace(",", ""))) for r in rows[hdr + 1:] if len(r) > mv]
# last step = the last 1/5 of the launches (3 warm-up + 2 timed steps)
per = len(data) // 5
last = data[-per:]
agg = collections.OrderedDict()
for k, v in last:
    k = k.split("(")[0]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v for _, v in last)
print(f"launches in the last step: {len(last)}, total {tot/1e3:.1f} us")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v/1e3:9.1f} us  {100*v/tot:5.1f}%  x{n:3d}  {k[:90]}")
PY
