#!/bin/bash
mkdir -p gpurun_out
NCU=$(command -v ncu || echo /usr/local/cuda/bin/ncu)
timeout 300 $NCU --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/um2_launches.csv python scripts/smpl_standalone.py 65536 bf16 1 > gpurun_out/um2_launches.log 2>&1; echo "list exit=$?"
timeout 600 $NCU --set full --clock-control none --import-source on -k regex:k_smpl_lbs_um2 -s 2 -c 1 -f -o gpurun_out/prof_k_smpl_lbs_um2 \
   python scripts/smpl_standalone.py 16384 bf16 1 > gpurun_out/ncu_k_smpl_lbs_um2.log 2>&1; echo "ncu exit=$?"
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/um2_launches.csv")))
i0=next(i for i,r in enumerate(rows) if r and r[0]=="ID")
h=rows[i0]; kn=h.index("Kernel Name"); mv=h.index("Metric Value")
for r in rows[i0+1:][-5:]:
    print(r[kn][:70], r[mv])
PY
