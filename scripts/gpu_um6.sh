#!/bin/bash
mkdir -p gpurun_out
NCU=$(command -v ncu || echo /usr/local/cuda/bin/ncu)
timeout 400 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -p no:cacheprovider -x -k "smpl and not older" > gpurun_out/um_smpl.log 2>&1; echo "smpl tests exit=$? $(tail -1 gpurun_out/um_smpl.log)"
grep -E "^FAILED|^ERROR|Error|assert|^E " gpurun_out/um_smpl.log | head -12
timeout 120 python scripts/smpl_standalone.py 65536 bf16 5 random
timeout 300 $NCU --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/um2_launches.csv python scripts/smpl_standalone.py 65536 bf16 1 > gpurun_out/um2_launches.log 2>&1; echo "list exit=$?"
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/um2_launches.csv")))
i0=next(i for i,r in enumerate(rows) if r and r[0]=="ID")
h=rows[i0]; kn=h.index("Kernel Name"); mv=h.index("Metric Value")
for r in rows[i0+1:][-5:]:
    print(r[kn][:70], r[mv])
PY
