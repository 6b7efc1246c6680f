#!/bin/bash
# same-box A/B of an environment switch: usage gpu_ab.sh VAR valA valB [repeats]
mkdir -p gpurun_out
VAR=$1; A=$2; B=$3; REP=${4:-2}
for r in $(seq 1 $REP); do
 for v in $A $B; do
  export $VAR=$v
  timeout 600 python bench.py --steps 100 --warmup 10 --no-smpl --no-fold --no-live --cpu-budget 0.5 > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/ab.json").read().strip().splitlines()[-1])
print("$VAR=$v", "ms/step", round(d["ms_per_step"],4), "median", round(d["step_ms"]["median"],4), {k:round(x,4) for k,x in d["stages_ms"].items() if k!="pack"})
PY
 done
done
