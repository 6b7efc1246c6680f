#!/bin/bash
# same-box sweep of one env var over values: gpu_ab3.sh VAR v1 v2 ...
mkdir -p gpurun_out
VAR=$1; shift
for v in "$@"; do
  export $VAR=$v
  timeout 300 python bench.py --steps 100 --warmup 10 --no-smpl --no-fold --no-live --cpu-budget 0.5 > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/ab.json").read().strip().splitlines()[-1])
print("$VAR=$v", "ms/step", round(d["ms_per_step"],4), "median", round(d["step_ms"]["median"],4), {k:round(x,4) for k,x in d["stages_ms"].items() if k!="pack"})
PY
done
