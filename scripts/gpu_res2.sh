#!/bin/bash
mkdir -p gpurun_out
for cfg in "7 0" "6 1" "5 2" "6 0" "5 0"; do
  set -- $cfg
  export TP_GRU_WS=$1 TP_GRU_RS=$2
  timeout 600 python bench.py --steps 50 --warmup 5 --no-smpl --no-live --cpu-budget 1 > gpurun_out/bench_ws$1_rs$2.json 2> gpurun_out/bench_ws.err; echo "bench ws=$1 rs=$2 exit=$?"; tail -2 gpurun_out/bench_ws.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_ws$1_rs$2.json").read().strip().splitlines()[-1])
print("ws=$1 rs=$2 value",round(d["value"]),"ms/step",round(d["ms_per_step"],4), "k2", round(d["stages_ms"]["k2_recurrence_l0"],4))
PY
done
