#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -p no:cacheprovider -k "gru" > gpurun_out/gru.log 2>&1; echo "gru exit=$? $(tail -1 gpurun_out/gru.log)"
for cfg in "2 8" "3 8" "4 8" "2 4" "nores 0"; do
  set -- $cfg
  if [ $1 = nores ]; then export TP_GRU_NO_RES=1; else export TP_GRU_WS=$1 TP_GRU_RS=$2; fi
  timeout 600 python bench.py --steps 20 --warmup 3 --no-smpl --cpu-budget 1 > gpurun_out/bench_b1_$1_$2.json 2> gpurun_out/bench_ws.err; echo "bench ws=$1 rs=$2 exit=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_b1_$1_$2.json").read().strip().splitlines()[-1])
print("ws=$1 rs=$2 k2(B=32)", round(d["stages_ms"]["k2_recurrence_l0"],4), "live", round(d["live"]["p50_ms"],4), "windowed", round(d["live"]["windowed"]["p50_ms"],4))
PY
done
