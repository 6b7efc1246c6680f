#!/bin/bash
# same-box A/B of "VAR set" vs "VAR unset": usage gpu_ab2.sh VAR [repeats]
mkdir -p gpurun_out
VAR=$1; REP=${2:-2}
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q --no-header -p no:cacheprovider -k "golden or full_size or edge" > gpurun_out/e2e.log 2>&1; echo "tests exit=$? $(tail -1 gpurun_out/e2e.log)"
for r in $(seq 1 $REP); do
 for v in unset set; do
  if [ $v = set ]; then export $VAR=1; else unset $VAR; fi
  timeout 600 python bench.py --steps 100 --warmup 10 --no-smpl --no-fold --cpu-budget 0.5 > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/ab.json").read().strip().splitlines()[-1])
print("$VAR $v", "ms/step", round(d["ms_per_step"],4), "median", round(d["step_ms"]["median"],4), {k:round(x,4) for k,x in d["stages_ms"].items() if k!="pack"}, "live", round(d["live"]["p50_ms"],4), round(d["live"]["windowed"]["p50_ms"],4))
PY
 done
done
