#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_stream.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/e2e.log 2>&1; echo "e2e exit=$? $(tail -1 gpurun_out/e2e.log)"
grep -E "^FAILED|^ERROR|Error:" gpurun_out/e2e.log | head -10
for mode in fused nofused; do
  if [ $mode = nofused ]; then export TP_NO_FUSED_HEADS=1; else unset TP_NO_FUSED_HEADS; fi
  timeout 600 python bench.py --steps 50 --warmup 5 --no-smpl --no-fold --cpu-budget 1 > gpurun_out/bench_$mode.json 2> gpurun_out/bench_$mode.err; echo "bench $mode exit=$?"; tail -2 gpurun_out/bench_$mode.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$mode.json").read().strip().splitlines()[-1])
print("$mode value",round(d["value"]),"ms/step",round(d["ms_per_step"],4),"e2e",round(d["e2e"]["value"]), "launches", d["gpu_launches"])
print({k:round(v,4) for k,v in d["stages_ms"].items()})
print("live",round(d["live"]["p50_ms"],4), "windowed", round(d["live"]["windowed"]["p50_ms"],4))
PY
done
