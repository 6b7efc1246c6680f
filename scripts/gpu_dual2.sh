#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -p no:cacheprovider -x -k "gru" > gpurun_out/gru.log 2>&1; echo "gru exit=$? $(tail -1 gpurun_out/gru.log)"
grep -E "^FAILED|^ERROR|Error|assert" gpurun_out/gru.log | head -10
timeout 600 python -m pytest tests/test_gpu_e2e.py tests/test_stream.py tests/test_vibe.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/e2e.log 2>&1; echo "e2e exit=$? $(tail -1 gpurun_out/e2e.log)"
grep -E "^FAILED|^ERROR|Error:" gpurun_out/e2e.log | head -10
for v in unset set; do
  if [ $v = set ]; then export TP_GRU_NO_DUAL=1; else unset TP_GRU_NO_DUAL; fi
  timeout 300 python bench.py --steps 100 --warmup 10 --no-smpl --no-fold --cpu-budget 0.5 > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/ab.json").read().strip().splitlines()[-1])
print("NO_DUAL $v", "ms/step", round(d["ms_per_step"],4), {k:round(x,4) for k,x in d["stages_ms"].items() if k!="pack"}, "live", round(d["live"]["p50_ms"],4), "windowed", round(d["live"]["windowed"]["p50_ms"],4))
PY
done
