#!/bin/bash
# A/B of the resident-weight recurrence kernel against the streaming TMA kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -p no:cacheprovider -k "gru" > gpurun_out/gru.log 2>&1; echo "gru exit=$? $(tail -1 gpurun_out/gru.log)"
grep -E "^FAILED|^ERROR|Error:" gpurun_out/gru.log | head -10
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/e2e.log 2>&1; echo "e2e exit=$? $(tail -1 gpurun_out/e2e.log)"
grep -E "^FAILED|^ERROR|Error:" gpurun_out/e2e.log | head -10
for mode in res nores; do
  if [ $mode = nores ]; then export TP_GRU_NO_RES=1; else unset TP_GRU_NO_RES; fi
  timeout 600 python bench.py --steps 50 --warmup 5 --no-smpl --cpu-budget 1 > gpurun_out/bench_$mode.json 2> gpurun_out/bench_$mode.err; echo "bench $mode exit=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$mode.json").read().strip().splitlines()[-1])
print("$mode value",round(d["value"]),"ms/step",round(d["ms_per_step"],4),"e2e",round(d["e2e"]["value"]))
print({k:round(v,4) for k,v in d["stages_ms"].items()})
print("live",d.get("live"))
PY
done
