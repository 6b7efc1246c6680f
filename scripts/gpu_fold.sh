#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q --no-header -p no:cacheprovider -k "folded or live" -s > gpurun_out/fold.log 2>&1; echo "fold exit=$? $(tail -1 gpurun_out/fold.log)"
grep -E "^FAILED|^ERROR|Error:|folded|layer-by-layer" gpurun_out/fold.log | head -20
timeout 900 python bench.py --steps 50 --warmup 5 --no-smpl --cpu-budget 1 > gpurun_out/bench_fold.json 2> gpurun_out/bench_fold.err; echo "bench exit=$?"; tail -3 gpurun_out/bench_fold.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_fold.json").read().strip().splitlines()[-1])
print("value",round(d["value"]),"ms/step",round(d["ms_per_step"],4),"e2e",round(d["e2e"]["value"]))
print("folded",d.get("folded"))
print("live",d.get("live"))
PY
