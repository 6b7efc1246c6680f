#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_e2e.py -m gpu -q --no-header -p no:cacheprovider -x -k "golden or full_size" > gpurun_out/tc_e2e.log 2>&1; echo "e2e exit=$? $(tail -1 gpurun_out/tc_e2e.log)"
grep -E "^FAILED|^ERROR|Error|assert " gpurun_out/tc_e2e.log | head
for prec in fp32 fp32_tc; do
timeout 300 python bench.py --steps 20 --warmup 5 --precision $prec --no-smpl --no-live --no-fold --no-train --cpu-budget 0 > gpurun_out/bench_$prec.json 2> gpurun_out/bench_$prec.err; echo "bench $prec exit=$?"; tail -2 gpurun_out/bench_$prec.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$prec.json").read().strip().splitlines()[-1])
print("$prec value",round(d["value"]),"ms/step",round(d["ms_per_step"],4)); print({k:round(v,4) for k,v in d["stages_ms"].items()})
PY
done
