#!/bin/bash
cd "$(dirname "$0")/.."
timeout -s KILL 400 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "smpl" 2>&1 | tail -3
timeout -s KILL 400 python -m pytest tests/test_gpu_e2e.py -m gpu -q -x 2>&1 | tail -3
B="--no-live --no-smpl --no-fold --no-train --no-fp32 --no-hmr --cpu-budget 0"
timeout -s KILL 200 python bench.py $B > gpurun_out/bench_fin.json 2> gpurun_out/bench_fin.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/bench_fin.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["stages_ms"])
P
