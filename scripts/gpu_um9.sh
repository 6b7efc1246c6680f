#!/bin/bash
cd "$(dirname "$0")/.."
timeout -s KILL 200 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "umma or interleaved" 2>&1 | tail -3
timeout -s KILL 120 python scripts/umma_trace.py 2>&1 | tail -9 | cut -c1-260
timeout -s KILL 400 python -m pytest tests/test_gpu_e2e.py tests/test_stream.py tests/test_vibe.py -m gpu -x -q 2>&1 | tail -3
B="--no-live --no-smpl --no-fold --no-train --no-fp32 --no-hmr --cpu-budget 0"
timeout -s KILL 200 python bench.py $B > gpurun_out/bench_um9.json 2> gpurun_out/bench_um9.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/bench_um9.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["stages_ms"])
P
