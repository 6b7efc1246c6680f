#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -3
timeout 600 python bench.py --config train --steps 10 --warmup 3 --cpu-budget 10 > gpurun_out/bench_train_fp32.json 2> gpurun_out/bench_train_fp32.err; echo "train bench exit=$?"; tail -3 gpurun_out/bench_train_fp32.err
timeout 600 python bench.py --config train --steps 10 --warmup 3 --cpu-budget 0 --tc-grads > gpurun_out/bench_train_tc.json 2> gpurun_out/bench_train_tc.err; echo "train bench (tc grads) exit=$?"; tail -3 gpurun_out/bench_train_tc.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_train_fp32.json", "gpurun_out/bench_train_tc.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["training"], d.get("cpu_baseline"))
    except Exception as e:
        print(f, "ERR", e)
PY
