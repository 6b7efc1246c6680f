#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q --no-header -p no:cacheprovider -s 2>&1 | grep -E "passed|failed|full-size|Error|assert" | tail -12
for cfg in "fp32:" "bf16:--tc-grads"; do
  prec=${cfg%%:*}; extra=${cfg#*:}
  timeout 600 python bench.py --config train --steps 10 --warmup 3 --cpu-budget 0 --train-precision $prec $extra > gpurun_out/bench_train_$prec.json 2> gpurun_out/bench_train_$prec.err; echo "train bench $prec exit=$?"; tail -2 gpurun_out/bench_train_$prec.err
done
python - <<'PY'
import json
for f in ("gpurun_out/bench_train_fp32.json", "gpurun_out/bench_train_bf16.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["training"]["loss"], d["training"]["launches_per_step"])
    except Exception as e:
        print(f, "ERR", e)
PY
