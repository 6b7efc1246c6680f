#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -k "umma" -q --no-header -p no:cacheprovider -x 2>&1 | tail -3
timeout 300 python scripts/umma_trace.py 2>&1 | tail -9
timeout 600 python bench.py --steps 20 --warmup 5 --cpu-budget 1 --no-live --no-fold --no-smpl > gpurun_out/bench_umma.json 2> gpurun_out/bench_umma.err; echo "bench exit=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_umma.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d.get("stages_ms"))
PY
