#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_kernels.py -m gpu -q --no-header -p no:cacheprovider -k "smpl or golden or full_size or graph or pipeline or live or gemm or tcgen05 or edge or gru" > gpurun_out/e2e.log 2>&1; echo "tests exit=$? $(tail -1 gpurun_out/e2e.log)"
grep -E "^FAILED|^ERROR|Error:" gpurun_out/e2e.log | head -10
timeout 600 python bench.py --steps 50 --warmup 5 --no-smpl --no-fold --cpu-budget 1 > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench exit=$?"; tail -2 gpurun_out/bench_q.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_q.json").read().strip().splitlines()[-1])
print("value",round(d["value"]),"ms/step",round(d["ms_per_step"],4),"e2e",round(d["e2e"]["value"]), "launches", d["gpu_launches"])
print({k:round(v,4) for k,v in d["stages_ms"].items()})
print("live",round(d["live"]["p50_ms"],4), "windowed", round(d["live"]["windowed"]["p50_ms"],4))
PY
