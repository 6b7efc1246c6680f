#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -p no:cacheprovider -k "smpl or tcgen05 or gemm" > gpurun_out/smpl.log 2>&1; echo "smpl+gemm exit=$? $(tail -1 gpurun_out/smpl.log)"
grep -E "^FAILED|^ERROR|Error|assert" gpurun_out/smpl.log | head
python scripts/smpl_standalone.py 65536 bf16 5
TP_SMPL_CHUNK=1024 python scripts/smpl_standalone.py 65536 bf16 5
python scripts/smpl_standalone.py 1024 bf16 20
timeout 600 python bench.py --steps 50 --warmup 5 --no-smpl --no-live --no-fold --cpu-budget 1 > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench exit=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_q.json").read().strip().splitlines()[-1])
print("value",round(d["value"]),"ms/step",round(d["ms_per_step"],4)); print({k:round(v,4) for k,v in d["stages_ms"].items()})
PY
