#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/e2e.log 2>&1; echo "e2e exit=$? $(tail -1 gpurun_out/e2e.log)"
grep -E "^FAILED|^ERROR|Error:" gpurun_out/e2e.log | head -10
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench bf16 exit=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_bf16.json").read().strip().splitlines()[-1])
print("value",round(d["value"]),"ms/step",round(d["ms_per_step"],4),"e2e",round(d["e2e"]["value"]))
print({k:round(v,4) for k,v in d["stages_ms"].items()})
print("live",d.get("live")); print("smpl",d.get("smpl_standalone")); print("cpu",d.get("cpu_baseline")); print("roof",d.get("roofline"))
PY
tail -5 gpurun_out/bench_bf16.err
