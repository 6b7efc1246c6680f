#!/bin/bash
# full GPU parity suites + traces + bench (bf16) in one call
bash scripts/gpu_check.sh
timeout 300 python scripts/ief_trace.py 2>&1 | grep -E "cta 0|layer [0-3]:|span|regressor" | head -8
timeout 300 python scripts/gru_trace.py 2>&1 | grep -E "cta   0|cta 100"
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench bf16 exit=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_bf16.json").read().strip().splitlines()[-1])
print("value",round(d["value"]),"ms/step",round(d["ms_per_step"],4),"e2e",round(d["e2e"]["value"]))
print({k:round(v,4) for k,v in d["stages_ms"].items()})
print("live",d.get("live")); print("smpl",d.get("smpl_standalone"))
PY
tail -3 gpurun_out/bench_bf16.err
