#!/bin/bash
# full GPU regression: parity suites + default bench line (+ the same without the tcgen05 recurrence, for the A/B)
mkdir -p gpurun_out
bash scripts/gpu_check.sh 2>&1 | tail -8
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_full.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
print("stages", d.get("stages_ms"))
print("live", d.get("live"))
print("released", d.get("released_config"))
print("folded", d.get("folded", {}).get("ms_per_step"))
print("smpl", d.get("smpl_standalone"))
PY
