#!/bin/bash
cd "$(dirname "$0")/.."
B="--no-live --no-smpl --no-fold --no-train --no-fp32 --no-hmr --cpu-budget 0"
timeout -s KILL 200 python bench.py $B > gpurun_out/bench_2cta.json 2> gpurun_out/bench_2cta.err
TP_TC_NO_2CTA=1 timeout -s KILL 200 python bench.py $B > gpurun_out/bench_no2cta.json 2> gpurun_out/bench_no2cta.err
python - <<'P'
import json
for f in ("bench_2cta", "bench_no2cta"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["stages_ms"], d["max_vertex_err_m_vs_cpu_baseline"])
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/{f}.err").read()[-500:])
P
timeout -s KILL 300 python -m pytest tests/test_gpu_e2e.py -m gpu -x -q -k "golden or full_size or programmatic" 2>&1 | tail -3
