#!/bin/bash
cd "$(dirname "$0")/.."
timeout -s KILL 200 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm_bf16_tcgen05" 2>&1 | tail -5
timeout -s KILL 300 python -m pytest tests/test_gpu_e2e.py -m gpu -x -q -k "golden or full_size or programmatic" 2>&1 | tail -4
timeout -s KILL 300 python bench.py > gpurun_out/bench_k1.json 2> gpurun_out/bench_k1.err; tail -c 300 gpurun_out/bench_k1.err
python - <<'P'
import json
for f in ("bench_k1",):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["stages_ms"], d["fp32"]["fp32_tc"]["ms_per_step"], d["training"]["ms_per_step"], d["smpl_standalone"]["ms"])
    except Exception as e:
        print(f, "failed", e)
P
