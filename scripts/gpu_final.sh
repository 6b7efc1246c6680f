#!/bin/bash
# round-end evidence run: smoke, full suites, both bench arms, both precisions
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$? $(tail -1 gpurun_out/smoke.log)"
timeout 1500 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit=$? $(tail -1 gpurun_out/pytest_gpu.log)"
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit=$?"
timeout 900 python bench.py > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench bf16 exit=$?"
timeout 900 python bench.py --precision fp32 --no-live --no-smpl > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "bench fp32 exit=$?"
python - <<'PY'
import json
for f in ("bench_ref","bench_bf16","bench_fp32"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value",round(d["value"]),"ms/step",round(d["ms_per_step"],4),"e2e",round(d["e2e"]["value"]), d.get("clocks"))
    except Exception as e: print(f, "ERR", e)
PY
