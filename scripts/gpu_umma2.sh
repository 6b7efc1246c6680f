#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_check.sh 2>&1 | tail -12
timeout 600 python bench.py --steps 20 --warmup 5 --cpu-budget 3 > gpurun_out/bench_umma.json 2> gpurun_out/bench_umma.err; echo "bench exit=$?"
TP_GRU_NO_UMMA=1 timeout 600 python bench.py --steps 20 --warmup 5 --cpu-budget 1 --no-live --no-fold --no-smpl > gpurun_out/bench_noumma.json 2> gpurun_out/bench_noumma.err; echo "bench(no umma) exit=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_umma.json", "gpurun_out/bench_noumma.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d.get("stages_ms"), d.get("live"), d.get("released_config"))
    except Exception as e:
        print(f, "ERR", e)
PY
