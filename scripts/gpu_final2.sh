#!/bin/bash
cd "$(dirname "$0")/.."
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout -s KILL 1000 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout -s KILL 400 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench exit=$?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/final_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["traffic"], d["roofline"]["frac"], d["gpu_launches"])
P
