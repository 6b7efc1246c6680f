#!/bin/bash
cd "$(dirname "$0")/.."
timeout -s KILL 90 python scripts/heads_hang_debug.py > gpurun_out/hang_debug.txt 2>&1
grep -q completed gpurun_out/hang_debug.txt || { tail -5 gpurun_out/hang_debug.txt | cut -c1-300; exit 1; }
timeout -s KILL 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout -s KILL 400 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 300 gpurun_out/bench_full.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/bench_full.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["stages_ms"])
print(d["live"]["p50_ms"], d["live"]["windowed"]["p50_ms"], d["released_config"]["ms_per_step"], d["folded"]["ms_per_step"], d["fp32"]["fp32_tc"]["ms_per_step"])
P
