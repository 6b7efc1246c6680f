#!/bin/bash
cd "$(dirname "$0")/.."
timeout -s KILL 400 python bench.py > gpurun_out/bench_b1.json 2> gpurun_out/bench_b1.err; tail -c 400 gpurun_out/bench_b1.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/bench_b1.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["stages_ms"])
print(d["hmr_feature_extractor"])
P
