#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_stream.py tests/test_vibe.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/stream.log 2>&1; echo "stream exit=$? $(tail -1 gpurun_out/stream.log)"
grep -E "^FAILED|^ERROR|Error:" gpurun_out/stream.log | head -10
timeout 900 python bench.py --steps 50 --warmup 5 --no-smpl > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench bf16 exit=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_bf16.json").read().strip().splitlines()[-1])
print("value",round(d["value"]),"ms/step",round(d["ms_per_step"],4),"e2e",round(d["e2e"]["value"]))
print("live",d.get("live"))
PY
tail -5 gpurun_out/bench_bf16.err
