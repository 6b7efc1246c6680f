#!/bin/bash
# whole GPU suite (the driver's command) + the default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$? $(tail -1 gpurun_out/pytest_gpu.log)"
grep -E "^FAILED|^ERROR|Error|assert " gpurun_out/pytest_gpu.log | head
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit=$?"; tail -3 gpurun_out/bench_default.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "lean", round(d["e2e_lean"]["value"]))
print("legs", d["e2e_legs_ms"])
print("stages", d.get("stages_ms"))
print("roofline", {k: d["roofline"][k] for k in ("kernel", "frac", "traffic", "l2_to_sm")})
print("fp32", d.get("fp32")); print("acc", d.get("max_vertex_err_m_vs_cpu_baseline"))
print("smpl", d.get("smpl_standalone"))
print("train", d.get("training"))
print("cpu", d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("kind"))
PY
