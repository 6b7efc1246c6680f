#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -p no:cacheprovider -x -k "gemm or tcgen05 or smpl" > gpurun_out/g.log 2>&1; echo "gemm exit=$? $(tail -1 gpurun_out/g.log)"
grep -E "^FAILED|^ERROR|Error|assert" gpurun_out/g.log | head -10
timeout 600 python -m pytest tests/test_gpu_e2e.py tests/test_vibe.py -m gpu -q --no-header -p no:cacheprovider -k "golden or full_size or edge or vibe" > gpurun_out/e2e.log 2>&1; echo "e2e exit=$? $(tail -1 gpurun_out/e2e.log)"
grep -E "^FAILED|^ERROR|Error:" gpurun_out/e2e.log | head -10
for r in 1 2; do
 for v in unset set; do
  if [ $v = set ]; then export TP_TC_TALL=1; else unset TP_TC_TALL; fi
  timeout 300 python bench.py --steps 100 --warmup 10 --no-smpl --no-fold --no-live --cpu-budget 0.5 > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/ab.json").read().strip().splitlines()[-1])
print("NO_TALL $v", "ms/step", round(d["ms_per_step"],4), {k:round(x,4) for k,x in d["stages_ms"].items() if k!="pack"})
PY
 done
done
python scripts/smpl_standalone.py 65536 bf16 5
TP_TC_TALL=1 python scripts/smpl_standalone.py 65536 bf16 5
