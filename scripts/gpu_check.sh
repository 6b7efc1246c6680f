#!/bin/bash
# Runs the GPU parity suites in separate, time-limited processes (a deadlocked kernel then
# costs one suite, not the whole call) and leaves logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, pytest args...
  local name=$1 t=$2; shift 2
  timeout "$t" python -m pytest "$@" -q --no-header -p no:cacheprovider > "gpurun_out/$name.log" 2>&1
  echo "== $name exit=$? $(tail -1 gpurun_out/$name.log)"
}
run safe 600 tests/test_gpu_kernels.py -m gpu -k "geometry or gemm_f32 or pack_rows or smpl or skinny"
run gru 600 tests/test_gpu_kernels.py -m gpu -k "gru"
run tc 300 tests/test_gpu_kernels.py -m gpu -k "tcgen05"
run e2e 1200 tests/test_gpu_e2e.py -m gpu -s
grep -h -E "FAILED|ERROR|passed|failed|Error|error:" gpurun_out/*.log | sort | uniq -c | sort -rn | head -40
