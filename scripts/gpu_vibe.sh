#!/bin/bash
# VIBE bootstrap + >64-row parity on the GPU, then a timing of the released bootstrap config.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vibe.py tests/test_gpu_kernels.py -m gpu -q --no-header -p no:cacheprovider -k "vibe or unpack or Vibe" > gpurun_out/vibe.log 2>&1; echo "vibe exit=$? $(tail -1 gpurun_out/vibe.log)"
grep -E "^FAILED|^ERROR|Error:" gpurun_out/vibe.log | head -10
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q --no-header -p no:cacheprovider -k "edge_shapes" > gpurun_out/edge.log 2>&1; echo "edge exit=$? $(tail -1 gpurun_out/edge.log)"
grep -E "^FAILED|^ERROR|Error:" gpurun_out/edge.log | head -10
timeout 600 python scripts/vibe_time.py 2>&1 | tail -12
