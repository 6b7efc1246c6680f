#!/bin/bash
cd "$(dirname "$0")/.."
timeout -s KILL 90 python scripts/heads_hang_debug.py > gpurun_out/hang_debug.txt 2>&1
grep -q completed gpurun_out/hang_debug.txt || { tail -5 gpurun_out/hang_debug.txt | cut -c1-300; exit 1; }
timeout -s KILL 120 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "ief_cluster" 2>&1 | tail -3
timeout -s KILL 300 python -m pytest tests/test_gpu_e2e.py -m gpu -x -q -k "golden or programmatic or fused_heads or fence" 2>&1 | tail -3
scripts/gpu_cl1.sh | grep -v k_heads_base
grep "k_ief_cluster cta 0" -A3 gpurun_out/cl_trace_b1.txt
