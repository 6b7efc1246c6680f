#!/bin/bash
cd "$(dirname "$0")/.."
timeout -s KILL 120 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm_bf16_tcgen05" 2>&1 | tail -8
