#!/bin/bash
# first contact of k_gru_umma with a GPU: debug probe (also with every A operand in shared memory), then the tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python scripts/umma_debug.py > gpurun_out/umma_debug.log 2>&1; echo "debug exit=$?"
TP_UM_TMEMK=0 timeout 300 python scripts/umma_debug.py > gpurun_out/umma_debug_smemA.log 2>&1; echo "debug(smem A) exit=$?"
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -k "umma" -q --no-header -p no:cacheprovider -x > gpurun_out/umma_test.log 2>&1; echo "test exit=$?"
tail -5 gpurun_out/umma_debug.log; tail -5 gpurun_out/umma_debug_smemA.log; tail -15 gpurun_out/umma_test.log
