#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_loss.py tests/test_gpu_train.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/loss.log 2>&1; echo "loss+train exit=$? $(tail -1 gpurun_out/loss.log)"
grep -E "^FAILED|^ERROR|Error|assert|^E " gpurun_out/loss.log | head -12
