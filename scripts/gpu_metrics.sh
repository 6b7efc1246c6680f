#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_metrics.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/metrics.log 2>&1; echo "metrics exit=$? $(tail -1 gpurun_out/metrics.log)"
grep -E "^FAILED|^ERROR|Error|assert" gpurun_out/metrics.log | head -20
timeout 300 python scripts/metrics_time.py 2>&1 | tail -6
