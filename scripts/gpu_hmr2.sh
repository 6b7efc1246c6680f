#!/bin/bash
cd "$(dirname "$0")/.."
timeout -s KILL 200 python -m pytest tests/test_hmr.py -m gpu -q 2>&1 | tail -2
scripts/gpu_hmr_list.sh
