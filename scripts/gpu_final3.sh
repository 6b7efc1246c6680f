#!/bin/bash
cd "$(dirname "$0")/.."
timeout -s KILL 1000 python -m pytest tests -m gpu -q 2>&1 | tail -3
bash scripts/gpu_round2b_profiles.sh
