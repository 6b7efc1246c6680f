#!/bin/bash
cd "$(dirname "$0")/.."
timeout -s KILL 90 python scripts/heads_hang_debug.py > gpurun_out/hang_debug.txt 2>&1
grep -q completed gpurun_out/hang_debug.txt || { tail -5 gpurun_out/hang_debug.txt | cut -c1-300; exit 1; }
timeout -s KILL 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout -s KILL 300 python bench.py > gpurun_out/bench_cl.json 2> gpurun_out/bench_cl.err; tail -c 300 gpurun_out/bench_cl.err
