#!/bin/bash
cd "$(dirname "$0")/.."
timeout -s KILL 120 python scripts/ief_cluster_trace.py > gpurun_out/cl_trace.txt 2>&1
B=1 timeout -s KILL 120 python scripts/ief_cluster_trace.py > gpurun_out/cl_trace_b1.txt 2>&1
tail -30 gpurun_out/cl_trace.txt
