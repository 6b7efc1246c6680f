#!/bin/bash
mkdir -p gpurun_out
echo "== resident kernel ws=5 rs=2"; TP_GRU_WS=5 TP_GRU_RS=2 timeout 300 python scripts/gru_trace.py 2>&1 | tail -8
echo "== resident kernel ws=2 rs=5"; TP_GRU_WS=2 TP_GRU_RS=5 timeout 300 python scripts/gru_trace.py 2>&1 | tail -8
echo "== streaming kernel"; TP_GRU_NO_RES=1 timeout 300 python scripts/gru_trace.py 2>&1 | tail -8
