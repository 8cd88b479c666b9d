#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
echo "== queue-driven kernel timeline S=1 (prior split)" > $O/r2J_trace.log
MLO_B200_LIB=$PWD/scratch/libmlo_b200_trace.so timeout 300 python scratch/trace_persistent.py 1 >> $O/r2J_trace.log 2>&1
tail -22 $O/r2J_trace.log
