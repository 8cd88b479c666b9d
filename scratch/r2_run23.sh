#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
for Y in lidar3d-ndt.yaml; do
for S in 1 8; do
echo "== queue-driven kernel timeline S=$S $Y"
MLO_B200_LIB=$PWD/scratch/libmlo_b200_trace.so timeout 300 python scratch/trace_persistent.py $S $Y 2>&1 | tail -22
done; done > $O/r2w_trace_ndt.log 2>&1
cat $O/r2w_trace_ndt.log
