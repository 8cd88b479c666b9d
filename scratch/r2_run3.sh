#!/bin/bash
# round 2, GPU call 3: where does the match phase of k_icp_block spend its time?
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r2c_trace_block.log
for CLU in 1 4 8; do
  echo "== trace S=1 cluster=$CLU" >> $O/r2c_trace_block.log
  MLO_BLOCK_CLUSTER=$CLU MLO_B200_LIB=$PWD/scratch/libmlo_b200_trace.so timeout 300 python scratch/trace_block.py 1 >> $O/r2c_trace_block.log 2>&1
done
cat $O/r2c_trace_block.log
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,pstate --format=csv -lms 100 > $O/r2c_clocks_S1.csv &
SMI=$!
MLO_ALIGN_PATH=3 timeout 300 python bench.py --sequences 1 --no-cpu-baseline --workload sequence --scans 200 2>/dev/null | cut -c1-300
kill $SMI
sort $O/r2c_clocks_S1.csv | uniq -c | sort -rn | head -8
