#!/bin/bash
# round 2, GPU call 6: octet-per-query block kernel
cd "$(dirname "$0")/.."
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > $O/r2f_gpu_tests.log 2>&1; tail -8 $O/r2f_gpu_tests.log
rm -f $O/r2f_trace_block.log
for CLU in 1 8; do
  echo "== trace S=1 cluster=$CLU prior off" >> $O/r2f_trace_block.log
  MLO_ICP_PRIOR=0 MLO_BLOCK_CLUSTER=$CLU MLO_B200_LIB=$PWD/scratch/libmlo_b200_trace.so timeout 300 python scratch/trace_block.py 1 >> $O/r2f_trace_block.log 2>&1
done
echo "== trace S=1 cluster=8 with prior" >> $O/r2f_trace_block.log
MLO_BLOCK_CLUSTER=8 MLO_B200_LIB=$PWD/scratch/libmlo_b200_trace.so timeout 300 python scratch/trace_block.py 1 >> $O/r2f_trace_block.log 2>&1
cat $O/r2f_trace_block.log
run() { # S env args
  echo "== S=$1 $2 $3"
  env $2 timeout 400 python bench.py --sequences $1 --no-cpu-baseline $3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'scans/s', {k:round(v,3) for k,v in d['phases']['host_wall_timed_pass'].items()}, {k:round(v,3) for k,v in d['phases']['device_events_pass'].items()})"
}
{
run 1 MLO_ALIGN_PATH=2 "--workload sequence --scans 120"
run 1 MLO_ALIGN_PATH=3 "--workload sequence --scans 120"
run 1 "MLO_ALIGN_PATH=3 MLO_BLOCK_CLUSTER=4" "--workload sequence --scans 120"
run 32 MLO_ALIGN_PATH=2 "--workload sequence --scans 60"
run 32 MLO_ALIGN_PATH=3 "--workload sequence --scans 60"
run 32 "MLO_ALIGN_PATH=3 MLO_BLOCK_CLUSTER=2" "--workload sequence --scans 60"
run 64 MLO_ALIGN_PATH=3 "--workload sequence --scans 60"
run 128 MLO_ALIGN_PATH=3 "--workload sequence --scans 40"
run 8 MLO_ALIGN_PATH=2 "--workload ndt --scans 40"
run 8 MLO_ALIGN_PATH=3 "--workload ndt --scans 40"
} > $O/r2f_block_ab.log 2>&1
cat $O/r2f_block_ab.log
