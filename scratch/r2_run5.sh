#!/bin/bash
# round 2, GPU call 5: block-kernel trace + ncu source-level capture of k_icp_block (S=1), new tests
cd "$(dirname "$0")/.."
O=gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -x -q) > $O/r2e_gpu_tests.log 2>&1; tail -5 $O/r2e_gpu_tests.log
rm -f $O/r2e_trace_block.log
for CLU in 1 8; do
  echo "== trace S=1 cluster=$CLU" >> $O/r2e_trace_block.log
  MLO_ICP_PRIOR=0 MLO_BLOCK_CLUSTER=$CLU MLO_B200_LIB=$PWD/scratch/libmlo_b200_trace.so timeout 300 python scratch/trace_block.py 1 >> $O/r2e_trace_block.log 2>&1
done
cat $O/r2e_trace_block.log
MLO_ALIGN_PATH=3 MLO_BLOCK_CLUSTER=1 MLO_ICP_PRIOR=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_icp_block -s 20 -c 1 \
   -o $O/r2e_prof_icp_block_S1 -f python bench.py --sequences 1 --no-cpu-baseline --workload sequence --scans 16 > $O/r2e_ncu_block.log 2>&1
ls -la $O/r2e_prof_icp_block_S1.ncu-rep; tail -3 $O/r2e_ncu_block.log
