#!/bin/bash
# round 2, GPU call 1: parity of every align path, block-kernel timeline, block vs queue A/B on fleets, headline bench
cd "$(dirname "$0")/.."
O=gpurun_out
nproc > $O/r2_nproc.txt; nvidia-smi topo -m > $O/r2_topo.txt 2>&1; free -g >> $O/r2_nproc.txt
(time timeout 1200 python -m pytest tests -m gpu -x -q) > $O/r2_gpu_tests.log 2>&1; tail -5 $O/r2_gpu_tests.log
for S in 1 32; do
  echo "== trace S=$S" >> $O/r2_trace_block.log
  MLO_B200_LIB=$PWD/scratch/libmlo_b200_trace.so timeout 300 python scratch/trace_block.py $S >> $O/r2_trace_block.log 2>&1
done
tail -12 $O/r2_trace_block.log
run() { # S env args
  echo "== S=$1 $2 $3"
  env $2 timeout 400 python bench.py --sequences $1 --no-cpu-baseline $3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'scans/s', {k:round(v,3) for k,v in d['phases']['host_wall_timed_pass'].items()}, {k:round(v,3) for k,v in d['phases']['device_events_pass'].items()})"
}
{
run 1 MLO_ALIGN_PATH=2 "--workload sequence --scans 120"
run 1 MLO_ALIGN_PATH=3 "--workload sequence --scans 120"
run 1 "MLO_ALIGN_PATH=3 MLO_BLOCK_THREADS=256" "--workload sequence --scans 120"
run 32 MLO_ALIGN_PATH=2 "--workload sequence --scans 60"
run 32 MLO_ALIGN_PATH=3 "--workload sequence --scans 60"
run 64 MLO_ALIGN_PATH=3 "--workload sequence --scans 60"
run 128 MLO_ALIGN_PATH=3 "--workload sequence --scans 40"
run 8 MLO_ALIGN_PATH=2 "--workload ndt --scans 40"
run 8 MLO_ALIGN_PATH=3 "--workload ndt --scans 40"
} > $O/r2_block_ab.log 2>&1
cat $O/r2_block_ab.log
(time timeout 600 python bench.py --steps 10) > $O/r2_bench_default.json 2> $O/r2_bench_default.err; cat $O/r2_bench_default.json; tail -3 $O/r2_bench_default.err
(time timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_paths.py -x -q -k "every_align_path or block_kernel") > $O/r2_sanitizer_memcheck_paths.log 2>&1; tail -4 $O/r2_sanitizer_memcheck_paths.log
