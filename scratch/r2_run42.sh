#!/bin/bash
# round 2, GPU call 42: work queue with one 64-bit word per slot (sequence | item)
cd "$(dirname "$0")/.."
O=gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q -x) > $O/r2P_gpu_tests.log 2>&1; grep -n "passed\|failed\|^FAILED" $O/r2P_gpu_tests.log
echo "== queue-driven kernel timeline S=1" > $O/r2P_trace.log
MLO_B200_LIB=$PWD/scratch/libmlo_b200_trace.so timeout 300 python scratch/trace_persistent.py 1 >> $O/r2P_trace.log 2>&1
grep "next_published\|total span\|pop_match0 -> chunk0" $O/r2P_trace.log
run() { # S env args
  echo "== S=$1 $2 $3"
  env $2 timeout 300 python bench.py --sequences $1 $3 2> $O/r2P_last.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'scans/s', {k:round(v,3) for k,v in d['phases']['host_wall_timed_pass'].items()})" || tail -5 $O/r2P_last.err
}
{
run 1 X=1 "--workload sequence --scans 300 --no-cpu-baseline"
run 32 X=1 "--workload sequence --scans 200 --no-cpu-baseline"
run 128 X=1 "--workload sequence --scans 40 --no-cpu-baseline"
} > $O/r2P_seq.log 2>&1
cut -c1-330 $O/r2P_seq.log
(time timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_paths.py -x -q -k "every_align_path") > $O/r2P_racecheck.log 2>&1; grep -n "passed\|failed\|SUMMARY" $O/r2P_racecheck.log
