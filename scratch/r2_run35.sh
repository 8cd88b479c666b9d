#!/bin/bash
# round 2, GPU call 35: pairings prefetched to shared memory during the solve + transposing reductions in the latency-bound chunks
cd "$(dirname "$0")/.."
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $O/r2I_gpu_tests.log 2>&1; grep -n "passed\|failed" $O/r2I_gpu_tests.log
echo "== queue-driven kernel timeline S=1" > $O/r2I_trace.log
MLO_B200_LIB=$PWD/scratch/libmlo_b200_trace.so timeout 300 python scratch/trace_persistent.py 1 >> $O/r2I_trace.log 2>&1
tail -20 $O/r2I_trace.log
run() { # S env args
  echo "== S=$1 $2 $3"
  env $2 timeout 600 python bench.py --sequences $1 $3 2> $O/r2I_last.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'scans/s', {k:round(v,3) for k,v in d['phases']['host_wall_timed_pass'].items()}, (d.get('quality') or {}).get('parity_vs_oracle'))" || tail -5 $O/r2I_last.err
}
{
run 1 X=1 "--workload sequence --scans 300 --cpu-scans 300"
run 32 X=1 "--workload sequence --scans 200 --cpu-scans 200"
run 8 X=1 "--workload ndt --scans 120 --no-cpu-baseline"
run 128 X=1 "--workload sequence --scans 40 --no-cpu-baseline"
} > $O/r2I_seq.log 2>&1
cut -c1-600 $O/r2I_seq.log
timeout 600 python bench.py --steps 10 --no-cpu-baseline --sub-records none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],3),'launch_us',round(d['roofline']['avg_launch_us'],1),'launches',d['gpu_launches'])"
(time timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_paths.py -x -q -k "every_align_path") > $O/r2I_racecheck.log 2>&1; grep -n "passed\|failed\|SUMMARY" $O/r2I_racecheck.log
