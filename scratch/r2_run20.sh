#!/bin/bash
# round 2, GPU call 20: warp-parallel prior Jacobian + stage-in overlapped with the partial sum
cd "$(dirname "$0")/.."
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $O/r2t_gpu_tests.log 2>&1; tail -4 $O/r2t_gpu_tests.log
echo "== queue-driven kernel timeline S=1" > $O/r2t_trace.log
MLO_B200_LIB=$PWD/scratch/libmlo_b200_trace.so timeout 300 python scratch/trace_persistent.py 1 >> $O/r2t_trace.log 2>&1
tail -22 $O/r2t_trace.log
run() { # S env args
  echo "== S=$1 $2 $3"
  env $2 timeout 600 python bench.py --sequences $1 $3 2> $O/r2t_last.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'scans/s', {k:round(v,3) for k,v in d['phases']['host_wall_timed_pass'].items()}, d.get('quality'))" || tail -5 $O/r2t_last.err
}
{
run 1 X=1 "--workload sequence --scans 300 --no-cpu-baseline"
run 32 X=1 "--workload sequence --scans 150 --cpu-scans 150"
run 8 X=1 "--workload ndt --scans 80 --no-cpu-baseline"
} > $O/r2t_seq.log 2>&1
cut -c1-600 $O/r2t_seq.log
timeout 600 python bench.py --steps 10 --no-cpu-baseline --sub-records none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],3),'launch_us',round(d['roofline']['avg_launch_us'],1),'launches',d['gpu_launches'])"
