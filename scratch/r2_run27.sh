#!/bin/bash
# round 2, GPU call 27: octet plane search in the warp-per-query chunks (point-to-plane / NDT pipelines on small batches)
cd "$(dirname "$0")/.."
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $O/r2A_gpu_tests.log 2>&1; grep -n "passed\|failed" $O/r2A_gpu_tests.log
for S in 1; do
echo "== queue-driven kernel timeline S=$S lidar3d-ndt.yaml"
MLO_B200_LIB=$PWD/scratch/libmlo_b200_trace.so timeout 300 python scratch/trace_persistent.py $S lidar3d-ndt.yaml 2>&1 | tail -18
done > $O/r2A_trace_ndt.log 2>&1
cat $O/r2A_trace_ndt.log
run() { # S env args
  echo "== S=$1 $2 $3"
  env $2 timeout 600 python bench.py --sequences $1 $3 2> $O/r2A_last.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'scans/s', {k:round(v,3) for k,v in d['phases']['host_wall_timed_pass'].items()}, (d.get('quality') or {}).get('parity_vs_oracle'))" || tail -5 $O/r2A_last.err
}
{
run 8 X=1 "--workload ndt --scans 120 --cpu-scans 120"
run 8 X=1 "--workload sequence_pt2pl --scans 200 --cpu-scans 200"
run 1 X=1 "--workload ndt --scans 120 --no-cpu-baseline"
run 32 X=1 "--workload sequence_pt2pl --scans 100 --no-cpu-baseline"
} > $O/r2A_seq.log 2>&1
cut -c1-700 $O/r2A_seq.log
