#!/bin/bash
# A/B on one box: current library vs the build before the octet plane search (default pipeline, 32 sequences)
cd "$(dirname "$0")/.."
O=gpurun_out
run() { # S env args
  echo "== S=$1 $2 $3"
  env $2 timeout 600 python bench.py --sequences $1 $3 2> $O/r2C_last.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'scans/s', {k:round(v,3) for k,v in d['phases']['host_wall_timed_pass'].items()})" || tail -5 $O/r2C_last.err
}
{
run 32 X=1 "--workload sequence --scans 200 --no-cpu-baseline"
run 32 MLO_B200_LIB=$PWD/scratch/libmlo_b200_prev.so "--workload sequence --scans 200 --no-cpu-baseline"
run 32 X=1 "--workload sequence --scans 200 --no-cpu-baseline"
run 32 MLO_B200_LIB=$PWD/scratch/libmlo_b200_prev.so "--workload sequence --scans 200 --no-cpu-baseline"
run 1 X=1 "--workload sequence --scans 300 --no-cpu-baseline"
run 1 MLO_B200_LIB=$PWD/scratch/libmlo_b200_prev.so "--workload sequence --scans 300 --no-cpu-baseline"
} > $O/r2C_ab.log 2>&1
cut -c1-330 $O/r2C_ab.log
