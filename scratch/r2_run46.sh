#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
run() { # S env args
  echo "== S=$1 $2 $3"
  env $2 timeout 200 python bench.py --sequences $1 $3 2> $O/r2S_last.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'scans/s', {k:round(v,3) for k,v in d['phases']['host_wall_timed_pass'].items()})" || tail -5 $O/r2S_last.err
}
{
run 128 X=1 "--workload sequence --scans 60 --no-cpu-baseline"
run 128 X=1 "--workload sequence --scans 60 --no-cpu-baseline --no-prefetch"
run 128 X=2 "--workload sequence --scans 60 --no-cpu-baseline"
run 128 X=2 "--workload sequence --scans 60 --no-cpu-baseline --no-prefetch"
} > $O/r2S_s128.log 2>&1
cut -c1-330 $O/r2S_s128.log
