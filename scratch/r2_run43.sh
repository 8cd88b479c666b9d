#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
run() { # S env args
  echo "== S=$1 $2 $3"
  env $2 timeout 300 python bench.py --sequences $1 $3 2> $O/r2Q_last.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'scans/s', {k:round(v,3) for k,v in d['phases']['host_wall_timed_pass'].items()})" || tail -5 $O/r2Q_last.err
}
{
run 128 X=1 "--workload sequence --scans 40 --no-cpu-baseline"
run 128 X=2 "--workload sequence --scans 40 --no-cpu-baseline"
run 128 X=3 "--workload sequence --scans 60 --no-cpu-baseline"
run 64 X=1 "--workload sequence --scans 60 --no-cpu-baseline"
run 32 X=1 "--workload sequence --scans 200 --no-cpu-baseline"
} > $O/r2Q_seq.log 2>&1
cut -c1-330 $O/r2Q_seq.log
timeout 600 python -m pytest tests/test_gpu_paths.py tests/test_host_layer.py -m gpu -q -x 2>&1 | tail -2
