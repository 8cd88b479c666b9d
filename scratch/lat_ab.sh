#!/bin/bash
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for f in 4 1; do echo "== trace floor $f"; MLO_QPW_FLOOR=$f MLO_B200_LIB=/root/repo/scratch/libmlo_b200_trace.so timeout 200 python scratch/trace_persistent.py 1 2>&1 | tail -10; done
run() { # S extra-env extra-args
  echo "== S=$1 $2 $3"
  env $2 timeout 300 python bench.py --sequences $1 --no-cpu-baseline $3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'scans/s', {k:round(v,3) for k,v in d['phases']['host_wall_timed_pass'].items() if 'align' in k or 'filter' in k})"
}
for f in 4 2 1; do
run 1 MLO_QPW_FLOOR=$f "--workload sequence --scans 120"
run 32 MLO_QPW_FLOOR=$f "--workload sequence --scans 60"
done
