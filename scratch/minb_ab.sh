#!/bin/bash
run() { # S extra-env extra-args
  echo "== S=$1 $2 $3"
  env $2 timeout 300 python bench.py --sequences $1 --no-cpu-baseline $3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'scans/s', {k:round(v,3) for k,v in d['phases']['host_wall_timed_pass'].items() if 'align' in k})"
}
for m in 4 2; do
run 1 MLO_PERS_MINB=$m "--workload sequence --scans 100"
run 8 MLO_PERS_MINB=$m "--workload sequence --scans 60"
run 32 MLO_PERS_MINB=$m "--workload sequence --scans 60"
done
