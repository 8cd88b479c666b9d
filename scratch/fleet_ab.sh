#!/bin/bash
# A/B of the align launch policy for fleets (S sequences in lock step)
for S in 32 64; do
  for mode in "MLO_PERSISTENT=1" "MLO_PERSISTENT=2" "MLO_PERSISTENT=0"; do
    echo "== S=$S $mode"
    env $mode timeout 300 python bench.py --workload sequence --sequences $S --scans 40 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'scans/s', {k:round(v,3) for k,v in d['phases']['host_wall_timed_pass'].items()})"
  done
done
