#!/bin/bash
run() { echo "== $1"; env $1 timeout 200 python bench.py --workload sequence --sequences $2 --scans 100 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['phases']['host_wall_timed_pass']['align_ms_per_step'],3))"; }
run "X=1" 1
run "MLO_LOCALMAP_CAPACITY_VOXELS=262144" 1
run "MLO_TABLE_FACTOR=2" 1
run "MLO_LOCALMAP_CAPACITY_VOXELS=262144" 32
run "MLO_TABLE_FACTOR=2" 32
