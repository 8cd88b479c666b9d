#!/bin/bash
# round 2, GPU call 33: ballot drain (32-bit segment minimum, winner lane updates the 64-bit best) vs the default
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_paths.py -q -k "drain_variants" 2>&1 | tail -3
benchq() { # env
  echo "== $1"
  env $1 timeout 600 python bench.py --steps 10 --no-cpu-baseline --sub-records none 2>$O/r2G_last.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],3),'launch_us',round(d['roofline']['avg_launch_us'],1),'launches',d['gpu_launches'])" || tail -5 $O/r2G_last.err
}
{
for v in 3 18 19 3 18; do benchq "MLO_WL_VARIANT=$v"; done
} > $O/r2G_wl_runs_ab.log 2>&1
cat $O/r2G_wl_runs_ab.log
