#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
benchq() { # env
  echo "== $1"
  env $1 timeout 600 python bench.py --steps 10 --no-cpu-baseline --sub-records none 2>$O/r2F_last.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],3),'launch_us',round(d['roofline']['avg_launch_us'],1),'launches',d['gpu_launches'])" || tail -5 $O/r2F_last.err
}
{
benchq "MLO_STREAM_GROUPS=3"
benchq "MLO_STREAM_GROUPS=4"
benchq "MLO_STREAM_GROUPS=2"
benchq "MLO_STREAM_GROUPS=4 MLO_CHECK_EVERY=8"
benchq "MLO_STREAM_GROUPS=3 MLO_CHECK_EVERY=8"
benchq "MLO_STREAM_GROUPS=3 MLO_CHECK_EVERY=6"
} > $O/r2F_groups_ab.log 2>&1
cat $O/r2F_groups_ab.log
