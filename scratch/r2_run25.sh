#!/bin/bash
# round 2, GPU call 25: split-key drain (native 32-bit shared atomics) vs the default
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_paths.py -x -q -k "drain_variants" 2>&1 | tail -3
benchq() { # env
  echo "== $1"
  env $1 timeout 600 python bench.py --steps 10 --no-cpu-baseline --sub-records none 2>$O/r2y_last.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],3),'launch_us',round(d['roofline']['avg_launch_us'],1),'launches',d['gpu_launches'])" || tail -5 $O/r2y_last.err
}
{
for v in 3 12 13 3 12; do benchq "MLO_WL_VARIANT=$v"; done
} > $O/r2y_wl_a32_ab.log 2>&1
cat $O/r2y_wl_a32_ab.log
