#!/bin/bash
# round 2, GPU call 21: where to hand the tail of a large batch over to the queue-driven kernel
cd "$(dirname "$0")/.."
O=gpurun_out
benchq() { # env
  echo "== $1"
  env $1 timeout 600 python bench.py --steps 10 --no-cpu-baseline --sub-records none 2>$O/r2u_last.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],3),'launch_us',round(d['roofline']['avg_launch_us'],1),'launches',d['gpu_launches'])" || tail -5 $O/r2u_last.err
}
{
benchq "MLO_TAIL_QUERIES_PER_SM=512"
benchq "MLO_TAIL_QUERIES_PER_SM=1024"
benchq "MLO_TAIL_QUERIES_PER_SM=2048"
benchq "MLO_TAIL_QUERIES_PER_SM=4096"
benchq "MLO_TAIL_QUERIES_PER_SM=1024 MLO_CHECK_EVERY=2"
benchq "MLO_TAIL_QUERIES_PER_SM=2048 MLO_CHECK_EVERY=2"
benchq "MLO_TAIL_QUERIES_PER_SM=512 MLO_CHECK_EVERY=8"
} > $O/r2u_tail_ab.log 2>&1
cat $O/r2u_tail_ab.log
