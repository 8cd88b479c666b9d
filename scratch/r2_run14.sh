#!/bin/bash
# round 2, GPU call 14: block-per-cloud decimation (k_decim_cta) tests + A/B; warp-partial variants of the work-list kernel
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_filter_kernels.py -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_paths.py -x -q -k "drain_variants" 2>&1 | tail -3
benchq() { # env
  echo "== $1"
  env $1 timeout 600 python bench.py --steps 10 --no-cpu-baseline --sub-records none 2>$O/r2n_last.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],3),'launch_us',round(d['roofline']['avg_launch_us'],1),'launches',d['gpu_launches'])" || tail -5 $O/r2n_last.err
}
{
benchq "MLO_FILTER_KERNEL=1"
benchq "MLO_FILTER_KERNEL=2"
benchq "MLO_FILTER_KERNEL=0"
benchq "MLO_WL_VARIANT=10"
benchq "MLO_WL_VARIANT=11"
benchq "MLO_WL_WARPS=1 MLO_WL_MIN_BLOCKS=24"
benchq "MLO_WL_WARPS=1 MLO_WL_MIN_BLOCKS=32"
} > $O/r2n_ab.log 2>&1
cat $O/r2n_ab.log
run() { # S env args
  echo "== S=$1 $2 $3"
  env $2 timeout 600 python bench.py --sequences $1 $3 2> $O/r2n_last.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'scans/s', {k:round(v,3) for k,v in d['phases']['host_wall_timed_pass'].items()})" || tail -5 $O/r2n_last.err
}
{
run 32 MLO_FILTER_KERNEL=1 "--workload sequence --scans 100 --no-cpu-baseline"
run 32 MLO_FILTER_KERNEL=2 "--workload sequence --scans 100 --no-cpu-baseline"
run 8 MLO_FILTER_KERNEL=1 "--workload sequence --scans 100 --no-cpu-baseline"
run 8 MLO_FILTER_KERNEL=2 "--workload sequence --scans 100 --no-cpu-baseline"
run 1 MLO_FILTER_KERNEL=1 "--workload sequence --scans 200 --no-cpu-baseline"
run 1 MLO_FILTER_KERNEL=2 "--workload sequence --scans 200 --no-cpu-baseline"
run 128 MLO_FILTER_KERNEL=1 "--workload sequence --scans 40 --no-cpu-baseline"
run 128 MLO_FILTER_KERNEL=2 "--workload sequence --scans 40 --no-cpu-baseline"
} > $O/r2n_seq.log 2>&1
cat $O/r2n_seq.log | cut -c1-400
