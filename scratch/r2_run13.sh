#!/bin/bash
# round 2, GPU call 13: contiguous-range drain of the work-list kernel (variants 6-9) vs the default (3)
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_paths.py -x -q -k "drain_variants or large_batch_launch" 2>&1 | tail -5
benchq() { # env
  echo "== $1"
  env $1 timeout 600 python bench.py --steps 10 --no-cpu-baseline --sub-records none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],3),'launch_us',round(d['roofline']['avg_launch_us'],1),'launches',d['gpu_launches'])"
}
{
for v in 3 6 7 8 9 3 6; do benchq "MLO_WL_VARIANT=$v"; done
} > $O/r2m_wl_oct_ab.log 2>&1
cat $O/r2m_wl_oct_ab.log
(time timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_paths.py -x -q -k "every_align_path") > $O/r2m_sanitizer_racecheck.log 2>&1; tail -4 $O/r2m_sanitizer_racecheck.log
