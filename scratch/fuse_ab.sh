#!/bin/bash
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run() { # S extra-env extra-args
  echo "== S=$1 $2 $3"
  env $2 timeout 300 python bench.py --sequences $1 --no-cpu-baseline $3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'scans/s', {k:round(v,3) for k,v in d['phases']['host_wall_timed_pass'].items() if 'align' in k or 'filter' in k}, d['quality'])"
}
for f in 1 0; do
run 1 MLO_FUSE_INNER=$f "--workload sequence --scans 120"
run 32 MLO_FUSE_INNER=$f "--workload sequence --scans 60"
echo "== default MLO_FUSE_INNER=$f"
MLO_FUSE_INNER=$f timeout 300 python bench.py --steps 10 --cpu-budget 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value'],1),'scans/s e2e',round(d['e2e']['value'],1),'frac',round(r['frac'],4),'avg_us',round(r['avg_launch_us'],1),'share',round(r['kernel_share_of_step'],3), d['quality']['parity_vs_oracle'], 'launches', d['gpu_launches'])"
done
