#!/bin/bash
for v in 0 1 2 3 0 1; do
  echo "== MLO_WL_VARIANT=$v"
  MLO_WL_VARIANT=$v timeout 300 python bench.py --steps 10 --cpu-budget 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value'],1),'scans/s e2e',round(d['e2e']['value'],1),'frac',round(r['frac'],4),'avg_us',round(r['avg_launch_us'],1),'share',round(r['kernel_share_of_step'],3), d['quality']['parity_vs_oracle'])"
done
