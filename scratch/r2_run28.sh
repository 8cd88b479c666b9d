#!/bin/bash
# round 2, GPU call 28 (2 GPUs): the driver's own N=2 command lines, both arms
cd "$(dirname "$0")/.."
O=gpurun_out
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5) > $O/r2B_bench_n2.json 2> $O/r2B_bench_n2.err
tail -4 $O/r2B_bench_n2.err
python - <<'PY'
import json
for l in open('gpurun_out/r2B_bench_n2.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('N=2 value',round(d['value']),'e2e',round(d['e2e']['value']),'subs',{k:round(v['value']) for k,v in d.get('sub_records',{}).items()}, 'clocks', d['clocks'])
PY
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 4 --warmup 3) > $O/r2B_ref_n2.json 2> $O/r2B_ref_n2.err
grep -c "^{" $O/r2B_ref_n2.json; cut -c1-160 $O/r2B_ref_n2.json | head -3; tail -3 $O/r2B_ref_n2.err
