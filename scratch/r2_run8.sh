#!/bin/bash
# round 2, GPU call 8: new sequence-workload machinery (GPU ray caster, streaming chunks), fleet insert fix, cp.async.bulk A/B
cd "$(dirname "$0")/.."
O=gpurun_out
run() { # S env args
  echo "== S=$1 $2 $3"
  env $2 timeout 600 python bench.py --sequences $1 $3 2> $O/r2h_last.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'scans/s cpu',d.get('cpu_baseline'),'quality',d.get('quality'), {k:round(v,3) for k,v in d['phases']['host_wall_timed_pass'].items()})" || tail -5 $O/r2h_last.err
}
{
run 32 X=1 "--workload sequence --scans 200 --cpu-scans 100"
run 128 X=1 "--workload sequence --scans 100 --no-cpu-baseline"
run 1 X=1 "--workload sequence --scans 300 --cpu-scans 100"
run 8 X=1 "--workload ndt --scans 80 --cpu-scans 40"
run 8 X=1 "--workload sequence_pt2pl --scans 80 --cpu-scans 40"
} > $O/r2h_seq.log 2>&1
cat $O/r2h_seq.log | cut -c1-900
benchq() { # env
  echo "== $1"
  env $1 timeout 600 python bench.py --steps 10 --no-cpu-baseline --sub-records none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],3),'launch_us',round(d['roofline']['avg_launch_us'],1),'launches',d['gpu_launches'])"
}
{
benchq "MLO_WL_VARIANT=3"
benchq "MLO_WL_VARIANT=4"
benchq "MLO_WL_VARIANT=5"
benchq "MLO_WL_VARIANT=0"
} > $O/r2h_wl_bulk_ab.log 2>&1
cat $O/r2h_wl_bulk_ab.log
(time timeout 900 python bench.py --sub-scans 200 --cpu-scans 100) > $O/r2h_bench_small_subs.json 2> $O/r2h_bench_small_subs.err; tail -3 $O/r2h_bench_small_subs.err; python -c "
import json
d=json.load(open('gpurun_out/r2h_bench_small_subs.json'))
print('value',d['value'],'e2e',d['e2e']['value'])
for k,v in d['sub_records'].items(): print(k, round(v['value'],1), v.get('cpu_baseline',{}).get('value'), v.get('parity_vs_oracle'), v.get('speedup_vs_cpu'))
"
