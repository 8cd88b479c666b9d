#!/bin/bash
# round 2, GPU call 2: cluster block kernel + decimation v2
cd "$(dirname "$0")/.."
O=gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -x -q) > $O/r2b_gpu_tests.log 2>&1; tail -5 $O/r2b_gpu_tests.log
rm -f $O/r2b_trace_block.log
for S in 1 32; do
  for CLU in 0 1 4; do
  echo "== trace S=$S cluster=$CLU" >> $O/r2b_trace_block.log
  MLO_BLOCK_CLUSTER=$CLU MLO_B200_LIB=$PWD/scratch/libmlo_b200_trace.so timeout 300 python scratch/trace_block.py $S >> $O/r2b_trace_block.log 2>&1
  done
done
cat $O/r2b_trace_block.log
run() { # S env args
  echo "== S=$1 $2 $3"
  env $2 timeout 400 python bench.py --sequences $1 --no-cpu-baseline $3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'scans/s', {k:round(v,3) for k,v in d['phases']['host_wall_timed_pass'].items()}, {k:round(v,3) for k,v in d['phases']['device_events_pass'].items()})"
}
{
run 1 MLO_ALIGN_PATH=2 "--workload sequence --scans 120"
run 1 MLO_ALIGN_PATH=3 "--workload sequence --scans 120"
run 1 "MLO_ALIGN_PATH=3 MLO_BLOCK_CLUSTER=4" "--workload sequence --scans 120"
run 32 MLO_ALIGN_PATH=2 "--workload sequence --scans 60"
run 32 MLO_ALIGN_PATH=3 "--workload sequence --scans 60"
run 32 "MLO_ALIGN_PATH=3 MLO_BLOCK_CLUSTER=2" "--workload sequence --scans 60"
run 64 MLO_ALIGN_PATH=3 "--workload sequence --scans 60"
run 128 MLO_ALIGN_PATH=3 "--workload sequence --scans 40"
run 8 MLO_ALIGN_PATH=2 "--workload ndt --scans 40"
run 8 MLO_ALIGN_PATH=3 "--workload ndt --scans 40"
} > $O/r2b_block_ab.log 2>&1
cat $O/r2b_block_ab.log
benchq() { # env
  echo "== $1"
  env $1 timeout 600 python bench.py --steps 10 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],3),'launch_us',round(d['roofline']['avg_launch_us'],1),'launches',d['gpu_launches'])"
}
{
benchq "MLO_FILTER_GROUP_MB=72"
benchq "MLO_FILTER_GROUP_MB=24"
benchq "MLO_FILTER_GROUP_MB=144"
benchq "MLO_FILTER_GROUP_MB=100000"
benchq "MLO_FILTER_GROUP_MB=72 MLO_TAIL_PATH=2"
} > $O/r2b_bench_ab.log 2>&1
cat $O/r2b_bench_ab.log
MLO_BENCH_CUPROF=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3000 --csv \
    --log-file $O/r2b_launches_default_B512.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/r2b_ncu1.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2b_launches_default_B512.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    name=r[4].split('(')[0][:60]; v=float(r[-1].replace(',',''));  agg[name][0]+=1; agg[name][1]+=v
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{k:62s} n={v[0]:5d} total_us={v[1]/1e3:10.1f} share={v[1]/tot:6.3f}")
PY
