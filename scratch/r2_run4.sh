#!/bin/bash
# round 2, GPU call 4: map growth (small tables), decimation PPT, trace again
cd "$(dirname "$0")/.."
O=gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -x -q) > $O/r2d_gpu_tests.log 2>&1; tail -5 $O/r2d_gpu_tests.log
rm -f $O/r2d_trace_block.log
for CLU in 1 8; do
  echo "== trace S=1 cluster=$CLU (initial map capacity 2^17, grows on demand)" >> $O/r2d_trace_block.log
  MLO_BLOCK_CLUSTER=$CLU MLO_B200_LIB=$PWD/scratch/libmlo_b200_trace.so timeout 300 python scratch/trace_block.py 1 >> $O/r2d_trace_block.log 2>&1
done
cat $O/r2d_trace_block.log
run() { # S env args
  echo "== S=$1 $2 $3"
  env $2 timeout 400 python bench.py --sequences $1 --no-cpu-baseline $3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'scans/s', {k:round(v,3) for k,v in d['phases']['host_wall_timed_pass'].items()}, {k:round(v,3) for k,v in d['phases']['device_events_pass'].items()})"
}
{
run 1 MLO_ALIGN_PATH=2 "--workload sequence --scans 120"
run 1 MLO_ALIGN_PATH=3 "--workload sequence --scans 120"
run 1 "MLO_ALIGN_PATH=3 MLO_ICP_PRIOR=0" "--workload sequence --scans 120"
run 32 MLO_ALIGN_PATH=2 "--workload sequence --scans 60"
run 32 MLO_ALIGN_PATH=3 "--workload sequence --scans 60"
run 32 "MLO_ALIGN_PATH=3 MLO_FILTER_PPT=1" "--workload sequence --scans 60"
run 64 MLO_ALIGN_PATH=3 "--workload sequence --scans 60"
run 128 MLO_ALIGN_PATH=3 "--workload sequence --scans 40"
run 8 MLO_ALIGN_PATH=2 "--workload ndt --scans 40"
run 8 MLO_ALIGN_PATH=3 "--workload ndt --scans 40"
} > $O/r2d_block_ab.log 2>&1
cat $O/r2d_block_ab.log
benchq() { # env
  echo "== $1"
  env $1 timeout 600 python bench.py --steps 10 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],3),'launch_us',round(d['roofline']['avg_launch_us'],1),'launches',d['gpu_launches'])"
}
{
benchq "MLO_FILTER_PPT=4"
benchq "MLO_FILTER_PPT=2"
benchq "MLO_FILTER_PPT=1"
} > $O/r2d_bench_ab.log 2>&1
cat $O/r2d_bench_ab.log
MLO_BENCH_CUPROF=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3000 --csv \
    --log-file $O/r2d_launches_default_B512.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/r2d_ncu1.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2d_launches_default_B512.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    name=r[4].split('(')[0][:60]; v=float(r[-1].replace(',',''));  agg[name][0]+=1; agg[name][1]+=v
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{k:62s} n={v[0]:5d} total_us={v[1]/1e3:10.1f} share={v[1]/tot:6.3f}")
PY
