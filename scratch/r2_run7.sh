#!/bin/bash
# round 2, GPU call 7: fleet throughput (host pool, launch-sequence threshold), launch list of a fleet step, ncu of the decimation claim kernel
cd "$(dirname "$0")/.."
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > $O/r2g_gpu_tests.log 2>&1; tail -6 $O/r2g_gpu_tests.log
run() { # S env args
  echo "== S=$1 $2 $3"
  env $2 timeout 400 python bench.py --sequences $1 --no-cpu-baseline $3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'scans/s', {k:round(v,3) for k,v in d['phases']['host_wall_timed_pass'].items()}, {k:round(v,3) for k,v in d['phases']['device_events_pass'].items()})"
}
{
run 32 X=1 "--workload sequence --scans 60"
run 32 MLO_HOST_THREADS=1 "--workload sequence --scans 60"
run 64 X=1 "--workload sequence --scans 60"
run 64 MLO_LARGE_BATCH_QUERIES=30000 "--workload sequence --scans 60"
run 128 X=1 "--workload sequence --scans 40"
run 128 MLO_LARGE_BATCH_QUERIES=30000 "--workload sequence --scans 40"
run 128 "MLO_LARGE_BATCH_QUERIES=30000 MLO_HOST_THREADS=1" "--workload sequence --scans 40"
run 256 MLO_LARGE_BATCH_QUERIES=30000 "--workload sequence --scans 30"
run 1 X=1 "--workload sequence --scans 120"
run 8 X=1 "--workload ndt --scans 40"
} > $O/r2g_fleet_ab.log 2>&1
cat $O/r2g_fleet_ab.log
MLO_BENCH_CUPROF=1 MLO_LARGE_BATCH_QUERIES=30000 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 6000 --csv \
    --log-file $O/r2g_launches_fleet_S128.csv python bench.py --workload sequence --sequences 128 --scans 10 --no-cpu-baseline > $O/r2g_ncu2.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2g_launches_fleet_S128.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    name=r[4].split('(')[0][:60]; v=float(r[-1].replace(',',''));  agg[name][0]+=1; agg[name][1]+=v
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{k:62s} n={v[0]:5d} total_us={v[1]/1e3:10.1f} share={v[1]/tot:6.3f}")
PY
MLO_BENCH_CUPROF=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_decim_claim -s 0 -c 1 \
   -o $O/r2g_prof_decim_claim_B512 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/r2g_ncu3.log 2>&1
ls -la $O/*.ncu-rep | tail -3
(time timeout 600 python bench.py) > $O/r2g_bench_default.json 2> $O/r2g_bench_default.err; cut -c1-600 $O/r2g_bench_default.json
