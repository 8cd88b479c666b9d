#!/bin/bash
# round 2, GPU call 15: full GPU suite with the block-per-cloud decimation as default for large batches; ncu evidence for it;
# fresh %globaltimer timeline of the queue-driven kernel; default bench
cd "$(dirname "$0")/.."
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $O/r2o_gpu_tests.log 2>&1; tail -5 $O/r2o_gpu_tests.log
export MLO_BENCH_CUPROF=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1200 --csv \
    --log-file $O/r2o_launches_default_B512.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --sub-records none > $O/r2o_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_decim_cta -c 2 \
    -o $O/r2o_prof_decim_cta_B512 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --sub-records none > $O/r2o_ncu2.log 2>&1
unset MLO_BENCH_CUPROF
ls -la $O/*.ncu-rep | tail -3; tail -2 $O/r2o_ncu1.log $O/r2o_ncu2.log
for S in 1 32; do
  echo "== queue-driven kernel timeline S=$S"
  MLO_B200_LIB=$PWD/scratch/libmlo_b200_trace.so timeout 300 python scratch/trace_persistent.py $S 2>&1 | tail -14
done > $O/r2o_trace_persistent.log 2>&1
cat $O/r2o_trace_persistent.log
(time timeout 1500 python bench.py) > $O/r2o_bench_full.json 2> $O/r2o_bench_full.err; tail -4 $O/r2o_bench_full.err; python -c "
import json
d=json.load(open('gpurun_out/r2o_bench_full.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'cpu',d['cpu_baseline']['value'],'frac',d['roofline']['frac'], 'share', d['roofline']['kernel_share_of_step'])
for k,v in d['sub_records'].items(): print(k, round(v['value'],1), (v.get('cpu_baseline') or {}).get('value'), v.get('speedup_vs_cpu'))
"
