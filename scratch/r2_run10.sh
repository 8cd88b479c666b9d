#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > $O/r2j_gpu_tests.log 2>&1; tail -4 $O/r2j_gpu_tests.log
timeout 900 python scratch/diag_fleet_parity.py 16 100 2>&1 | grep -v "Exception ignored\|Traceback\|oracle_py.py\|TypeError\|__del__" > $O/r2j_diag.log; cat $O/r2j_diag.log | cut -c1-300
run() { # S env args
  echo "== S=$1 $2 $3"
  env $2 timeout 600 python bench.py --sequences $1 $3 2> $O/r2j_last.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'scans/s cpu',(d.get('cpu_baseline') or {}).get('value'),'parity',(d.get('quality') or {}).get('parity_vs_oracle'), {k:round(v,3) for k,v in d['phases']['host_wall_timed_pass'].items()})" || tail -5 $O/r2j_last.err
}
{
run 1 X=1 "--workload sequence --scans 120 --no-cpu-baseline"
run 1 X=1 "--workload sequence --scans 300 --no-cpu-baseline"
run 32 X=1 "--workload sequence --scans 200 --cpu-scans 200"
run 128 X=1 "--workload sequence --scans 100 --no-cpu-baseline"
} > $O/r2j_seq.log 2>&1
cat $O/r2j_seq.log | cut -c1-700
(time timeout 1500 python bench.py) > $O/r2j_bench_full.json 2> $O/r2j_bench_full.err; tail -8 $O/r2j_bench_full.err; python -c "
import json
d=json.load(open('gpurun_out/r2j_bench_full.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'cpu',d['cpu_baseline']['value'])
for k,v in d['sub_records'].items(): print(k, round(v['value'],1), (v.get('cpu_baseline') or {}).get('value'), v.get('parity_vs_oracle'), v.get('speedup_vs_cpu'), v.get('phases_ms_per_step'))
"
