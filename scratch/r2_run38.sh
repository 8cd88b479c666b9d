#!/bin/bash
# round 2, GPU call 38: final state - full GPU suite, smoke(), default bench line, 128-sequence fleet twice, racecheck on the solve paths
cd "$(dirname "$0")/.."
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > $O/r2L_gpu_tests.log 2>&1; grep -n "passed\|failed" $O/r2L_gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
(time timeout 1500 python bench.py) > $O/r2L_bench_full.json 2> $O/r2L_bench_full.err; tail -4 $O/r2L_bench_full.err; python -c "
import json
d=json.load(open('gpurun_out/r2L_bench_full.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'cpu',d['cpu_baseline']['value'],'frac',d['roofline']['frac'], 'share', d['roofline']['kernel_share_of_step'], 'launches', d['gpu_launches'])
for k,v in d['sub_records'].items(): print(k, round(v['value'],1), (v.get('cpu_baseline') or {}).get('value'), v.get('speedup_vs_cpu'), {a:b for a,b in (v.get('parity_vs_oracle') or {}).items() if a not in ('first_deviations','tolerance')})
"
for i in 1 2; do timeout 300 python bench.py --workload sequence --sequences 128 --scans 40 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('S=128', round(d['value'],1),'scans/s', {k:round(v,3) for k,v in d['phases']['host_wall_timed_pass'].items()})"; done
(time timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_paths.py -x -q -k "every_align_path or prior_prepared") > $O/r2L_racecheck.log 2>&1; grep -n "passed\|failed\|SUMMARY" $O/r2L_racecheck.log
