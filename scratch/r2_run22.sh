#!/bin/bash
# round 2, GPU call 22: final state - full GPU suite, smoke(), the default bench line, the reference arm
cd "$(dirname "$0")/.."
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > $O/r2v_gpu_tests.log 2>&1; grep -n "passed\|failed" $O/r2v_gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
(time timeout 1500 python bench.py) > $O/r2v_bench_full.json 2> $O/r2v_bench_full.err; tail -4 $O/r2v_bench_full.err; python -c "
import json
d=json.load(open('gpurun_out/r2v_bench_full.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'cpu',d['cpu_baseline']['value'],'frac',d['roofline']['frac'], 'share', d['roofline']['kernel_share_of_step'], 'launches', d['gpu_launches'])
for k,v in d['sub_records'].items(): print(k, round(v['value'],1), (v.get('cpu_baseline') or {}).get('value'), v.get('speedup_vs_cpu'), {a:b for a,b in (v.get('parity_vs_oracle') or {}).items() if a not in ('first_deviations','tolerance')})
"
(time timeout 900 python bench.py --impl reference) > $O/r2v_bench_reference.json 2> $O/r2v_bench_reference.err; cut -c1-200 $O/r2v_bench_reference.json
