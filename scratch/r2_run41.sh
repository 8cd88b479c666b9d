#!/bin/bash
# round 2, last GPU call: smoke() + the default bench line (short) + the reference arm on the final tree
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
(time timeout 900 python bench.py --cpu-budget 8 --sub-scans 300 --cpu-scans 300) > $O/r2O_bench.json 2> $O/r2O_bench.err; tail -3 $O/r2O_bench.err; python -c "
import json
d=json.load(open('gpurun_out/r2O_bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'cpu',d['cpu_baseline']['value'],'frac',d['roofline']['frac'], 'parity', d['quality']['parity_vs_oracle']['resident'])
for k,v in d['sub_records'].items(): print(k, round(v['value'],1), (v.get('cpu_baseline') or {}).get('value'), v.get('speedup_vs_cpu'))
"
