#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
(time timeout 1500 python bench.py) > $O/r2l_bench_full.json 2> $O/r2l_bench_full.err; tail -6 $O/r2l_bench_full.err; python -c "
import json
d=json.load(open('gpurun_out/r2l_bench_full.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'cpu',d['cpu_baseline']['value'],'frac',d['roofline']['frac'])
for k,v in d['sub_records'].items(): print(k, round(v['value'],1), (v.get('cpu_baseline') or {}).get('value'), v.get('speedup_vs_cpu'), {a:b for a,b in (v.get('parity_vs_oracle') or {}).items() if a not in ('first_deviations','tolerance')}, v.get('ape_rmse_vs_ground_truth_m'))
"
(time timeout 900 python bench.py --impl reference) > $O/r2l_bench_reference.json 2> $O/r2l_bench_reference.err; cut -c1-400 $O/r2l_bench_reference.json; tail -3 $O/r2l_bench_reference.err
(time timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_paths.py tests/test_conventions.py -x -q -k "every_align_path or block_kernel or iteration_log or under_every_convention") > $O/r2l_sanitizer_memcheck.log 2>&1; tail -4 $O/r2l_sanitizer_memcheck.log
(time timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_paths.py -x -q -k "block_kernel_geometries or every_align_path") > $O/r2l_sanitizer_racecheck.log 2>&1; tail -4 $O/r2l_sanitizer_racecheck.log
