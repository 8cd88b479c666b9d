#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
export MLO_BENCH_CUPROF=1
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --profile-from-start off -k regex:k_decim_cta -c 2 --csv \
    --log-file $O/r2r_decim_cta.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --sub-records none > $O/r2r_ncu1.log 2>&1
grep "k_decim_cta" $O/r2r_decim_cta.csv | cut -d, -f5,12- | cut -c1-200
