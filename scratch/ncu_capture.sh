#!/bin/bash
# ncu evidence for profiles/: launch lists of the timed regions + one full capture of the dominant kernel.
set -x
export MLO_BENCH_CUPROF=1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1200 --csv \
    --log-file gpurun_out/launches_default_B512.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3000 --csv \
    --log-file gpurun_out/launches_fleet_S32.csv python bench.py --workload sequence --sequences 32 --scans 12 --no-cpu-baseline > gpurun_out/ncu_b2.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_match_accumulate_wl4 -s 2 -c 2 \
    -o gpurun_out/prof_wl4_B512 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b3.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_match_accumulate_wl4 -s 4 -c 2 \
    -o gpurun_out/prof_wl4_fleet_S64 -f python bench.py --workload sequence --sequences 64 --scans 8 --no-cpu-baseline > gpurun_out/ncu_b4.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_*.csv
tail -2 gpurun_out/ncu_b*.log
