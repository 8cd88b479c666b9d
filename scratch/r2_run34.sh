#!/bin/bash
# round 2, GPU call 34: ncu --set full of one queue-driven kernel launch (32-sequence fleet, default pipeline)
cd "$(dirname "$0")/.."
O=gpurun_out
export MLO_BENCH_CUPROF=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_icp_persistent -s 10 -c 1 \
    -o $O/r2H_prof_persistent_S32 -f python bench.py --workload sequence --sequences 32 --scans 12 --no-cpu-baseline > $O/r2H_ncu.log 2>&1
tail -2 $O/r2H_ncu.log; ls -la $O/r2H_prof_persistent_S32.ncu-rep
