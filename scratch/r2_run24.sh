#!/bin/bash
# round 2, GPU call 24: refresh the ncu --set full capture of the dominant kernel (final build) for roofline.traffic
cd "$(dirname "$0")/.."
O=gpurun_out
export MLO_BENCH_CUPROF=1 MLO_STREAM_GROUPS=1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_match_accumulate_wl4 -c 2 \
    -o $O/r2x_prof_wl4_B512 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --sub-records none > $O/r2x_ncu.log 2>&1
tail -3 $O/r2x_ncu.log; ls -la $O/r2x_prof_wl4_B512.ncu-rep
