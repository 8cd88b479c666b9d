#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_paths.py -q -k "drain_variants" 2>&1 | tail -3
export MLO_BENCH_CUPROF=1 MLO_STREAM_GROUPS=1 MLO_WL_VARIANT=12
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_match_accumulate_wl4 -c 1 \
    -o $O/r2z_prof_wl4_a32 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --sub-records none > $O/r2z_ncu.log 2>&1
tail -2 $O/r2z_ncu.log
