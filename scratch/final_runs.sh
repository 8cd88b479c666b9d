#!/bin/bash
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
(time timeout 400 python bench.py) > gpurun_out/final_default.json 2> gpurun_out/final_default.err; cat gpurun_out/final_default.json
(time timeout 600 python bench.py --workload sequence --sequences 32 --scans 120) > gpurun_out/final_seq32.json 2> gpurun_out/final_seq32.err; cat gpurun_out/final_seq32.json
(time timeout 600 python bench.py --workload sequence --sequences 1 --scans 300) > gpurun_out/final_seq1.json 2> gpurun_out/final_seq1.err; cat gpurun_out/final_seq1.json
(time timeout 600 python bench.py --workload ndt --sequences 8 --scans 60) > gpurun_out/final_ndt8.json 2> gpurun_out/final_ndt8.err; cat gpurun_out/final_ndt8.json
MLO_BENCH_CUPROF=1 MLO_STREAM_GROUPS=1 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_match_accumulate_wl4 -s 0 -c 1 \
    -o gpurun_out/prof_wl4_B512_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b5.log 2>&1
ls -la gpurun_out/prof_wl4_B512_full.ncu-rep
