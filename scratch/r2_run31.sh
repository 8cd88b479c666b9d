#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
(time timeout 1200 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_paths.py -x -q --durations=8 -k "full_size or hand_over or 2_pow_20 or round_trip or nearest_neighbour_full or 512") > $O/r2E_full_size.log 2>&1; tail -25 $O/r2E_full_size.log
