#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > $O/r2M_gpu_tests.log 2>&1; grep -n "passed\|failed\|^FAILED" $O/r2M_gpu_tests.log
(time timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_paths.py -x -q -k "every_align_path or prior_prepared") > $O/r2M_racecheck.log 2>&1; grep -n "passed\|failed\|SUMMARY" $O/r2M_racecheck.log
