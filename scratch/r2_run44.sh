#!/bin/bash
cd "$(dirname "$0")/.."
timeout 500 python scratch/diag_s128.py 128 50 2>&1 | tail -54 > gpurun_out/r2R_diag_s128.log; cat gpurun_out/r2R_diag_s128.log | cut -c1-170
