#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python scratch/diag_fleet_parity.py 16 1000 2>&1 | grep -v "Exception ignored\|Traceback\|oracle_py.py\|TypeError\|__del__" > $O/r2k_diag.log; cat $O/r2k_diag.log | cut -c1-330
