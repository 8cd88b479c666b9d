#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
echo "== default" > $O/r2i_diag.log
timeout 900 python scratch/diag_fleet_parity.py 16 100 >> $O/r2i_diag.log 2>&1
echo "== MLO_HOST_THREADS=1" >> $O/r2i_diag.log
MLO_HOST_THREADS=1 timeout 900 python scratch/diag_fleet_parity.py 16 100 >> $O/r2i_diag.log 2>&1
echo "== MLO_ICP_PRIOR=0" >> $O/r2i_diag.log
MLO_ICP_PRIOR=0 timeout 900 python scratch/diag_fleet_parity.py 16 100 >> $O/r2i_diag.log 2>&1
cat $O/r2i_diag.log | cut -c1-400
