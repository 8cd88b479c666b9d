#!/bin/bash
# round 2, GPU call 17: where the solve step's microseconds go (fine-grained %globaltimer events); k_decim_cta after the job staging
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_filter_kernels.py -x -q 2>&1 | tail -3
echo "== queue-driven kernel timeline S=1 (solve events)" > $O/r2q_trace_solve.log
MLO_B200_LIB=$PWD/scratch/libmlo_b200_trace.so timeout 300 python scratch/trace_persistent.py 1 >> $O/r2q_trace_solve.log 2>&1
cat $O/r2q_trace_solve.log | tail -30
export MLO_BENCH_CUPROF=1
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --profile-from-start off -k regex:k_decim_cta -c 2 --csv \
    --log-file $O/r2q_decim_cta.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --sub-records none > $O/r2q_ncu1.log 2>&1
unset MLO_BENCH_CUPROF
grep "k_decim_cta" $O/r2q_decim_cta.csv | cut -d, -f5,12- | cut -c1-200
