#!/bin/bash
# profiling pass (run under gpurun): launch lists + one full capture of the query kernel
set -x
O=gpurun_out/r1c; mkdir -p $O
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-fit > $O/bench_under_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_fit.csv python bench_fit.py --iters 2 --reps 1 --no-graph > $O/fit_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:query_tc_kernel -s 12 -c 1 -o $O/query_tc python bench.py --steps 1 --warmup 1 --no-cpu --no-fit > $O/full_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:query_bwd_tc_kernel -s 4 -c 1 -o $O/query_bwd_tc python bench_fit.py --iters 2 --reps 1 --no-graph > $O/full_bwd_under_ncu.log 2>&1
ls -la $O
