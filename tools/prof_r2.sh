#!/bin/bash
# round-2 profiling pass (run under gpurun): launch lists, full captures of the two dominant kernels, per-launch encoder
# metrics, encoder batch sweep
O=gpurun_out/r2p; mkdir -p $O
export CHORE_B200_ENCODER_GRAPH=0      # list the encoder's kernels one by one (the product path replays them as one CUDA graph)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-fit > $O/bench_under_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_fit.csv python bench_fit.py --iters 2 --reps 1 --no-graph > $O/fit_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:query_tc_kernel -s 4 -c 1 -o $O/query_tc_r2 python bench.py --steps 1 --warmup 1 --no-cpu --no-fit > $O/full_query_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_hx_kernel -s 456 -c 8 -o $O/conv_hx_r2 python tools/time_encoder.py --batches 1 --iters 1 > $O/full_conv_under_ncu.log 2>&1
for r in query_tc_r2 conv_hx_r2; do
  ncu -i $O/$r.ncu-rep --page details > $O/${r}_details.txt 2>&1
  ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>&1
done
timeout 400 ncu --metrics gpu__time_duration.sum,launch__grid_size,sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -c 900 --csv --log-file $O/encoder_metrics.csv python tools/time_encoder.py --batches 1 --iters 1 > $O/enc_metrics_under_ncu.log 2>&1
unset CHORE_B200_ENCODER_GRAPH
timeout 200 python tools/time_encoder.py --batches 1,4,32 > $O/encoder_sweep.json 2> $O/encoder_sweep.err
ls -la $O
