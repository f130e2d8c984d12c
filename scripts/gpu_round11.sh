#!/bin/bash
# ncu captures of the live-mask / FLAT kernel: C3 (and_sub), C5 shard (or), launch list of the default bench command
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:agg_kernel -s 2 -c 1 -f -o gpurun_out/prof_agg_c3_v7 \
   python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:agg_kernel -s 2 -c 1 -f -o gpurun_out/prof_agg_c5_v7 \
   python bench.py --workload c5 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_run_c5.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c3_v7.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch_run.log 2>&1
ls -la gpurun_out/
