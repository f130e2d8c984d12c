#!/bin/bash
mkdir -p gpurun_out
./scripts/_bin/microbench_gap 2>&1 | tee gpurun_out/microbench_gap.txt
# launch list (cold-cache, serialised) for the default bench command shape at reduced steps
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c3.csv \
   python bench.py --workload c3 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch_run.log 2>&1
# full capture of the dominant kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:agg_kernel -s 2 -c 1 -o gpurun_out/prof_agg_c3_v1 \
   python bench.py --workload c3 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_run.log 2>&1
ls -la gpurun_out
