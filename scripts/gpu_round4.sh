#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 300 python bench.py --workload c3 --steps 10 --no-e2e --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_c3_stream.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:agg_kernel -s 2 -c 1 -o gpurun_out/prof_agg_c3_v2 \
   python bench.py --workload c3 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_run.log 2>&1
