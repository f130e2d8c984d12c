#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
timeout 300 python bench.py --workload c3 --steps 10 --no-e2e --no-cpu 2>&1 | tail -3 | tee gpurun_out/bench_c3_stream.log
BMB200_GAP_MODE=1 timeout 300 python bench.py --workload c3 --steps 10 --no-e2e --no-cpu 2>&1 | tail -3 | tee gpurun_out/bench_c3_gather.log
timeout 300 python bench.py --workload c2 --steps 10 --no-e2e --no-cpu 2>&1 | tail -3 | tee gpurun_out/bench_c2.log
