#!/bin/bash
# one GPU trip: parity tests, then short benches; everything tee'd into gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv | tee gpurun_out/gpu.txt
nproc | tee -a gpurun_out/gpu.txt; free -g | head -2 | tee -a gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
timeout 300 python bench.py --workload c2 --steps 5 --no-e2e --no-cpu 2>&1 | tail -5 | tee gpurun_out/bench_c2_quick.log
timeout 300 python bench.py --workload c3 --cols 2048 --steps 5 --no-e2e 2>&1 | tail -5 | tee gpurun_out/bench_c3_small.log
timeout 600 python bench.py --workload c3 --steps 5 2>&1 | tail -5 | tee gpurun_out/bench_c3.log
