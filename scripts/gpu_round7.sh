#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 900 python bench.py --gpus 1 2>&1 | tail -1 | tee gpurun_out/bench_c3_full.json
timeout 900 python bench.py --workload c5 --no-e2e --cpu-cols 8 2>&1 | tail -1 | tee gpurun_out/bench_c5_shard.json
