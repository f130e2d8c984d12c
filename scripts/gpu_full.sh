#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.json
timeout 900 python bench.py --gpus 1 2>&1 | tail -1 | tee gpurun_out/bench_c3_full.json
timeout 600 python bench.py --workload c2 2>&1 | tail -1 | tee gpurun_out/bench_c2_full.json
