#!/bin/bash
mkdir -p gpurun_out
oracle/_ref/test_cxx_binding 2>&1 | tail -3 | tee gpurun_out/cxx_binding.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
timeout 600 python scripts/bench_blob.py 2>&1 | tail -1 | tee gpurun_out/bench_blob.json
timeout 600 python scripts/bench_scan.py 2>&1 | tail -1 > gpurun_out/bench_scan.json
timeout 300 python bench.py --no-e2e --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_c3_quick.json | cut -c1-200
