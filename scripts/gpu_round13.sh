#!/bin/bash
mkdir -p gpurun_out
oracle/_ref/test_cxx_binding 2>&1 | tail -6 | tee gpurun_out/cxx_binding.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
timeout 600 python scripts/bench_scan.py 2>&1 | tail -1 | tee gpurun_out/bench_scan.json | cut -c1-300
