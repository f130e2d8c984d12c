#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest "tests/test_gpu_parity.py::test_aggregate_random_mixed_vs_oracle" -x -q 2>&1 | grep -v "^$" | head -80 | tee gpurun_out/sanitizer.log
