#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 120 oracle/_ref/test_cxx_binding > gpurun_out/cxx_binding_full.log 2>&1; tail -4 gpurun_out/cxx_binding_full.log
BMB200_TRACE=1 timeout -s KILL 200 python scripts/bench_blob.py 256 64 6 4 > gpurun_out/diag_l6_b4.json 2> gpurun_out/diag_l6_b4.err
grep "pass 2\|item\|blob_walk_kernel\|blob_entropy_kernel" gpurun_out/diag_l6_b4.err | tail -9
timeout -s KILL 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_full.log 2>&1; echo "suite rc=$?"; tail -2 gpurun_out/pytest_gpu_full.log
