#!/bin/bash
# last GPU call of the round: whole suite on the final build + the bookmarked level-6 BLOB line with the pass-2 item trace
mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_full.log 2>&1; echo "suite rc=$?"; tail -2 gpurun_out/pytest_gpu_full.log
BMB200_TRACE=1 timeout -s KILL 60 python scripts/bench_blob.py 256 64 6 4 > gpurun_out/bench_blob_l6_b4.json 2> gpurun_out/bench_blob_l6_b4.err
grep "blob_walk_kernel\|blob_entropy_kernel\|item 0\|pass 2" gpurun_out/bench_blob_l6_b4.err | tail -4; cut -c1-330 gpurun_out/bench_blob_l6_b4.json
