#!/bin/bash
mkdir -p gpurun_out
BMB200_TRACE=1 timeout -s KILL 200 python scripts/bench_blob.py 256 64 6 4 > gpurun_out/diag_l6_b4.json 2> gpurun_out/diag_l6_b4.err
grep "pass 2\|item\|blob_walk_kernel\|blob_entropy_kernel" gpurun_out/diag_l6_b4.err | tail -9
