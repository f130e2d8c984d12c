#!/bin/bash
# BLOB phase timings + the whole GPU suite + smoke in one call
bash scripts/gpu_blob_trace.sh 2>&1 | grep -v "token table\|arena layout\|blob_decode_kernel\|free temporaries\|walk buffers\|host walk\|host token"
timeout -s KILL 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_full.log 2>&1; echo "suite rc=$?"; tail -4 gpurun_out/pytest_gpu_full.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
