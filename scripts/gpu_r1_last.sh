#!/bin/bash
# last GPU call of the round: whole suite + smoke on the final build, BLOB bench lines with the leaner interpolative loop
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_full.log 2>&1; echo "suite rc=$?"; tail -2 gpurun_out/pytest_gpu_full.log
timeout -s KILL 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
for a in "6 0" "6 4"; do
  set -- $a
  BMB200_TRACE=1 timeout -s KILL 100 python scripts/bench_blob.py 256 64 $1 $2 > gpurun_out/bench_blob_l$1_b$2.json 2> gpurun_out/bench_blob_l$1_b$2.err
  echo "== level $1 bookmarks $2"; grep "blob_walk_kernel\|blob_entropy_kernel\|item 0" gpurun_out/bench_blob_l$1_b$2.err | tail -3; cut -c1-330 gpurun_out/bench_blob_l$1_b$2.json
done
