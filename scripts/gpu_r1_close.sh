#!/bin/bash
# round-1 closing evidence: whole GPU suite, C++ binding checks, smoke(), racecheck of the BLOB kernels, BLOB bench lines with phase traces
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_full.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/pytest_gpu_full.log
timeout -s KILL 120 oracle/_ref/test_cxx_binding > gpurun_out/cxx_binding_full.log 2>&1; tail -1 gpurun_out/cxx_binding_full.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout -s KILL 150 compute-sanitizer --tool racecheck --print-limit 5 python scripts/san_blob.py 2>&1 | grep -E "RACECHECK SUMMARY|ERROR SUMMARY|san_blob ok|hazard|Race reported" | head -6 > gpurun_out/sanitizer_blob_racecheck.log; cat gpurun_out/sanitizer_blob_racecheck.log
for a in "6 0" "6 4" "4 0" "2 0"; do
  set -- $a
  BMB200_TRACE=1 timeout -s KILL 200 python scripts/bench_blob.py 256 64 $1 $2 > gpurun_out/bench_blob_l$1_b$2.json 2> gpurun_out/bench_blob_l$1_b$2.err
  echo "== level $1 bookmarks $2"; grep "blob_walk_kernel\|blob_entropy_kernel\|blob_decode_kernel" gpurun_out/bench_blob_l$1_b$2.err | tail -3; cut -c1-420 gpurun_out/bench_blob_l$1_b$2.json
done
