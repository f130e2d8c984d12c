#!/bin/bash
# bookmarks (segment-parallel walk) + sanitizer + ncu of the BLOB kernels + the whole suite
mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py -x -q -k "deserialize" > gpurun_out/pytest_blob.log 2>&1; echo "blob tests rc=$?"; tail -3 gpurun_out/pytest_blob.log
for tool in memcheck racecheck synccheck; do
  timeout -s KILL 150 compute-sanitizer --tool $tool --print-limit 5 python scripts/san_blob.py 2>&1 | grep -E "ERROR SUMMARY|san_blob ok|Invalid|hazard|at .*bmb200" | head -8 > gpurun_out/sanitizer_blob_$tool.log; echo "== $tool"; cat gpurun_out/sanitizer_blob_$tool.log
done
for a in "6 0" "6 4" "6 16" "4 0" "4 4"; do
  set -- $a
  BMB200_TRACE=1 timeout -s KILL 200 python scripts/bench_blob.py 256 64 $1 $2 > gpurun_out/bench_blob_l$1_b$2.json 2> gpurun_out/bench_blob_l$1_b$2.err
  echo "== level $1 bookmarks $2"; grep "blob_walk_kernel\|blob_entropy_kernel" gpurun_out/bench_blob_l$1_b$2.err | tail -2; cut -c1-330 gpurun_out/bench_blob_l$1_b$2.json
done
timeout -s KILL 240 ncu --set full --clock-control none --import-source on -k regex:blob_walk_kernel\|blob_entropy_kernel -s 2 -c 2 -f -o gpurun_out/prof_blob \
   python scripts/bench_blob.py 256 64 6 0 > gpurun_out/ncu_blob_run.log 2>&1
ncu -i gpurun_out/prof_blob.ncu-rep --page raw --csv > gpurun_out/ncu_blob_raw.csv 2>/dev/null; wc -c gpurun_out/ncu_blob_raw.csv
timeout -s KILL 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_full.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/pytest_gpu_full.log
