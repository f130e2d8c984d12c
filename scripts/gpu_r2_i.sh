#!/bin/bash
# round 2, run I: alternating pinned fetch buffers + store overlapped with the chunked D2H; full GPU suite; store thread count
mkdir -p gpurun_out
oracle/_ref/test_cxx_binding 2>&1 | tail -3 | tee gpurun_out/cxx_binding.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
for thr in 8 16 4; do
  echo "== BMB200_STORE_THREADS=$thr"
  BMB200_STORE_THREADS=$thr BMB200_TRACE=1 timeout 900 python bench.py --gpus 1 --steps 12 --warmup 3 --no-cpu --no-parity --no-e2e-check --no-e2e-slab --e2e-steps 1 > gpurun_out/bench_i_$thr.json 2> gpurun_out/bench_i_$thr.err
  grep -a "aggregator::run\|result_fetch_view" gpurun_out/bench_i_$thr.err | tail -4
  python -c "
import json; d=json.load(open('gpurun_out/bench_i_$thr.json')); print('warm ms', d['e2e']['ms_per_step'], 'device', d['ms_per_step'])"
done
