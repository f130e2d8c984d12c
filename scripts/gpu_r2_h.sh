#!/bin/bash
# round 2, run H: the warm call's D2H is 3x slower back to back than after an upload -- pinned lines parked in the private L2s of the store threads?
mkdir -p gpurun_out
for thr in 1 2 4 8; do
  echo "== BMB200_STORE_THREADS=$thr"
  BMB200_STORE_THREADS=$thr BMB200_TRACE=1 timeout 900 python bench.py --gpus 1 --steps 12 --warmup 3 --no-cpu --no-parity --no-e2e-check --no-e2e-slab --e2e-steps 1 > gpurun_out/bench_h_$thr.json 2> gpurun_out/bench_h_$thr.err
  grep -a "aggregator::run\|result_fetch_view" gpurun_out/bench_h_$thr.err | tail -4
  python -c "
import json; d=json.load(open('gpurun_out/bench_h_$thr.json')); print('warm ms', d['e2e']['ms_per_step'], 'device', d['ms_per_step'])"
done
