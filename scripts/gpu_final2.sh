#!/bin/bash
mkdir -p gpurun_out
for m in slot all slot all; do
  E=""; [ $m = all ] && E="BMB200_PACK_WAKE_ALL=1"
  env $E BMB200_TRACE=1 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --no-parity --no-e2e-check --no-e2e-slab --e2e-steps 3 2> gpurun_out/final2_$m.err | tail -1 > gpurun_out/final2_$m.json
  python -c "
import json; d=json.load(open('gpurun_out/final2_$m.json')); e=d['e2e']; print('wake=$m cold', round(e['cold']['ms_per_step'],1), round(e['cold']['split_ms']['device_set_assign(walk+layout+pack+H2D)'],1))"
  grep -a "issuing thread waited" gpurun_out/final2_$m.err | tail -3 | cut -c40-200
done
