#!/bin/bash
# last run of the round: the whole GPU suite on the final tree + the cold upload with the packers' non-temporal copy off / on
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_final.log
for nt in 0 1; do
  E=""; [ $nt = 1 ] && E="BMB200_PACK_NT=1"
  env $E BMB200_TRACE=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu --no-parity --no-e2e-check --no-e2e-slab --e2e-steps 3 2> gpurun_out/final_nt$nt.err | tail -1 > gpurun_out/final_nt$nt.json
  python -c "
import json; d=json.load(open('gpurun_out/final_nt$nt.json')); e=d['e2e']; print('NT=$nt cold', round(e['cold']['ms_per_step'],1), e['cold']['split_ms'], 'warm', round(e['ms_per_step'],3), 'device', round(d['ms_per_step'],4))"
  grep -a "issuing thread waited" gpurun_out/final_nt$nt.err | tail -2
done
