#!/bin/bash
# 1 GPU: the exchange machinery with a 1-rank communicator -- where does the per-step gap of the multi-GPU runs come from?
mkdir -p gpurun_out
for mode in plain self_nccl self_direct; do
  E=""; [ $mode = self_nccl ] && E="BENCH_SELF_EXCHANGE=1"; [ $mode = self_direct ] && E="BENCH_SELF_EXCHANGE=1 BMB200_EXCHANGE_DIRECT=1"
  for rep in 1 2; do
    env $E timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-parity 2>gpurun_out/j_$mode.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$mode', d['ms_per_step'], d['roofline']['kernel_ms'])"
  done
done 2>&1 | tee gpurun_out/exchange_self_n1.txt
tail -3 gpurun_out/j_self_nccl.err
