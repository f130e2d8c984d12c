#!/bin/bash
mkdir -p gpurun_out
P=29717
for mode in direct direct_defcarve nccl; do
  E=""; [ $mode = nccl ] && E="BMB200_EXCHANGE_NCCL=1"; [ $mode = direct_defcarve ] && E="BMB200_XCHG_DEFAULT_CARVEOUT=1"
  for rep in 1 2; do
    env $E timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --no-cpu --no-parity 2>gpurun_out/n2d_$mode.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$mode', d['ms_per_step'], d['roofline']['kernel_ms'], d['exchange']['mode'], d['exchange']['global_result_bits'])" | cut -c1-400
    P=$((P+1))
  done
done 2>&1 | tee gpurun_out/exchange_modes_n2.txt
