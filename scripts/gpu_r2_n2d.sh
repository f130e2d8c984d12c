#!/bin/bash
mkdir -p gpurun_out
P=30017
for steps in 20 10 40; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps $steps --warmup 5 --no-e2e --no-cpu --no-parity 2>gpurun_out/n2d_s$steps.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('steps $steps', d['ms_per_step'], d['roofline']['kernel_ms'], d['exchange']['mode'], d['exchange']['global_result_bits'])" | cut -c1-400
  P=$((P+1))
done 2>&1 | tee gpurun_out/exchange_aligned_n2.txt
