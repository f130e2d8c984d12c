#!/bin/bash
# 2 GPUs: peer-memory exchange (default) vs ncclAllGather, device-timed C3; C++ sharded aggregator in both modes
mkdir -p gpurun_out
P=29617
for mode in direct nccl; do
  E=""; [ $mode = nccl ] && E="BMB200_EXCHANGE_NCCL=1"
  for rep in 1 2; do
    env $E timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --no-cpu --no-parity 2>gpurun_out/n2c_$mode.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$mode', d['ms_per_step'], d['roofline']['kernel_ms'], d['exchange']['mode'], d['exchange']['global_result_bits'])" | cut -c1-400
    P=$((P+1))
  done
done 2>&1 | tee gpurun_out/exchange_modes_n2.txt
tail -5 gpurun_out/n2c_direct.err
timeout 300 bash scripts/gpu_r2_n2b.sh 2>&1 | head -6
BMB200_EXCHANGE_NCCL=1 timeout 300 bash scripts/gpu_r2_n2b.sh 2>&1 | head -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+5)) bench.py --gpus 2 --workload c5 --no-e2e --steps 10 2> gpurun_out/bench_c5_n2.err | tail -1 | tee gpurun_out/bench_c5_n2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c5 n2', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['exchange']['mode'], d['parity']['equal'])"
