#!/bin/bash
# round 2, 2-GPU run: the library's NCCL exchange (C ABI), bm::b200::sharded_aggregator, bench at N=2 (C3 with e2e, C5), C4 sharded
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n2.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -k "sharded or multi or nccl or exchange" 2>&1 | tail -5 | tee gpurun_out/pytest_n2.log
P=29517
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 20 --warmup 5 2> gpurun_out/bench_c3_n2.err | tail -1 | tee gpurun_out/bench_c3_n2.json | cut -c1-1500
tail -c 600 gpurun_out/bench_c3_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus 2 --workload c5 --no-e2e --steps 10 2> gpurun_out/bench_c5_n2.err | tail -1 | tee gpurun_out/bench_c5_n2.json | cut -c1-900
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+2)) bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -1 | tee gpurun_out/bench_ref_n2.json | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+3)) scripts/bench_c4_sharded.py 2>&1 | tail -1 | tee gpurun_out/bench_c4_n2.json | cut -c1-600
