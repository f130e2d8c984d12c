#!/bin/bash
# round 2, run A: parity suite + the new bench line (all-column parity, e2e on real bvectors, reference arm)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,memory.total --format=csv > gpurun_out/box.txt
nproc >> gpurun_out/box.txt; free -g | head -2 >> gpurun_out/box.txt; lscpu | grep -i "model name\|numa\|socket" >> gpurun_out/box.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
BMB200_TRACE=1 timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
tail -c 3000 gpurun_out/bench_c3.err
cut -c1-6000 gpurun_out/bench_c3.json
timeout 600 python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -c 1000 gpurun_out/bench_ref.err
cut -c1-3000 gpurun_out/bench_ref.json
