#!/bin/bash
# the round's evidence run: parity, ncu captures (full set for C3 / C2 / C5 + launch list), bench lines, reference arm, C4
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,memory.total --format=csv > gpurun_out/box.txt
oracle/_ref/test_cxx_binding 2>&1 | tail -4 | tee gpurun_out/cxx_binding.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
for w in c3 c2 c5; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:agg_kernel -s 2 -c 1 -f -o gpurun_out/prof_agg_$w \
     python bench.py --workload $w --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_run_$w.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c3.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch_run.log 2>&1
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.json | cut -c1-300
timeout 900 python bench.py --gpus 1 2>&1 | tail -1 | tee gpurun_out/bench_c3_full.json | cut -c1-400
timeout 600 python bench.py --workload c2 2>&1 | tail -1 | tee gpurun_out/bench_c2_full.json | cut -c1-300
timeout 600 python bench.py --workload c5 --no-e2e --no-cpu --steps 5 2>&1 | tail -1 | tee gpurun_out/bench_c5_shard.json | cut -c1-300
python scripts/bench_c4.py 2>&1 | tail -1 > gpurun_out/bench_c4.json
python -c "
import json
d=json.load(open('gpurun_out/bench_c4.json'))
for k,v in d['results'].items(): print(k, {a:round(b,3) for a,b in v['gpu'].items()}, v.get('parity'))"
