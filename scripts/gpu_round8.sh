#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:agg_kernel -s 2 -c 1 -o gpurun_out/prof_agg_c3_v6 \
   python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:agg_kernel -s 2 -c 1 -o gpurun_out/prof_agg_c2_v6 \
   python bench.py --workload c2 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_run_c2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c3.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch_run.log 2>&1
python scripts/bench_c4.py 2>&1 | tail -1 > gpurun_out/bench_c4.json
python -c "
import json
d=json.load(open('gpurun_out/bench_c4.json'))
for k,v in d['results'].items(): print(k, v['gpu'], v.get('parity'))"
