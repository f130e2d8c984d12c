#!/bin/bash
mkdir -p gpurun_out
oracle/_ref/test_cxx_binding 2>&1 | tail -5 | tee gpurun_out/cxx_binding.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
# launch list of the default bench command shape (cold-cache, serialised: compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c3.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:agg_kernel -s 2 -c 1 -o gpurun_out/prof_agg_c3_v5 \
   python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:agg_kernel -s 2 -c 1 -o gpurun_out/prof_agg_c2_v5 \
   python bench.py --workload c2 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_run_c2.log 2>&1
