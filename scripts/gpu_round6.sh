#!/bin/bash
mkdir -p gpurun_out
oracle/_ref/test_cxx_binding 2>&1 | tail -3 | tee gpurun_out/cxx_binding.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 300 python bench.py --workload c3 --steps 10 --no-e2e --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_c3_pad.log
timeout 300 python bench.py --workload c2 --steps 10 --no-e2e --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_c2_pad.log
python scripts/bench_c4.py 1000000 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for k,v in d['results'].items(): print(k, v['gpu'], v.get('parity'))"
