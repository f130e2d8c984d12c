#!/bin/bash
# live-mask + FLAT consumer: parity, then C3 / C2 / C5 with the flat form and with the raw GAP form
mkdir -p gpurun_out
oracle/_ref/test_cxx_binding 2>&1 | tail -4 | tee gpurun_out/cxx_binding.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
summ='import json,sys
d=json.loads(sys.stdin.read()); print(sys.argv[1], "ms", round(d["ms_per_step"],3), "frac", round(d["roofline"]["frac"],3), "bits", d.get("result_bits"))'
timeout 300 python bench.py --steps 10 --no-e2e --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_c3_flat.json | python -c "$summ" c3_flat
BMB200_GAP_LEGACY=1 timeout 300 python bench.py --steps 10 --no-e2e --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_c3_raw.json | python -c "$summ" c3_raw
timeout 300 python bench.py --steps 10 --no-e2e --no-cpu --workload c2 2>&1 | tail -1 | tee gpurun_out/bench_c2.json | python -c "$summ" c2
timeout 300 python bench.py --steps 5 --no-e2e --no-cpu --workload c5 2>&1 | tail -1 | tee gpurun_out/bench_c5.json | python -c "$summ" c5_flat
BMB200_GAP_LEGACY=1 timeout 300 python bench.py --steps 5 --no-e2e --no-cpu --workload c5 2>&1 | tail -1 | tee gpurun_out/bench_c5_raw.json | python -c "$summ" c5_raw
