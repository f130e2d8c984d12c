#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
summ='import json,sys
d=json.loads(sys.stdin.read()); print(sys.argv[1], "ms", round(d["ms_per_step"],3), "frac", round(d["roofline"]["frac"],3), "bits", d.get("result_bits"), "traffic", d["roofline"].get("traffic"))'
timeout 300 python bench.py --no-e2e --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_c3_quick.json | python -c "$summ" c3_flat
BMB200_GAP_LEGACY=1 timeout 300 python bench.py --steps 10 --no-e2e --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_c3_raw.json | python -c "$summ" c3_raw
BMB200_GAP_LEGACY=1 timeout 300 python bench.py --steps 5 --no-e2e --no-cpu --workload c5 2>&1 | tail -1 | tee gpurun_out/bench_c5_raw.json | python -c "$summ" c5_raw
