#!/bin/bash
mkdir -p gpurun_out
oracle/_ref/test_cxx_binding 2>&1 | tail -6 | tee gpurun_out/cxx_binding.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c3 ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'bits', d['result_bits'])"
python scripts/bench_c4.py 2>&1 | tail -1 > gpurun_out/bench_c4.json
python -c "
import json
d=json.load(open('gpurun_out/bench_c4.json'))
for k,v in d['results'].items(): print(k, {a:round(b,3) for a,b in v['gpu'].items()}, v.get('parity'))"
