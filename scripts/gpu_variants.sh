#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
for v in anti noanti; do
  BMB200_LIB=$PWD/scripts/_bin/libbmb200_$v.so timeout 300 python bench.py --workload c3 --steps 10 --no-e2e --no-cpu 2>&1 | tail -1 > gpurun_out/var_$v.log
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/var_$v.log').read().strip().splitlines()[-1])
    print('$v', 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'bits', d['result_bits'])
except Exception as e:
    print('$v', 'FAILED', open('gpurun_out/var_$v.log').read()[-300:])
PY
done 2>&1 | tee gpurun_out/variants.txt
