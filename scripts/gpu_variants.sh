#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
run() { v=$1; w=$2
  BMB200_LIB=$PWD/scripts/_bin/libbmb200_$v.so timeout 300 python bench.py --workload $w --steps 10 --no-e2e --no-cpu 2>&1 | tail -1 > gpurun_out/var_${v}_$w.log
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/var_${v}_$w.log').read().strip().splitlines()[-1])
    print('$v $w', 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'bits', d['result_bits'])
except Exception as e:
    print('$v $w', 'FAILED', open('gpurun_out/var_${v}_$w.log').read()[-400:])
PY
}
for w in ${WORKLOADS:-c3 c5 c2}; do for v in $VARIANTS; do run $v $w; done; done 2>&1 | tee gpurun_out/variants.txt
