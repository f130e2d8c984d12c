#!/bin/bash
mkdir -p gpurun_out
run() { v=$1; np=$2
  if [ $np = 1 ]; then export BMB200_NO_PAD=1; else unset BMB200_NO_PAD; fi
  BMB200_LIB=$PWD/scripts/_bin/libbmb200_$v.so timeout 300 python bench.py --workload c3 --steps 10 --no-e2e --no-cpu 2>&1 | tail -1 > gpurun_out/var_${v}_$np.log
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/var_${v}_$np.log').read().strip().splitlines()[-1])
    print('$v nopad=$np', 'ms', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'bits', d['result_bits'])
except Exception as e:
    print('$v', 'FAILED', open('gpurun_out/var_${v}_$np.log').read()[-300:])
PY
}
for rep in 1 2; do
run old 1; run new 1; run new 0; run unr2 0; run unr2 1; run unr2g32 0
done 2>&1 | tee gpurun_out/variants.txt
