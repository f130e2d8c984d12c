#!/bin/bash
# round 2, run F: final kernel (wide chunk + adaptive slots) vs the narrow-chunk variants on ONE box, e2e with the in-place result store, upload trace
mkdir -p gpurun_out
for rep in 1 2; do
for v in default narrow narrow_2slot; do
  for w in c3 c5 c2; do
    if [ $v = default ]; then L=""; else L="BMB200_LIB=$PWD/scripts/_bin/libbmb200_$v.so"; fi
    env $L timeout 300 python bench.py --workload $w --steps 10 --no-e2e --no-cpu --no-parity 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v $w', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['result_bits'])"
  done
done
done 2>&1 | tee gpurun_out/variants_f.txt
oracle/_ref/test_cxx_binding 2>&1 | tail -3 | tee gpurun_out/cxx_binding.log
BMB200_TRACE=1 timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
grep -a "set_upload_vectors" gpurun_out/bench_c3.err | tail -8
tail -c 400 gpurun_out/bench_c3.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_c3.json'))
    print('C3 ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'parity', d['parity']['equal'], d['parity']['ref_ms'])
    print('e2e', json.dumps(d['e2e'])[:1800])
except Exception as e: print('bench json', e)
PY
timeout 600 python scripts/bench_c4.py 2>&1 | tail -1 > gpurun_out/bench_c4.json
python -c "
import json
d=json.load(open('gpurun_out/bench_c4.json'))
for k,v in d['results'].items(): print(k, {a:round(b,3) for a,b in v['gpu'].items()}, v.get('parity'), v.get('fraction_of_sector_bound'))"
