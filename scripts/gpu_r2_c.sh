#!/bin/bash
# round 2, run C: aligned live mask (lean decode fixed), rs index v2, full GPU suite, C++ binding output, variants, ncu
mkdir -p gpurun_out
oracle/_ref/test_cxx_binding 2>&1 | tail -15 | tee gpurun_out/cxx_binding.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
for v in old mode1 mode1_1slot mode2_1slot; do
  for w in c3 c5; do
    BMB200_LIB=$PWD/scripts/_bin/libbmb200_$v.so timeout 300 python bench.py --workload $w --steps 10 --no-e2e --no-cpu --no-parity 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v $w', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['result_bits'])"
  done
done
for w in c3 c5 c2; do
  timeout 300 python bench.py --workload $w --steps 10 --no-e2e --no-cpu --no-parity 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('default $w', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['result_bits'])"
done
BMB200_TRACE=1 timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
grep -a "set_upload_vectors" gpurun_out/bench_c3.err | tail -8
tail -c 600 gpurun_out/bench_c3.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_c3.json'))
    print('C3 ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'parity', d['parity']['equal'], d['parity']['host_synth_s'], d['parity']['ref_build_s'], d['parity']['ref_ms'])
    print('e2e', json.dumps(d['e2e'])[:1800])
except Exception as e: print('bench json', e)
PY
timeout 600 python scripts/bench_c4.py 2>&1 | tail -1 > gpurun_out/bench_c4.json
python -c "
import json
d=json.load(open('gpurun_out/bench_c4.json'))
for k,v in d['results'].items(): print(k, {a:round(b,3) for a,b in v['gpu'].items()}, v.get('parity'))"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:agg_kernel -s 2 -c 1 -f -o gpurun_out/prof_agg_c5 python bench.py --workload c5 --steps 1 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/ncu_c5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:agg_kernel -s 2 -c 1 -f -o gpurun_out/prof_agg_c3 python bench.py --workload c3 --steps 1 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/ncu_c3.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:rs_ -c 12 -f -o gpurun_out/prof_rs python scripts/bench_c4.py 200000 > gpurun_out/ncu_rs.log 2>&1
ls -la gpurun_out/*.ncu-rep
