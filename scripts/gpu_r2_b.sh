#!/bin/bash
# round 2, run B: upload pipeline after the condvar fix, kernel variants (lean pair decode / 4 KB slot) on C3 + C5, ncu of the C5 sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "upload_vectors or e2e_harness or leaks or cxx_binding" 2>&1 | tail -5 | tee gpurun_out/pytest_b.log
BMB200_TRACE=1 timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
grep -a "set_upload_vectors" gpurun_out/bench_c3.err | tail -12
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c3.json'))
print('C3 ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'parity', d['parity']['equal'], d['parity']['host_synth_s'], d['parity']['ref_build_s'], d['parity']['ref_ms'])
print('e2e', json.dumps(d['e2e'])[:1500])
PY
for v in old lean1slot; do
  for w in c3 c5; do
    BMB200_LIB=$PWD/scripts/_bin/libbmb200_$v.so timeout 300 python bench.py --workload $w --steps 10 --no-e2e --no-cpu --no-parity 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v $w', round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"
  done
done
for w in c3 c5 c2; do
  timeout 300 python bench.py --workload $w --steps 10 --no-e2e --no-cpu --no-parity 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('lean(default) $w', round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:agg_kernel -s 2 -c 1 -f -o gpurun_out/prof_agg_c5 python bench.py --workload c5 --steps 1 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/ncu_c5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:agg_kernel -s 2 -c 1 -f -o gpurun_out/prof_agg_c3 python bench.py --workload c3 --steps 1 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/ncu_c3.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 900 python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cut -c1-2500 gpurun_out/bench_ref.json
