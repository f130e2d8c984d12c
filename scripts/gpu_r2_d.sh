#!/bin/bash
# round 2, run D: single out-of-line copy of the FLAT sweep (instruction-cache footprint), variants, correctness subset
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "golden or random_mixed or flat or gap_stream or large_groups or edge or c3_quarter or pipeline" 2>&1 | tail -4 | tee gpurun_out/pytest_d.log
for rep in 1 2; do
for w in c3 c5 c2; do
  timeout 300 python bench.py --workload $w --steps 10 --no-e2e --no-cpu --no-parity 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('default $w', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['result_bits'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
done
for v in oneslot unroll2; do
  for w in c3 c5; do
    BMB200_LIB=$PWD/scripts/_bin/libbmb200_$v.so timeout 300 python bench.py --workload $w --steps 10 --no-e2e --no-cpu --no-parity 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v $w', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['result_bits'])"
  done
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:agg_kernel -s 2 -c 1 -f -o gpurun_out/prof_agg_c5 python bench.py --workload c5 --steps 1 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/ncu_c5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:agg_kernel -s 2 -c 1 -f -o gpurun_out/prof_agg_c3 python bench.py --workload c3 --steps 1 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/ncu_c3.log 2>&1
ls -la gpurun_out/*.ncu-rep
