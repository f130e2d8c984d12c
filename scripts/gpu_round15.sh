#!/bin/bash
mkdir -p gpurun_out
oracle/_ref/test_cxx_binding 2>&1 | tail -3 | tee gpurun_out/cxx_binding.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 600 python scripts/bench_scan.py 2>&1 | tail -1 | tee gpurun_out/bench_scan.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for w,r in d['results'].items():
    print(w, r['kinds'])
    for k in ('find_eq','find_gt','find_range'):
        x=r[k]; print('  ',k, round(x['gpu_ms_per_launch'],3),'ms/64', round(x['gpu_plane_GBps']),'GB/s', 'ref ms/search', round(x['ref_ms_per_search_1core'],3), 'x', round(x['speedup_vs_1core']), x['counts_equal_first16'])"
