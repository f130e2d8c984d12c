#!/bin/bash
# round 2, run G: slab allocator road (DMA of host slabs + device gather), bind cache + trace of the warm call
mkdir -p gpurun_out
oracle/_ref/test_cxx_binding 2>&1 | tail -3 | tee gpurun_out/cxx_binding.log
timeout 900 python -m pytest tests -m gpu -x -q -k "upload_slabs or e2e_harness or upload_vectors" 2>&1 | tail -5 | tee gpurun_out/pytest_g.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "upload_slabs" 2>&1 | tail -4 | tee gpurun_out/sanitizer_slabs.log
BMB200_TRACE=1 timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
grep -a "set_upload_slabs\|set_upload_vectors" gpurun_out/bench_c3.err | tail -14
grep -a "aggregator::run" gpurun_out/bench_c3.err | tail -6
tail -c 300 gpurun_out/bench_c3.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_c3.json'))
    print('C3 ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'parity', d['parity']['equal'], d['parity']['ref_ms'])
    print('e2e', json.dumps(d['e2e'])[:2600])
except Exception as e: print('bench json', e)
PY
