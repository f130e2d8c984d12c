#!/bin/bash
# round-1 closing run: the new rows first (deserialize-to-device for entropy-coded BLOBs, sharded rank/select), each under its
# own timeout, then the whole GPU suite, the C++ binding checks, smoke(), a short C3 bench and the BLOB bench lines
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,memory.total --format=csv > gpurun_out/box.txt
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -x -q -k "deserialize or sharded_rs" > gpurun_out/pytest_new.log 2>&1; echo "new rc=$?"; tail -15 gpurun_out/pytest_new.log
timeout -s KILL 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_full.log 2>&1; echo "suite rc=$?"; tail -8 gpurun_out/pytest_gpu_full.log
timeout -s KILL 120 oracle/_ref/test_cxx_binding > gpurun_out/cxx_binding_full.log 2>&1; tail -3 gpurun_out/cxx_binding_full.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout -s KILL 200 python scripts/bench_blob.py 256 64 6 > gpurun_out/bench_blob_l6.json 2> gpurun_out/bench_blob_l6.err; cat gpurun_out/bench_blob_l6.json; tail -3 gpurun_out/bench_blob_l6.err
timeout -s KILL 200 python scripts/bench_blob.py 256 64 4 > gpurun_out/bench_blob_l4.json 2> gpurun_out/bench_blob_l4.err; cat gpurun_out/bench_blob_l4.json
timeout -s KILL 200 python scripts/bench_blob.py 256 64 2 > gpurun_out/bench_blob_l2.json 2> gpurun_out/bench_blob_l2.err; cat gpurun_out/bench_blob_l2.json
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c3_quick.json 2> gpurun_out/bench_c3_quick.err; cut -c1-600 gpurun_out/bench_c3_quick.json
