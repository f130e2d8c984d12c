#!/bin/bash
# full-set ncu captures of the kernels the main evidence run does not reach: rank / select, SHIFT-R-AND, then GAP merge, slab gather and
# the BLOB decode kernels (separate passes: -c counts matching launches in order, the shift tests alone launch 14)
mkdir -p gpurun_out
if [ "$1" != "part2" ]; then
timeout 600 ncu --set full --clock-control none -k regex:"rs_rank_kernel|rs_select_kernel" -c 6 -f -o gpurun_out/prof_rs_q python scripts/bench_c4.py 10000000 > gpurun_out/ncu_rs_q.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"shift_and_kernel" -c 14 -f -o gpurun_out/prof_rest \
   python -m pytest tests -m gpu -q -k "shift_right_and" > gpurun_out/ncu_rest.log 2>&1
ncu -i gpurun_out/prof_rs_q.ncu-rep --page raw --csv > gpurun_out/ncu_rs_q_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_rest.ncu-rep --page raw --csv > gpurun_out/ncu_rest_raw.csv 2>/dev/null
fi
timeout 600 ncu --set full --clock-control none -k regex:"gap_merge_kernel|slab_gather_kernel|blob_decode_kernel|blob_walk_kernel|blob_entropy_kernel" -c 12 -f -o gpurun_out/prof_rest2 \
   python -m pytest tests -m gpu -q -k "binop_result or upload_slabs or deserialize_to_device_vs_golden" > gpurun_out/ncu_rest2.log 2>&1
ncu -i gpurun_out/prof_rest2.ncu-rep --page raw --csv > gpurun_out/ncu_rest2_raw.csv 2>/dev/null
rm -f gpurun_out/prof_rs_q.ncu-rep gpurun_out/prof_rest.ncu-rep gpurun_out/prof_rest2.ncu-rep
tail -3 gpurun_out/ncu_rest2.log; wc -l gpurun_out/ncu_rest2_raw.csv
