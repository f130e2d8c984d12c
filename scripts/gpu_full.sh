#!/bin/bash
# the round's evidence run on ONE GPU: parity suites, ncu captures (full set for C3 / C2 / C5 + launch list + the other kernels),
# bench lines (C3 with e2e warm / cold / slab allocator, reference arm, C2, C5 shard), C4 with sector accounting, scan
R=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,memory.total --format=csv > gpurun_out/box.txt
nproc >> gpurun_out/box.txt; lscpu | grep -i "model name\|numa\|socket" >> gpurun_out/box.txt
oracle/_ref/test_cxx_binding 2>&1 | tail -4 | tee gpurun_out/cxx_binding.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
for w in c3 c2 c5; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:agg_kernel -s 2 -c 1 -f -o gpurun_out/prof_agg_$w \
     python bench.py --workload $w --steps 1 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/ncu_full_run_$w.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c3.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/ncu_launch_run.log 2>&1
# the other kernels of the path: one full-set capture each (scan, SHIFT-R-AND, BLOB decode, GAP merge, slab gather, rs_index build / rank / select)
timeout 600 ncu --set full --clock-control none -k regex:"scan_kernel|shift_and_kernel|blob_decode_kernel|blob_walk_kernel|blob_entropy_kernel|gap_merge_kernel|slab_gather_kernel" -c 16 -f -o gpurun_out/prof_other \
   python -m pytest tests -m gpu -q -k "scan_vs_oracle or shift_right_and or deserialize_to_device_vs_golden or binop_result or upload_slabs" > gpurun_out/ncu_other.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:rs_ -c 12 -f -o gpurun_out/prof_rs python scripts/bench_c4.py 1000000 > gpurun_out/ncu_rs.log 2>&1
BMB200_TRACE=1 timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/bench_c3_full.err | tail -1 | tee gpurun_out/bench_c3_full.json | cut -c1-400
grep -a "set_upload_vectors\|set_upload_slabs" gpurun_out/bench_c3_full.err | tail -9 > gpurun_out/trace_upload_c3.txt
grep -a "aggregator::run\|result_fetch_view" gpurun_out/bench_c3_full.err | tail -6 > gpurun_out/trace_warm_c3.txt
timeout 900 python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 2>/dev/null | tail -1 | tee gpurun_out/bench_ref.json | cut -c1-300
timeout 600 python bench.py --workload c2 --steps 20 --no-e2e-slab 2>/dev/null | tail -1 | tee gpurun_out/bench_c2_full.json | cut -c1-300
timeout 900 python bench.py --workload c5 --no-e2e --steps 10 2>/dev/null | tail -1 | tee gpurun_out/bench_c5_shard.json | cut -c1-300
python scripts/bench_c4.py 2>&1 | tail -1 > gpurun_out/bench_c4.json
python scripts/bench_scan.py 2>&1 | tail -1 > gpurun_out/bench_scan.json
python -c "
import json
d=json.load(open('gpurun_out/bench_c4.json'))
for k,v in d['results'].items(): print(k, {a:round(b,3) for a,b in v['gpu'].items()}, v.get('parity'), v.get('fraction_of_sector_bound'))"
python scripts/ncu_summary.py c3_agg_kernel_and_sub=gpurun_out/prof_agg_c3.ncu-rep c2_agg_kernel_or=gpurun_out/prof_agg_c2.ncu-rep c5_agg_kernel_or=gpurun_out/prof_agg_c5.ncu-rep > gpurun_out/ncu_agg_kernel.json
ncu -i gpurun_out/prof_other.ncu-rep --page raw --csv > gpurun_out/ncu_other_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_rs.ncu-rep --page raw --csv > gpurun_out/ncu_rs_raw.csv 2>/dev/null
# gpurun brings back at most 64 MiB: keep the two reports whose source pages are read (C3, C5), export the rest to CSV and drop them
ncu -i gpurun_out/prof_agg_c2.ncu-rep --page raw --csv > gpurun_out/ncu_agg_c2_raw.csv 2>/dev/null
rm -f gpurun_out/prof_other.ncu-rep gpurun_out/prof_rs.ncu-rep gpurun_out/prof_agg_c2.ncu-rep
du -sh gpurun_out; ls -la gpurun_out | tail -32
