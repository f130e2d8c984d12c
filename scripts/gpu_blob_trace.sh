#!/bin/bash
# where does bmb200_set_upload_blobs spend its time?  (BMB200_TRACE=1 phase timings; the traced numbers are serialised)
mkdir -p gpurun_out
for l in 6 4 2; do
  BMB200_TRACE=1 timeout -s KILL 200 python scripts/bench_blob.py 256 64 $l > gpurun_out/trace_blob_l$l.json 2> gpurun_out/trace_blob_l$l.err
  echo "== level $l"; tail -14 gpurun_out/trace_blob_l$l.err; cut -c1-400 gpurun_out/trace_blob_l$l.json
done
