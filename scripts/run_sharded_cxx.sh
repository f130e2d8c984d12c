#!/bin/bash
# two ranks of oracle/_ref/test_sharded (bm::b200::sharded_aggregator vs bm::aggregator), one per GPU, communicator id through a file
mkdir -p gpurun_out
rm -f /tmp/ncclid.bin
NCCL_DEBUG=WARN oracle/_ref/test_sharded 0 2 /tmp/ncclid.bin > gpurun_out/sharded_r0.log 2>&1 &
P0=$!
NCCL_DEBUG=WARN oracle/_ref/test_sharded 1 2 /tmp/ncclid.bin > gpurun_out/sharded_r1.log 2>&1 &
P1=$!
wait $P0; echo "rank0 rc=$?"; wait $P1; echo "rank1 rc=$?"
tail -15 gpurun_out/sharded_r0.log; tail -15 gpurun_out/sharded_r1.log
ldconfig -p | grep -i nccl
