#!/usr/bin/env python
"""BASELINE config 4 on N GPUs: the 2^32-bit vector is block-range sharded, every rank builds a local rs_index over its
shard and the query batches (same on every rank) are answered through bitmagic_b200.sharding.ShardedRS -- one all_gather of the shard
cardinalities at construction, one all_reduce(SUM) of the answer vector per batch (NCCL).

  python scripts/bench_c4_sharded.py [n_queries]                                   # N = 1
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_c4_sharded.py [n_queries]

The vector is the concatenation of the ranks' shards, shard g = iid 1 % density with seed 7 + g (the counter-based generator is keyed
by position inside a set, so each rank generates its own shard; at N = 1 this is exactly scripts/bench_c4.py's vector).  Timing: CUDA
events on the launching stream, barrier + synchronize on both sides, max over ranks; rank 0 prints one JSON line.  Checks that hold at
any N: count_to(select(r)) == r for every found r, select finds exactly the ranks 1..total, count_to is non-decreasing in pos and
count_to(last bit) == total == the all-reduced shard totals."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import bitmagic_b200 as bm                                                   # noqa: E402
from bitmagic_b200.sharding import ShardedRS, device_rs_callables, shard_range   # noqa: E402


def main():
    nq = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    nb = 65536
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = bm.Context(local)
    stream = torch.cuda.current_stream(dev)
    ctx.set_stream(stream.cuda_stream)
    lo, hi = shard_range(nb, world, rank)
    out = {}
    for label, optimize in (("bit_blocks", False), ("optimized", True)):
        dset = bm.DeviceSet.synth(ctx, 1, hi - lo, np.array([0.01]), np.array([7 + rank], np.uint64), optimize)
        rs = bm.DeviceRS(ctx, dset, 0); ctx.sync()
        srs = ShardedRS(rs.total(), *device_rs_callables(rs), nb, dist if world > 1 else None, dev)
        total = srs.grand_total
        gen = torch.Generator(device="cpu"); gen.manual_seed(8)                 # the same queries on every rank
        pos = torch.randint(0, nb * 65536, (nq,), generator=gen, dtype=torch.int64).to(dev)
        rk = torch.randint(1, total + 1, (nq,), generator=gen, dtype=torch.int64).to(dev)

        def timed(fn, reps=3):
            fn(); torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(reps):
                fn()
            b.record(stream); torch.cuda.synchronize(dev)
            ms = torch.tensor([a.elapsed_time(b) / reps], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return float(ms.item())

        rank_ms = timed(lambda: srs.rank(pos))
        sel_ms = timed(lambda: srs.select(rk))
        # size-independent checks
        sp, sf = srs.select(rk)
        back = srs.rank(sp)
        ok_roundtrip = bool(sf.all().item()) and bool((back == rk).all().item())
        beyond = srs.select(torch.tensor([0, total + 1], dtype=torch.int64, device=dev))[1]
        r_sorted = srs.rank(torch.sort(pos).values)
        ok_monotone = bool((r_sorted[1:] >= r_sorted[:-1]).all().item())
        r_end = srs.rank(torch.tensor([nb * 65536 - 1, nb * 65536 + 5], dtype=torch.int64, device=dev))
        ok_total = int(r_end[0].item()) == total == int(r_end[1].item())
        out[label] = {"bits_set": total, "rank_ms": rank_ms, "rank_Mq_per_s": nq / rank_ms / 1e3, "select_ms": sel_ms, "select_Mq_per_s": nq / sel_ms / 1e3,
                      "checks": {"rank_of_select_is_identity": ok_roundtrip, "select_rejects_0_and_total_plus_1": not bool(beyond.any().item()),
                                 "rank_monotone": ok_monotone, "rank_at_end_is_total": ok_total}}
        rs.free(); dset.free()
    if rank == 0:
        print(json.dumps({"workload": "c4 sharded: rank/select over a block-range sharded 2^32-bit vector, 1% density", "n_gpus": world,
                          "n_queries": nq, "exchange": "all_gather(totals) once + all_reduce(sum) of the answers per batch", "results": out}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
