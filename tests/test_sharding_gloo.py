"""N>1 host logic on CPU: world_size-2 gloo.  Each rank aggregates ITS block range (here with the oracle,
the CPU checker), the package's exchange_popcounts() merges the shards, and every rank must hold exactly
what a single unsharded pass produces."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, n_blocks, q):
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bitmagic_b200 as bm
    from bitmagic_b200.sharding import exchange_popcounts, shard_range
    import gen, orclib
    rng = np.random.default_rng(123)                      # same inputs on every rank
    ps = bm.PackedSet.pack(gen.mixed_vectors(rng, 6, n_blocks, p_null=0.3, p_gap=0.5))
    lo, hi = shard_range(n_blocks, world, rank)
    g0, g1 = [0, 1], list(range(2, 6))
    pop_local = orclib.oracle_aggregate(ps, bm.OP_AND_SUB, g0, g1, bm.F_OPT_COMPRESS, lo, hi)[1] if hi > lo else np.zeros(0, np.uint32)
    full, card = exchange_popcounts(torch.from_numpy(pop_local.astype(np.int32)), n_blocks, dist)
    ref = orclib.oracle_aggregate(ps, bm.OP_AND_SUB, g0, g1, bm.F_OPT_COMPRESS)[1]
    ok = np.array_equal(full.numpy().astype(np.uint32), ref) and int(card.item()) == int(ref.sum())
    q.put((rank, ok, lo, hi))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_blocks", [512, 600])
def test_block_range_sharding_world2(n_blocks):
    world, port = 2, 29500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, n_blocks, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
    res.sort()
    assert all(ok for _, ok, _, _ in res), res
    assert res[0][2] == 0 and res[0][3] == res[1][2] and res[1][3] == n_blocks
    assert res[0][3] % 256 == 0


def test_shard_range_properties():
    from bitmagic_b200.sharding import shard_range, shard_sizes
    for n_blocks in (1, 255, 256, 257, 4096, 16384, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            rs = [shard_range(n_blocks, world, r) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n_blocks
            for a, b in zip(rs, rs[1:]):
                assert a[1] == b[0] and (a[1] % 256 == 0 or a[1] == n_blocks)
            assert sum(shard_sizes(n_blocks, world)) == n_blocks
    assert shard_sizes(16384, 8) == [2048] * 8


def test_shift_and_shards_carry_a_one_column_halo():
    from bitmagic_b200.sharding import shard_range, shard_range_with_halo
    for nb, world in ((16384, 8), (1000, 3), (256, 2)):
        prev_hi = 0
        for r in range(world):
            st, lo, hi = shard_range_with_halo(nb, world, r)
            assert (lo, hi) == shard_range(nb, world, r) and lo == prev_hi
            assert st == max(0, lo - 1)
            prev_hi = hi
        assert prev_hi == nb


def _rs_worker(rank, world, port, n_blocks, q):
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bitmagic_b200 as bm
    from bitmagic_b200.sharding import ShardedRS, shard_range
    import gen, orclib
    rng = np.random.default_rng(77)                       # same vector and same queries on every rank
    vec = gen.mixed_vectors(rng, 1, n_blocks, p_null=0.3, p_full=0.1, p_gap=0.4)[0]
    whole = bm.PackedSet.pack([vec])
    lo, hi = shard_range(n_blocks, world, rank)
    shard = bm.PackedSet.pack([vec.slice(lo, hi)]) if hi > lo else None
    total = int(sum(vec.slice(lo, hi).count() for _ in [0])) if hi > lo else 0

    def local_rank(p):
        return torch.from_numpy(orclib.oracle_rank(shard, 0, p.numpy().astype(np.uint64)).astype(np.int64))

    def local_select(r):
        pos, found = orclib.oracle_select(shard, 0, r.numpy().astype(np.uint64))
        return torch.from_numpy(pos.astype(np.int64)), torch.from_numpy(found)

    srs = ShardedRS(total, local_rank, local_select, n_blocks, dist)
    card = vec.count()
    pos = np.concatenate([rng.integers(0, n_blocks * 65536, 4000), [0, 65535, 65536, 256 * 65536 - 1, 256 * 65536, n_blocks * 65536 - 1,
                          n_blocks * 65536, n_blocks * 65536 + 12345]]).astype(np.int64)
    rk = np.concatenate([rng.integers(1, card + 1, 4000), [0, 1, card, card + 1, card + 1000]]).astype(np.int64)
    got_rank = srs.rank(torch.from_numpy(pos)).numpy()
    got_pos, got_found = srs.select(torch.from_numpy(rk))
    want_rank = orclib.oracle_rank(whole, 0, pos.astype(np.uint64)).astype(np.int64)
    want_pos, want_found = orclib.oracle_select(whole, 0, rk.astype(np.uint64))
    ok = (srs.grand_total == card and np.array_equal(got_rank, want_rank) and np.array_equal(got_found.numpy(), want_found)
          and np.array_equal(got_pos.numpy()[want_found], want_pos.astype(np.int64)[want_found]))
    q.put((rank, bool(ok), lo, hi))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_blocks", [512, 700])
def test_sharded_rank_select_world2(n_blocks):
    """rank / select over a block-range sharded vector (ShardedRS: one all_gather of shard cardinalities, one all_reduce of the
    answers) == the unsharded oracle, including positions past the end, rank 0 and ranks above the cardinality."""
    world, port = 2, 31500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_rs_worker, args=(r, world, port, n_blocks, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res), res
