"""Multi-GPU sharding of the aggregation path: block-range partition + one popcount exchange.

Every block column is independent for AND/OR/XOR/SUB (the reference loops (i,j) with no carried state,
src/bmaggregator.h:1113-1121,1184-1218), so rank g of G owns a contiguous range of block columns of EVERY
vector, aligned to 256-block superblocks (rs_index rows never straddle ranks).  The only exchange is the
per-block popcount vector (4 B per column) and the global cardinality -- enqueued on the aggregation stream
through torch.distributed (NCCL on GPUs, gloo in the CPU tests): ONE all_gather; the cardinality is its local sum.
"""
from __future__ import annotations

import torch

SUPERBLOCK = 256


def shard_range(n_blocks: int, world: int, rank: int, align: int = SUPERBLOCK) -> tuple[int, int]:
    """[lo, hi) block columns owned by `rank`; superblock-aligned, sizes differ by at most one superblock."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    n_sb = (n_blocks + align - 1) // align
    lo_sb = (n_sb * rank) // world
    hi_sb = (n_sb * (rank + 1)) // world
    return min(lo_sb * align, n_blocks), min(hi_sb * align, n_blocks)


def shard_range_with_halo(n_blocks: int, world: int, rank: int, halo: int = 1, align: int = SUPERBLOCK) -> tuple[int, int, int]:
    """(lo_stored, lo, hi): the range a rank must HOLD for SHIFT-R-AND.  result[p] = AND_k v_k[p - s_k] reads up to n-1 bits of
    block lo - 1 (csrc/shift_kernel.cuh), so a shard stores `halo` extra block columns in front of the ones it owns and
    aggregates with nb_from = lo - lo_stored; everything else is the plain block-range partition."""
    lo, hi = shard_range(n_blocks, world, rank, align)
    return max(0, lo - halo), lo, hi


def shard_sizes(n_blocks: int, world: int, align: int = SUPERBLOCK) -> list[int]:
    return [hi - lo for lo, hi in (shard_range(n_blocks, world, r, align) for r in range(world))]


def exchange_popcounts(pop_local: torch.Tensor, n_blocks: int, dist=None, out: torch.Tensor | None = None):
    """All ranks end with the full per-block popcount vector [n_blocks] and the global cardinality.
    pop_local: int32 tensor with this rank's shard (length shard_sizes[rank]); runs on pop_local's device/stream."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return pop_local, pop_local.sum(dtype=torch.int64).reshape(1)
    world = dist.get_world_size()
    sizes = shard_sizes(n_blocks, world)
    if out is None:
        out = torch.empty(n_blocks, dtype=pop_local.dtype, device=pop_local.device)
    if len(set(sizes)) == 1:
        dist.all_gather_into_tensor(out, pop_local)
    else:                                   # ragged shards: pad to the largest, gather, drop the padding
        m = max(sizes)
        padded = torch.zeros(m, dtype=pop_local.dtype, device=pop_local.device)
        padded[: pop_local.numel()] = pop_local
        buf = torch.empty(world * m, dtype=pop_local.dtype, device=pop_local.device)
        dist.all_gather_into_tensor(buf, padded)
        o = 0
        for r, sz in enumerate(sizes):
            out[o:o + sz] = buf[r * m: r * m + sz]
            o += sz
    # the per-block vector is complete on every rank now, so the global cardinality needs no second collective
    card = out.sum(dtype=torch.int64).reshape(1)
    return out, card
