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


class ShardedRS:
    """rank / select over a vector whose block columns are sharded by block range.

    Rank g holds columns [lo, hi) of the vector and a LOCAL rs_index over them (bmb200_rs_build on its shard: the reference's
    rs_index, src/bmrs.h:40-155, with superblock rows that never straddle ranks because shards are superblock-aligned).  The
    only state shared between ranks is the G shard cardinalities: ONE all_gather at construction, then an exclusive scan
    (`prefix`).  Every rank sees the same query batch; a query is answered by the rank that owns it

        count_to(pos)  (inclusive rank, src/bm.h:3120):  owner = the shard holding bit `pos`;
                        answer = prefix[owner] + local count_to(pos - lo * 65536); positions past the last shard -> grand total
        select(r)      (1-based, src/bm.h:5350):          owner = the shard with prefix[g] < r <= prefix[g] + total[g];
                        answer = lo * 65536 + local select(r - prefix[g]); r == 0 or r > grand total -> not found

    and the answers meet in ONE all_reduce(SUM) of the answer vector (non-owners contribute zeros) on the stream the local
    kernels ran on.  `local_rank(pos_local) -> ranks` and `local_select(r_local) -> (pos_local, found)` are the shard-local
    batched kernels (`device_rs_callables` wraps bmb200_rank_batch_dev / bmb200_select_batch_dev; the gloo test wraps the oracle).
    """

    def __init__(self, local_total: int, local_rank, local_select, n_blocks: int, dist=None, device="cpu"):
        self.dist = dist if (dist is not None and dist.is_initialized() and dist.get_world_size() > 1) else None
        self.world = self.dist.get_world_size() if self.dist else 1
        self.rank_id = self.dist.get_rank() if self.dist else 0
        self.n_blocks, self.device = n_blocks, device
        self.local_rank, self.local_select = local_rank, local_select
        self.lo, self.hi = shard_range(n_blocks, self.world, self.rank_id)
        mine = torch.tensor([int(local_total)], dtype=torch.int64, device=device)
        if self.dist:
            allt = torch.empty(self.world, dtype=torch.int64, device=device)
            self.dist.all_gather_into_tensor(allt, mine)
        else:
            allt = mine
        self.totals = allt
        self.prefix = torch.cumsum(allt, 0) - allt                  # exclusive scan of the shard cardinalities
        self.grand_total = int(allt.sum().item())
        self._p0 = int(self.prefix[self.rank_id].item()); self._t = int(allt[self.rank_id].item())   # host copies: no sync per query batch

    def _reduce(self, t: torch.Tensor) -> torch.Tensor:
        if self.dist:
            self.dist.all_reduce(t)                                  # SUM; every query has exactly one non-zero contributor
        return t

    def rank(self, pos: torch.Tensor) -> torch.Tensor:
        """pos: int64 tensor of global bit positions (same batch on every rank) -> int64 inclusive ranks, complete on every rank."""
        lo_bit, hi_bit = self.lo * 65536, self.hi * 65536
        out = torch.zeros_like(pos)
        own = (pos >= lo_bit) & (pos < hi_bit)
        idx = own.nonzero(as_tuple=True)[0]
        if idx.numel():
            out[idx] = self.local_rank((pos[idx] - lo_bit).contiguous()) + self._p0
        if self.rank_id == self.world - 1:                           # past the last indexed bit: the grand total (src/bm.h:3132-3136)
            out[pos >= self.n_blocks * 65536] = self.grand_total
        return self._reduce(out)

    def select(self, r: torch.Tensor):
        """r: int64 tensor of 1-based global ranks -> (int64 positions, bool found), complete on every rank."""
        p0, t = self._p0, self._t
        pos = torch.zeros_like(r); found = torch.zeros_like(r)
        own = (r > p0) & (r <= p0 + t)
        idx = own.nonzero(as_tuple=True)[0]
        if idx.numel():
            lp, lf = self.local_select((r[idx] - p0).contiguous())
            pos[idx] = torch.where(lf, lp + self.lo * 65536, torch.zeros_like(lp))
            found[idx] = lf.to(found.dtype)
        pos = self._reduce(pos); found = self._reduce(found)
        return pos, found > 0


def device_rs_callables(rs):
    """(local_rank, local_select) over a bitmagic_b200.DeviceRS for ShardedRS: CUDA int64 tensors in and out, the batched
    kernels read / write them in place through the *_dev entry points of the C ABI (no host round trip)."""
    def same_stream(t: torch.Tensor):
        # the kernels run on the CONTEXT's stream while torch produced `t` (and will consume the answers) on ITS current stream:
        # both must be the same stream, otherwise nothing orders the two (ctx.set_stream(torch.cuda.current_stream().cuda_stream))
        cur = torch.cuda.current_stream(t.device).cuda_stream
        if rs.ctx.get_stream() != cur:
            raise RuntimeError("device_rs_callables: the bmb200 context runs on another CUDA stream than torch's current stream; "
                               "call ctx.set_stream(torch.cuda.current_stream().cuda_stream) first")

    def local_rank(pos_local: torch.Tensor) -> torch.Tensor:
        same_stream(pos_local)
        out = torch.empty_like(pos_local)
        rs.rank_dev(pos_local.data_ptr(), pos_local.numel(), out.data_ptr())
        return out

    def local_select(r_local: torch.Tensor):
        same_stream(r_local)
        pos = torch.empty_like(r_local)
        found = torch.empty(r_local.numel(), dtype=torch.uint8, device=r_local.device)
        rs.select_dev(r_local.data_ptr(), r_local.numel(), pos.data_ptr(), found.data_ptr())
        return pos, found > 0
    return local_rank, local_select
