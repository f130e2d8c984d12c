"""Host-side mirror of the reference operator surface for the hot path.

Same names, argument meaning and error behaviour as the reference:
  * ``Aggregator``            <- ``bm::aggregator<BV>``            (reference src/bmaggregator.h:56-1060)
        add / reset / set_optimization / combine_or / combine_and / combine_and_sub / combine_shift_right_and
  * ``bit_and / bit_or / bit_xor / bit_sub`` <- 3-operand ``bvector::bit_*`` (src/bm.h:1745-1850)
  * ``count_and / count_or / count_xor / count_sub`` <- ``bm::count_*`` (src/bmalgo.h:48-51)
  * ``RSIndex / build_rs_index / count_to / select`` <- ``rs_index``, ``bvector::build_rs_index``,
        ``count_to``, ``select`` (src/bmrs.h:40-155, src/bm.h:2531,3120,5350)

Everything below the signatures is: pack the host block trees into the column-major arena, call the
C ABI (libbmb200.so, CUDA), unpack the result.  No set algebra happens in Python.
"""
from __future__ import annotations

import numpy as np

from . import capi
from .capi import (BLK_NULL, F_COUNT_ONLY, F_OPT_COMPRESS, F_OPT_NONE, F_OR_TARGET, OP_AND, OP_AND_SUB, OP_OR, OP_SHIFT_R_AND, OP_XOR)
from .hostfmt import BVector, PackedSet, result_to_bvector

OPT_NONE = 0       # bvector::opt_none
OPT_COMPRESS = 3   # bvector::opt_compress (reference src/bm.h:132-138)


def _run(ctx, vectors, op, g0, g1, opt_mode, count_only=False, spare_blocks=0):
    n_blocks = max(v.n_blocks for v in vectors) + spare_blocks
    dset = capi.DeviceSet.upload_vectors(ctx, vectors, n_blocks)
    try:
        flags = (F_OPT_COMPRESS if opt_mode else F_OPT_NONE) | (F_COUNT_ONLY if count_only else 0)
        res = capi.aggregate(ctx, dset, op, g0, g1, flags)
        try:
            total, any_ = res.total()
            if count_only:
                return None, total, any_
            kind, off, bits, gaps = res.fetch()
            return result_to_bvector(kind, off, bits, gaps), total, any_
        finally:
            res.free()
    finally:
        dset.free()


class Aggregator:
    """``bm::aggregator<bvector>`` with the block algebra on the GPU.

    Group 0 is the OR / AND group, group 1 the SUB group (reference src/bmaggregator.h:383-388).
    The target is *replaced* (resize_target(init_clear=true), src/bmaggregator.h:2215-2219).
    """

    def __init__(self, ctx: capi.Context | None = None):
        self.ctx = ctx or capi.default_context()
        self._groups: list[list[BVector]] = [[], []]
        self._opt = OPT_NONE

    def set_optimization(self, opt: int = OPT_COMPRESS):
        self._opt = opt

    def add(self, bv: BVector, agr_group: int = 0) -> int:
        if agr_group not in (0, 1):
            raise ValueError("agr_group must be 0 or 1")
        self._groups[agr_group].append(bv)
        return len(self._groups[agr_group])

    def reset(self):
        self._groups = [[], []]

    def set_range_hint(self, frm: int, to: int) -> bool:
        """aggregator::set_range_hint (src/bmaggregator.h:974-994): narrows find_first_and_sub to the blocks of [frm, to]; a range
        inside one block also masks that block.  Returns True for such a one-block range."""
        self._range = (int(frm), int(to))
        return (frm >> 16) == (to >> 16)

    def reset_range_hint(self):
        self._range = None

    def find_first_and_sub(self, bv_src_and: list[BVector] | None = None, bv_src_sub: list[BVector] | None = None):
        """aggregator::find_first_and_sub (src/bmaggregator.h:1457-1549) -> (found, index of the first bit of AND(group 0) - OR(group 1)).
        One launch over the hinted block range, the first non-empty column is located from the popcounts, one block comes back."""
        a = self._groups[0] if bv_src_and is None else bv_src_and
        s = self._groups[1] if bv_src_sub is None else bv_src_sub
        if not a:
            return False, 0
        vecs = list(a) + list(s)
        n_blocks = max(v.n_blocks for v in vecs)
        rng = getattr(self, "_range", None)
        lo, hi = (0, n_blocks) if rng is None else (rng[0] >> 16, min(n_blocks, (rng[1] >> 16) + 1))
        if lo >= hi:
            return False, 0
        one_block = rng is not None and (rng[0] >> 16) == (rng[1] >> 16)
        b0, b1 = ((rng[0] & 65535, rng[1] & 65535) if one_block else (0, 65535))
        dset = capi.DeviceSet.upload_vectors(self.ctx, vecs, n_blocks)
        try:
            res = capi.aggregate(self.ctx, dset, OP_AND_SUB, np.arange(len(a)), np.arange(len(a), len(vecs)), F_OPT_NONE, lo, hi)
            try:
                _, pop, _, _ = res.meta()
                for c in np.flatnonzero(pop):
                    kind, bits, gaps = res.fetch_column(int(c))
                    if kind == capi.BLK_FULL:
                        return True, (lo + int(c)) * 65536 + b0
                    if kind == capi.BLK_BIT:
                        pos = np.flatnonzero(np.unpackbits(bits.view(np.uint8), bitorder="little"))
                        pos = pos[(pos >= b0) & (pos <= b1)]
                        if pos.size:
                            return True, (lo + int(c)) * 65536 + int(pos[0])
                    if one_block:
                        break
                return False, 0
            finally:
                res.free()
        finally:
            dset.free()

    # --- C-style entry points (src/bmaggregator.h:503-540) ---
    def combine_or(self, bv_src: list[BVector] | None = None) -> BVector:
        src = self._groups[0] if bv_src is None else bv_src
        if not src:
            return BVector(0)                                  # n == 0 => clear(), :1105-1109
        res, _, _ = _run(self.ctx, src, OP_OR, np.arange(len(src)), None, self._opt)
        return res

    def combine_and(self, bv_src: list[BVector] | None = None) -> BVector:
        if bv_src is None:
            # member form routes through combine_and_sub with an empty SUB group, :1030-1039
            res, _ = self.combine_and_sub(self._groups[0], [], any_=False)
            return res
        if not bv_src:
            return BVector(0)
        res, _, _ = _run(self.ctx, bv_src, OP_AND, np.arange(len(bv_src)), None, self._opt)
        return res

    def combine_and_sub(self, bv_src_and: "list[BVector] | Pipeline | None" = None, bv_src_sub: list[BVector] | None = None,
                        any_: bool = False):
        if isinstance(bv_src_and, Pipeline):               # template<class TPipe> void combine_and_sub(TPipe&), :1291-1453
            return _run_pipeline(self.ctx, bv_src_and)
        a = self._groups[0] if bv_src_and is None else bv_src_and
        s = self._groups[1] if bv_src_sub is None else bv_src_sub
        if not a:
            return BVector(0), False                           # empty AND group => clear + false, :1170-1174
        vecs = list(a) + list(s)
        # combine_and_sub always stores through opt_copy_bit_block(opt_compress), :1209
        res, total, found = _run(self.ctx, vecs, OP_AND_SUB, np.arange(len(a)),
                                 np.arange(len(a), len(vecs)), OPT_COMPRESS)
        return res, found

    def combine_shift_right_and(self, bv_src_and: list[BVector] | None = None, any_: bool = False):
        """aggregator::combine_shift_right_and (src/bmaggregator.h:2494-2530): T_0 = v_0, T_k = (T_{k-1} >> 1) & v_k.
        Returns (result, found).  One spare block column receives the bits carried out of the last source block
        (the reference keeps walking top blocks while carry-overs are pending, :2506-2511)."""
        src = self._groups[0] if bv_src_and is None else bv_src_and
        if not src:
            return BVector(0), False                           # :2499-2503
        res, total, found = _run(self.ctx, src, OP_SHIFT_R_AND, np.arange(len(src)), None, self._opt, spare_blocks=1)
        return res, found

    def count_and_sub(self, bv_src_and, bv_src_sub) -> int:
        """counts-only mode of the pipeline (src/bmaggregator.h:1397-1398)."""
        if not bv_src_and:
            return 0
        vecs = list(bv_src_and) + list(bv_src_sub)
        _, total, _ = _run(self.ctx, vecs, OP_AND_SUB, np.arange(len(bv_src_and)),
                           np.arange(len(bv_src_and), len(vecs)), OPT_NONE, count_only=True)
        return total


class Pipeline:
    """``bm::aggregator<BV>::pipeline<Opt>`` (reference src/bmaggregator.h:222-341): many (AND set, SUB set) argument groups
    over a shared family of vectors, executed by ``Aggregator.combine_and_sub(pipeline)`` -- here ONE batched launch.

    ``make_results`` / ``compute_counts`` mirror ``agg_run_options`` (agg_opt_only_counts, agg_opt_bvect_and_counts ...).
    """

    class ArgGroups:
        def __init__(self):
            self.arg_bv0: list[BVector] = []
            self.arg_bv1: list[BVector] = []

        def add(self, bv: BVector, agr_group: int = 0) -> int:
            if agr_group not in (0, 1):
                raise ValueError("agr_group must be 0 or 1")
            lst = self.arg_bv1 if agr_group else self.arg_bv0
            lst.append(bv)
            return len(lst)

    def __init__(self, make_results: bool = True, compute_counts: bool = False):
        self.make_results, self.compute_counts = make_results, compute_counts
        self._groups: list[Pipeline.ArgGroups] = []
        self._complete = False
        self._or_target = False
        self._unique: list[BVector] = []
        self.bv_res_vector: list[BVector | None] = []
        self.bv_count_vector: list[int] = []
        self.or_target: BVector | None = None

    def add(self) -> "Pipeline.ArgGroups":
        if self._complete:
            raise RuntimeError("pipeline is complete: cannot add()")
        self._groups.append(Pipeline.ArgGroups())
        return self._groups[-1]

    def set_or_target(self, enable: bool = True):
        self._or_target = enable

    def complete(self):
        """Collect the unique input vectors (the reference's pipeline_bcache, src/bmaggregator.h:187-203)."""
        seen: dict[int, int] = {}
        self._unique = []
        for g in self._groups:
            for bv in g.arg_bv0 + g.arg_bv1:
                if id(bv) not in seen:
                    seen[id(bv)] = len(self._unique)
                    self._unique.append(bv)
        self._index = seen
        self._complete = True

    def is_complete(self) -> bool:
        return self._complete

    def size(self) -> int:
        return len(self._groups)

    def unique_vectors(self) -> int:
        return len(self._unique)

    def get_bv_res_vector(self):
        return self.bv_res_vector

    def get_bv_count_vector(self):
        return self.bv_count_vector


def _run_pipeline(ctx, pipe: Pipeline):
    if not pipe.is_complete():
        raise RuntimeError("pipeline.complete() must be called before execution")
    ng = pipe.size()
    pipe.bv_res_vector = [None] * ng
    pipe.bv_count_vector = [0] * ng
    pipe.or_target = None
    if not ng or not pipe._unique:
        return
    n_blocks = max(v.n_blocks for v in pipe._unique)
    dset = capi.DeviceSet.upload_vectors(ctx, pipe._unique, n_blocks)
    try:
        groups = [([pipe._index[id(b)] for b in g.arg_bv0], [pipe._index[id(b)] for b in g.arg_bv1]) for g in pipe._groups]
        flags = F_OPT_COMPRESS | (0 if pipe.make_results else F_COUNT_ONLY) | (F_OR_TARGET if pipe._or_target else 0)
        res = capi.aggregate_batch(ctx, dset, OP_AND_SUB, groups, flags)
        try:
            totals = res.group_totals(ng)
            pipe.bv_count_vector = [int(t) for t in totals]
            if pipe.make_results:
                kind, off, bits, gaps = res.fetch()
                for g in range(ng):
                    if totals[g]:                      # empty results stay NULL pointers in the reference
                        sl = slice(g * n_blocks, (g + 1) * n_blocks)
                        pipe.bv_res_vector[g] = result_to_bvector(kind[sl], off[sl], bits, gaps)
            if pipe._or_target:
                o = res.or_target(n_blocks)
                try:
                    pipe.or_target = result_to_bvector(*o.fetch())
                finally:
                    o.free()
        finally:
            res.free()
    finally:
        dset.free()


def _binop(op, a: BVector, b: BVector, opt_mode: int, ctx=None):
    """bvector::bit_and / bit_or / bit_xor / bit_sub (src/bm.h:6185,5973,6072,6403) through bmb200_binop: the result's block
    kinds follow combine_operation_block_* like the reference's (GAP x GAP merged on the device, clones keep their kind)."""
    ctx = ctx or capi.default_context()
    dset = capi.DeviceSet.upload_vectors(ctx, [a, b], max(a.n_blocks, b.n_blocks))
    try:
        res = capi.binop(ctx, dset, {"and": OP_AND, "or": OP_OR, "xor": OP_XOR, "sub": capi.OP_SUB}[op], 0, 1,
                         capi.F_OPT_COMPRESS if opt_mode == OPT_COMPRESS else capi.F_OPT_NONE)
        try:
            return result_to_bvector(*res.fetch())
        finally:
            res.free()
    finally:
        dset.free()


def bit_and(a: BVector, b: BVector, opt_mode: int = OPT_NONE, ctx=None) -> BVector:
    return _binop("and", a, b, opt_mode, ctx)


def bit_or(a: BVector, b: BVector, opt_mode: int = OPT_NONE, ctx=None) -> BVector:
    return _binop("or", a, b, opt_mode, ctx)


def bit_xor(a: BVector, b: BVector, opt_mode: int = OPT_NONE, ctx=None) -> BVector:
    return _binop("xor", a, b, opt_mode, ctx)


def bit_sub(a: BVector, b: BVector, opt_mode: int = OPT_NONE, ctx=None) -> BVector:
    return _binop("sub", a, b, opt_mode, ctx)


def bit_or_and(target: BVector, a: BVector, b: BVector, opt_mode: int = OPT_NONE, ctx=None) -> BVector:
    """bvector::bit_or_and (src/bm.h:1787,6283): target | (a & b) -- two launches."""
    return bit_or(target, bit_and(a, b, opt_mode, ctx), opt_mode, ctx)


def merge(target: BVector, src: BVector, ctx=None) -> BVector:
    """bvector::merge (src/bm.h:1000,5883): target | src (the reference may steal src's blocks; here src is left untouched)."""
    return bit_or(target, src, OPT_NONE, ctx)


def _count(op, a, b, ctx=None) -> int:
    ctx = ctx or capi.default_context()
    if op == "sub":
        return _run(ctx, [a, b], OP_AND_SUB, [0], [1], OPT_NONE, count_only=True)[1]
    return _run(ctx, [a, b], {"and": OP_AND, "or": OP_OR, "xor": OP_XOR}[op], [0, 1], None, OPT_NONE, count_only=True)[1]


def count_and(a, b, ctx=None) -> int:
    return _count("and", a, b, ctx)


def count_or(a, b, ctx=None) -> int:
    return _count("or", a, b, ctx)


def count_xor(a, b, ctx=None) -> int:
    return _count("xor", a, b, ctx)


def count_sub(a, b, ctx=None) -> int:
    return _count("sub", a, b, ctx)


class RSIndex:
    """``bvector::rs_index_type`` built on the GPU; answers batched count_to / select."""

    def __init__(self, bv: BVector, ctx: capi.Context | None = None):
        self.ctx = ctx or capi.default_context()
        self._dset = capi.DeviceSet.upload_vectors(self.ctx, [bv], bv.n_blocks)
        self._rs = capi.DeviceRS(self.ctx, self._dset, 0)
        self.n_blocks = bv.n_blocks

    def count(self) -> int:
        return self._rs.total()

    def fields(self):
        """(bcount, sub_count, sblock_count) as register_super_block receives them."""
        return self._rs.export()

    def count_to(self, pos) -> np.ndarray:
        """Inclusive rank: bits set in [0, pos] (bvector::count_to)."""
        return self._rs.rank(np.atleast_1d(pos))

    def rank_corrected(self, pos, bit_at_pos) -> np.ndarray:
        """count_to minus the bit at pos (bvector::rank_corrected src/bm.h:3229)."""
        return self.count_to(pos) - np.asarray(bit_at_pos, dtype=np.uint64)

    def select(self, rank):
        """1-based select; returns (pos, found) (bvector::select)."""
        return self._rs.select(np.atleast_1d(rank))

    def close(self):
        self._rs.free()
        self._dset.free()


def build_rs_index(bv: BVector, ctx=None) -> RSIndex:
    return RSIndex(bv, ctx)
