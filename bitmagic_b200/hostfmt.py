"""Host-side containers: the block tree stays on the host (north_star), the GPU sees a packed arena.

``BVector`` mirrors what ``bm::bvector<>``'s ``blocks_manager`` holds (reference src/bmblocks.h:540-579):
per 64 Kbit block slot one of four states -- NULL, FULL, bit-block (2048 x u32) or GAP block
(u16 header + inclusive run ends, reference src/bmfunc.h:1696-1725).  ``PackedSet`` is the
column-major arena of include/bmb200.h (``bmb200_packed_set``) that the C ABI consumes.

This module is host bookkeeping only (numpy); all set algebra runs in libbmb200.so on the GPU.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .capi import (BLK_BIT, BLK_FULL, BLK_GAP, BLK_NULL, BLOCK_BITS, BLOCK_WORDS, GAP_THRESHOLD,
                   GAP_UNIT_WORDS)

_GAP_LEVEL_LIMITS = (124, 252, 508, 1276)   # glen(level) - 4, reference src/bmconst.h:396-403, bmfunc.h:5418


def gap_level(length: int) -> int:
    for lvl, lim in enumerate(_GAP_LEVEL_LIMITS):
        if length <= lim:
            return lvl
    return 3


def words_to_bits(words: np.ndarray) -> np.ndarray:
    """2048 x u32 -> 65536 x u8 (bit p of the block at index p)."""
    return np.unpackbits(np.ascontiguousarray(words, dtype="<u4").view(np.uint8), bitorder="little")


def bits_to_words(bits: np.ndarray) -> np.ndarray:
    return np.packbits(bits.astype(np.uint8), bitorder="little").view("<u4").astype(np.uint32)


def calc_change(words: np.ndarray) -> int:
    """Number of runs in a bit-block (bit_block_calc_change, reference src/bmfunc.h:6040)."""
    b = words_to_bits(words)
    return int(np.count_nonzero(b[1:] != b[:-1])) + 1


def bits_to_gap(words: np.ndarray) -> np.ndarray:
    """Bit-block -> GAP block: header | run ends (bit_block_to_gap, reference src/bmfunc.h:5540)."""
    b = words_to_bits(words)
    ends = np.flatnonzero(b[1:] != b[:-1]).astype(np.uint32)
    n = ends.size + 1
    out = np.empty(n + 1, dtype=np.uint16)
    out[1:n] = ends
    out[n] = 65535
    out[0] = int(b[0]) | (gap_level(n) << 1) | (n << 3)
    return out


def gap_to_bits(gap: np.ndarray) -> np.ndarray:
    """GAP block -> 2048 x u32 (gap_convert_to_bitset, reference src/bmfunc.h:5232)."""
    n = int(gap[0]) >> 3
    ends = gap[1:n + 1].astype(np.int64)
    starts = np.concatenate(([0], ends[:-1] + 1))
    vals = (np.arange(n) & 1) ^ (int(gap[0]) & 1)
    b = np.zeros(BLOCK_BITS + 1, dtype=np.int8)
    np.add.at(b, starts[vals == 1], 1)
    np.add.at(b, ends[vals == 1] + 1, -1)
    return bits_to_words(np.cumsum(b[:BLOCK_BITS]) > 0)


def gap_words(gap: np.ndarray) -> int:
    return (int(gap[0]) >> 3) + 1


class BVector:
    """Host mirror of one bm::bvector<>'s block tree over ``n_blocks`` 64 Kbit slots."""

    def __init__(self, n_blocks: int):
        self.n_blocks = int(n_blocks)
        self.kind = np.zeros(self.n_blocks, dtype=np.uint8)
        self.blocks: dict[int, np.ndarray] = {}

    # ---- construction ----
    @classmethod
    def from_positions(cls, positions, n_blocks: int) -> "BVector":
        """Set the given bit positions; blocks become bit-blocks (an un-optimized bvector)."""
        bv = cls(n_blocks)
        pos = np.unique(np.asarray(positions, dtype=np.uint64))
        if pos.size and int(pos[-1]) >= n_blocks * BLOCK_BITS:
            raise IndexError("bit position beyond the vector")
        nbs = (pos >> np.uint64(16)).astype(np.int64)
        for nb in np.unique(nbs):
            inb = (pos[nbs == nb] & np.uint64(0xFFFF)).astype(np.int64)
            b = np.zeros(BLOCK_BITS, dtype=np.uint8)
            b[inb] = 1
            bv.kind[nb] = BLK_BIT
            bv.blocks[int(nb)] = bits_to_words(b)
        return bv

    @classmethod
    def from_words(cls, words: np.ndarray) -> "BVector":
        """Dense u32 words (multiple of 2048) -> bit-blocks for every non-zero block."""
        w = np.ascontiguousarray(words, dtype=np.uint32).reshape(-1, BLOCK_WORDS)
        bv = cls(w.shape[0])
        for nb in range(w.shape[0]):
            if w[nb].any():
                bv.kind[nb] = BLK_BIT
                bv.blocks[nb] = w[nb].copy()
        return bv

    @classmethod
    def random(cls, n_blocks: int, density: float, rng: np.random.Generator) -> "BVector":
        words = bits_to_words(rng.random(n_blocks * BLOCK_BITS) < density)
        return cls.from_words(words)

    def set_full(self, nb: int):
        self.kind[nb] = BLK_FULL
        self.blocks.pop(nb, None)

    def set_gap(self, nb: int, gap: np.ndarray):
        self.kind[nb] = BLK_GAP
        self.blocks[nb] = np.ascontiguousarray(gap, dtype=np.uint16)

    def set_bits(self, nb: int, words: np.ndarray):
        self.kind[nb] = BLK_BIT
        self.blocks[nb] = np.ascontiguousarray(words, dtype=np.uint32)

    def slice(self, lo: int, hi: int) -> "BVector":
        """Block columns [lo, hi) as a vector of their own (the shard a rank holds under block-range sharding)."""
        out = BVector(hi - lo)
        out.kind[:] = self.kind[lo:hi]
        out.blocks = {nb - lo: b for nb, b in self.blocks.items() if lo <= nb < hi}
        return out

    def optimize(self) -> "BVector":
        """In place, like bvector::optimize(opt_compress): empty -> NULL, all-ones -> FULL,
        runs < 1276 -> GAP (reference src/bmblocks.h:1414-1437)."""
        for nb in range(self.n_blocks):
            if self.kind[nb] != BLK_BIT:
                continue
            w = self.blocks[nb]
            runs = calc_change(w)
            if runs == 1:
                self.kind[nb] = BLK_FULL if w[0] else BLK_NULL
                del self.blocks[nb]
            elif runs < GAP_THRESHOLD:
                self.kind[nb] = BLK_GAP
                self.blocks[nb] = bits_to_gap(w)
        return self

    # ---- inspection ----
    def block_words(self, nb: int) -> np.ndarray:
        k = self.kind[nb]
        if k == BLK_NULL:
            return np.zeros(BLOCK_WORDS, dtype=np.uint32)
        if k == BLK_FULL:
            return np.full(BLOCK_WORDS, 0xFFFFFFFF, dtype=np.uint32)
        if k == BLK_BIT:
            return self.blocks[nb]
        return gap_to_bits(self.blocks[nb])

    def to_words(self) -> np.ndarray:
        return np.concatenate([self.block_words(nb) for nb in range(self.n_blocks)]) if self.n_blocks else np.zeros(0, np.uint32)

    def count(self) -> int:
        return int(sum(int(np.unpackbits(self.block_words(nb).view(np.uint8)).sum()) for nb in range(self.n_blocks)))

    def positions(self) -> np.ndarray:
        out = []
        for nb in range(self.n_blocks):
            if self.kind[nb] != BLK_NULL:
                out.append(np.flatnonzero(words_to_bits(self.block_words(nb))).astype(np.uint64) + np.uint64(nb * BLOCK_BITS))
        return np.concatenate(out) if out else np.zeros(0, np.uint64)

    def compare(self, other: "BVector") -> int:
        """0 when logically equal (the reference's own parity criterion, bvector::compare src/bm.h:3776)."""
        n = max(self.n_blocks, other.n_blocks)
        z = np.zeros(BLOCK_WORDS, dtype=np.uint32)
        for nb in range(n):
            a = self.block_words(nb) if nb < self.n_blocks else z
            b = other.block_words(nb) if nb < other.n_blocks else z
            if not np.array_equal(a, b):
                return 1
        return 0

    def calc_stat(self) -> dict:
        """bit / GAP block counts like bvector::calc_stat (reference src/bm.h:4010)."""
        return {"bit_blocks": int(np.count_nonzero(self.kind == BLK_BIT)),
                "gap_blocks": int(np.count_nonzero(self.kind == BLK_GAP)),
                "full_blocks": int(np.count_nonzero(self.kind == BLK_FULL))}


@dataclass
class PackedSet:
    """Host copy of bmb200_packed_set: column-major arena of n_vec vectors x n_blocks columns."""
    n_vec: int
    n_blocks: int
    desc: np.ndarray       # u32 [n_blocks * n_vec]
    bit_base: np.ndarray   # u64 [n_blocks + 1]
    gap_base: np.ndarray   # u64 [n_blocks + 1], 16-byte units
    bit_pool: np.ndarray   # u32
    gap_pool: np.ndarray   # u16

    @classmethod
    def pack(cls, vectors: list[BVector], n_blocks: int | None = None, gap_flat: bool = True) -> "PackedSet":
        """Walk the host block trees column by column (what the binding does with get_block_ptr(i,j)).
        gap_flat: store GAP blocks in the flat-streamable form (BMB200_DESC_GAP_FLAT, include/bmb200.h): lead pad
        0xFFFF iff the first run is 0, zero fill to the 16-byte unit.  False = raw blocks, no flags."""
        nv = len(vectors)
        nb_tot = n_blocks if n_blocks is not None else max(v.n_blocks for v in vectors)
        desc = np.zeros(nb_tot * nv, dtype=np.uint32)
        bb = np.zeros(nb_tot + 1, dtype=np.uint64)
        gb = np.zeros(nb_tot + 1, dtype=np.uint64)
        bits, gaps = [], []
        for nb in range(nb_tot):
            nbit = ngap = 0
            for v, bv in enumerate(vectors):
                k = int(bv.kind[nb]) if nb < bv.n_blocks else BLK_NULL
                rel = 0
                if k == BLK_BIT:
                    rel = nbit; nbit += 1
                    bits.append(bv.blocks[nb])
                elif k == BLK_GAP:
                    g = bv.blocks[nb]
                    n = gap_words(g)
                    pad = 1 if (gap_flat and not (int(g[0]) & 1)) else 0
                    units = (n + pad + GAP_UNIT_WORDS - 1) // GAP_UNIT_WORDS
                    padded = np.zeros(units * GAP_UNIT_WORDS, dtype=np.uint16)
                    if pad:
                        padded[0] = 0xFFFF
                    padded[pad:pad + n] = g[:n]
                    rel = ngap | (pad << 29) | ((1 << 28) if gap_flat else 0); ngap += units
                    gaps.append(padded)
                desc[nb * nv + v] = k | ((rel << 2) & 0xFFFFFFFF)
            bb[nb + 1] = bb[nb] + np.uint64(nbit)
            gb[nb + 1] = gb[nb] + np.uint64(ngap)
        bit_pool = np.concatenate(bits).astype(np.uint32) if bits else np.zeros(0, np.uint32)
        gap_pool = np.concatenate(gaps).astype(np.uint16) if gaps else np.zeros(0, np.uint16)
        return cls(nv, nb_tot, desc, bb, gb, np.ascontiguousarray(bit_pool), np.ascontiguousarray(gap_pool))

    def block(self, v: int, nb: int):
        """(kind, data) of block nb of vector v."""
        d = int(self.desc[nb * self.n_vec + v])
        k, rel = d & 3, d >> 2
        if k == BLK_BIT:
            o = (int(self.bit_base[nb]) + rel) * BLOCK_WORDS
            return k, self.bit_pool[o:o + BLOCK_WORDS]
        if k == BLK_GAP:
            o = (int(self.gap_base[nb]) + (rel & 0x0FFFFFFF)) * GAP_UNIT_WORDS + (rel >> 29)
            n = (int(self.gap_pool[o]) >> 3) + 1
            return k, self.gap_pool[o:o + n]
        return k, None

    def vector(self, v: int) -> BVector:
        bv = BVector(self.n_blocks)
        for nb in range(self.n_blocks):
            k, data = self.block(v, nb)
            bv.kind[nb] = k
            if data is not None:
                bv.blocks[nb] = np.array(data)
        return bv

    def kinds(self) -> np.ndarray:
        return (self.desc & 3).reshape(self.n_blocks, self.n_vec)

    def stored_bytes(self) -> int:
        return int(self.bit_base[-1]) * BLOCK_WORDS * 4 + int(self.gap_base[-1]) * GAP_UNIT_WORDS * 2


def result_to_bvector(kind: np.ndarray, off: np.ndarray, bits: np.ndarray, gaps: np.ndarray) -> BVector:
    """Per-vector flat form returned by bmb200_result_fetch -> BVector."""
    bv = BVector(kind.size)
    for c in range(kind.size):
        k = int(kind[c])
        bv.kind[c] = k
        if k == BLK_BIT:
            o = int(off[c]) * BLOCK_WORDS
            bv.blocks[c] = np.array(bits[o:o + BLOCK_WORDS])
        elif k == BLK_GAP:
            o = int(off[c])
            n = (int(gaps[o]) >> 3) + 1
            bv.blocks[c] = np.array(gaps[o:o + n])
    return bv
