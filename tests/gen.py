"""Seeded input generators shared by the tests (host-side numpy only)."""
from __future__ import annotations

import numpy as np

import bitmagic_b200 as bm
from bitmagic_b200.hostfmt import bits_to_gap, bits_to_words, BLOCK_BITS


def gap_from_runs(ends, first):
    """GAP block from explicit inclusive run ends (last must be 65535) and first-run value."""
    ends = np.asarray(ends, dtype=np.uint16)
    n = ends.size
    out = np.empty(n + 1, dtype=np.uint16)
    out[1:] = ends
    lvl = 0 if n <= 124 else 1 if n <= 252 else 2 if n <= 508 else 3
    out[0] = (first & 1) | (lvl << 1) | (n << 3)
    return out


def block_with_runs(rng, n_runs):
    """Bit-block (words) with exactly n_runs runs at random boundaries."""
    cuts = np.sort(rng.choice(np.arange(1, BLOCK_BITS), size=n_runs - 1, replace=False)) if n_runs > 1 else np.zeros(0, int)
    bits = np.zeros(BLOCK_BITS, np.uint8)
    val = int(rng.integers(0, 2))
    prev = 0
    for c in list(cuts) + [BLOCK_BITS]:
        bits[prev:c] = val
        val ^= 1
        prev = c
    return bits_to_words(bits)


def mixed_vectors(rng, n_vec, n_blocks, p_null=0.15, p_full=0.05, p_gap=0.4):
    """Vectors whose blocks mix NULL / FULL / bit / GAP with varied densities and run structures."""
    vecs = []
    for _ in range(n_vec):
        v = bm.BVector(n_blocks)
        dens = float(10 ** rng.uniform(-3.3, -0.3))
        for nb in range(n_blocks):
            u = rng.random()
            if u < p_null:
                continue
            if u < p_null + p_full:
                v.set_full(nb)
            elif u < p_null + p_full + p_gap:
                style = rng.integers(0, 5)
                if style == 0:      # sparse random bits
                    w = bits_to_words(rng.random(BLOCK_BITS) < min(dens, 0.008))
                elif style == 1:    # few long runs
                    w = block_with_runs(rng, int(rng.integers(1, 40)))
                elif style == 2:    # many runs, near the GAP limit
                    w = block_with_runs(rng, int(rng.integers(900, 1276)))
                elif style == 3:    # word-aligned runs
                    bits = np.zeros(BLOCK_BITS, np.uint8)
                    for s in rng.choice(2048, size=20, replace=False):
                        bits[s * 32:(s + int(rng.integers(1, 6))) * 32] = 1
                    w = bits_to_words(bits)
                else:               # inverse of sparse (mostly ones)
                    w = bits_to_words(rng.random(BLOCK_BITS) >= 0.003)
                g = bits_to_gap(w)
                if (int(g[0]) >> 3) < 1276:
                    v.set_gap(nb, g)
                else:
                    v.set_bits(nb, w)
            else:
                v.set_bits(nb, bits_to_words(rng.random(BLOCK_BITS) < dens))
        vecs.append(v)
    return vecs


def edge_vectors(n_blocks=4):
    """Hand-made edge cases: all-zero / all-one GAP blocks, single-bit runs at word borders, FULL, NULL."""
    vs = []
    v = bm.BVector(n_blocks); v.set_gap(0, gap_from_runs([65535], 0)); v.set_gap(1, gap_from_runs([65535], 1)); vs.append(v)
    v = bm.BVector(n_blocks); v.set_gap(0, gap_from_runs([0, 65535], 1)); v.set_gap(1, gap_from_runs([65534, 65535], 0))
    v.set_gap(2, gap_from_runs([30, 31, 32, 63, 64, 65535], 0)); vs.append(v)
    v = bm.BVector(n_blocks); v.set_full(0); v.set_full(2); v.set_bits(1, np.full(2048, 0xFFFFFFFF, np.uint32)); vs.append(v)
    v = bm.BVector(n_blocks); v.set_bits(0, np.full(2048, 0xAAAAAAAA, np.uint32)); v.set_bits(3, np.full(2048, 0x55555555, np.uint32)); vs.append(v)
    v = bm.BVector(n_blocks); w = np.zeros(2048, np.uint32); w[0] = 1; w[2047] = 0x80000000; v.set_bits(1, w)
    v.set_gap(3, gap_from_runs([21823, 21824, 43647, 43648, 65535], 1)); vs.append(v)
    v = bm.BVector(n_blocks); vs.append(v)   # all NULL
    return vs


def entropy_vectors(rng, n_vec=14, n_blocks=10):
    """Vectors that drive bm::serializer<> (levels 3..6) through its entropy-coded encodings: iid bits over five decades of
    density and their complements (interpolative arrays, plain and inverted), run lists with regular strides (delta-range
    reduction, min0/min1), runs with isolated single-bit holes and spikes (GAP exception lists), clustered bits (windowed
    restore), dense blocks with word-aligned structure, very sparse super-blocks (super-block position lists)."""
    vecs = []
    for k in range(n_vec):
        v = bm.BVector(n_blocks)
        for nb in range(n_blocks):
            style = (k + nb) % 9
            bits = np.zeros(BLOCK_BITS, np.uint8)
            if style == 0:                       # iid, density 10^-4.5 .. 0.3
                bits = (rng.random(BLOCK_BITS) < 10 ** rng.uniform(-4.5, -0.5)).astype(np.uint8)
            elif style == 1:                     # complement of sparse
                bits = (rng.random(BLOCK_BITS) >= 10 ** rng.uniform(-4.0, -1.0)).astype(np.uint8)
            elif style == 2:                     # regular stride runs with jitter
                stride = int(rng.integers(20, 400)); rl = int(rng.integers(1, max(2, stride // 2)))
                for s in range(int(rng.integers(0, stride)), BLOCK_BITS - stride, stride):
                    s2 = s + int(rng.integers(0, 3)); bits[s2:s2 + rl + int(rng.integers(0, 2))] = 1
            elif style == 3:                     # long runs with isolated holes and spikes
                cuts = np.sort(rng.choice(np.arange(1, BLOCK_BITS), size=int(rng.integers(4, 300)), replace=False))
                val = int(rng.integers(0, 2)); prev = 0
                for c in list(cuts) + [BLOCK_BITS]:
                    bits[prev:c] = val; val ^= 1; prev = c
                flips = rng.choice(BLOCK_BITS, size=int(rng.integers(1, 200)), replace=False)
                bits[flips] ^= 1
            elif style == 4:                     # clusters
                for _ in range(int(rng.integers(1, 30))):
                    c = int(rng.integers(0, BLOCK_BITS - 600)); n = int(rng.integers(2, 300))
                    bits[c + rng.integers(0, 512, n)] = 1
            elif style == 5:                     # a handful of bits (super-block lists when the whole vector is like this)
                bits[rng.choice(BLOCK_BITS, size=int(rng.integers(1, 12)), replace=False)] = 1
            elif style == 6:                     # dense random words in a few waves
                w = np.zeros(2048, np.uint32)
                for s in rng.choice(64, size=int(rng.integers(1, 20)), replace=False):
                    w[s * 32:(s + 1) * 32] = rng.integers(0, 2**32, 32, dtype=np.uint64).astype(np.uint32)
                bits = np.unpackbits(w.view(np.uint8), bitorder="little")
            elif style == 7:                     # short runs of equal length, equal gaps (min0 / min1 > 1)
                g0 = int(rng.integers(3, 60)); g1 = int(rng.integers(2, 40)); p = int(rng.integers(0, 50))
                while p + g1 < BLOCK_BITS:
                    bits[p:p + g1 + int(rng.integers(0, 3))] = 1; p += g1 + g0 + int(rng.integers(0, 4))
            else:                                # mid density iid (bit-block or inverted array territory)
                bits = (rng.random(BLOCK_BITS) < rng.uniform(0.02, 0.98)).astype(np.uint8)
            if not bits.any():
                continue
            w = bits_to_words(bits)
            if int(np.count_nonzero(bits[1:] != bits[:-1])) + 1 < 1276 and rng.random() < 0.8:
                v.set_gap(nb, bits_to_gap(w))
            else:
                v.set_bits(nb, w)
        vecs.append(v)
    # whole vectors of very sparse blocks: the serializer folds each super-block into one position list (levels 5, 6)
    for dens in (1, 3, 40):
        v = bm.BVector(n_blocks)
        for nb in range(n_blocks):
            bits = np.zeros(BLOCK_BITS, np.uint8)
            bits[rng.choice(BLOCK_BITS, size=int(rng.integers(1, dens + 1)), replace=False)] = 1
            if nb % 4 != 3:
                v.set_gap(nb, bits_to_gap(bits_to_words(bits)))
        vecs.append(v)
    return vecs


def c1_vectors():
    """BASELINE.json configs[0] (SURVEY 8d "C1"): two vectors of 2^20 bits, each bit set iid with p = 0.10 (seeds 1, 2), no
    optimize() -> 16 bit-blocks each."""
    out = []
    for seed in (1, 2):
        bits = np.random.default_rng(seed).random(1 << 20) < 0.10
        v = bm.BVector(16)
        for nb in range(16):
            v.set_bits(nb, bits_to_words(bits[nb * BLOCK_BITS:(nb + 1) * BLOCK_BITS]))
        out.append(v)
    return out


def block_with_exact_runs(runs):
    """Bit-block of isolated bits with exactly `runs` runs (even counts start with a 1-run at bit 0)."""
    bits = np.zeros(BLOCK_BITS, np.uint8)
    n = (runs - 2) // 2 if runs % 2 == 0 else (runs - 1) // 2
    if runs % 2 == 0:
        bits[0] = 1
    bits[10 + 3 * np.arange(n)] = 1
    return bits_to_words(bits)


SB_MEMBER_RUNS = [3, 123, 124, 125, 126, 252, 253, 254, 508, 509, 510, 1275, 1276, 1277, 1279]


def superblock_threshold_vector(n_blocks=256):
    """One sparse super-block whose member blocks sit right at the GAP capacity levels (124 / 252 / 508 / 1276 runs): the serializer
    folds it into ONE set_sblock_bienc_v3 token at levels 5 / 6, and bm::deserialize rebuilds the members bit by bit under BM_GAP."""
    v = bm.BVector(n_blocks)
    for i, r in enumerate(SB_MEMBER_RUNS):
        w = block_with_exact_runs(r)
        if r < 1276:
            v.set_gap(2 * i, bits_to_gap(w))
        else:
            v.set_bits(2 * i, w)
    return v
