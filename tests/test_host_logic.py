"""Host-side containers (numpy only): packing, GAP encode/decode, compare; checked against the oracle."""
import numpy as np

import bitmagic_b200 as bm
from bitmagic_b200 import hostfmt as hf
import gen
import orclib


def test_gap_roundtrip_and_oracle_agreement():
    rng = np.random.default_rng(0)
    for runs in [1, 2, 5, 64, 500, 1275]:
        w = gen.block_with_runs(rng, runs)
        g = hf.bits_to_gap(w)
        assert (int(g[0]) >> 3) == runs == hf.calc_change(w)
        assert g[-1] == 65535
        assert np.array_equal(hf.gap_to_bits(g), w)
        out = np.zeros(2048, np.uint32)
        orclib.oracle().orc_gap_convert_to_bitset(orclib.ptr(out), orclib.ptr(g))
        assert np.array_equal(out, w)
        assert orclib.oracle().orc_gap_bit_count(orclib.ptr(g)) == int(hf.words_to_bits(w).sum())


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(1)
    vecs = gen.mixed_vectors(rng, 7, 5)
    ps = bm.PackedSet.pack(vecs)
    assert ps.desc.size == 7 * 5 and ps.bit_base[0] == 0 and ps.gap_base[0] == 0
    for v, bv in enumerate(vecs):
        back = ps.vector(v)
        assert np.array_equal(back.kind, bv.kind)
        assert back.compare(bv) == 0
    # column-major: blocks of one column are contiguous and ordered by vector
    for nb in range(5):
        rels = [int(ps.desc[nb * 7 + v]) >> 2 for v in range(7) if (int(ps.desc[nb * 7 + v]) & 3) == bm.BLK_BIT]
        assert rels == list(range(len(rels)))
    tmp = np.zeros(2048, np.uint32)
    import ctypes as C
    c = orclib._pc(ps)
    for v in range(7):
        for nb in range(5):
            orclib.oracle().orc_expand_block(C.byref(c), v, nb, orclib.ptr(tmp))
            assert np.array_equal(tmp, vecs[v].block_words(nb))


def test_optimize_kinds():
    v = bm.BVector(4)
    v.set_bits(0, np.zeros(2048, np.uint32)); v.set_bits(1, np.full(2048, 0xFFFFFFFF, np.uint32))
    v.set_bits(2, np.full(2048, 0xAAAAAAAA, np.uint32))
    w = np.zeros(2048, np.uint32); w[5] = 0xF0; v.set_bits(3, w)
    v.optimize()
    assert list(v.kind) == [bm.BLK_NULL, bm.BLK_FULL, bm.BLK_BIT, bm.BLK_GAP]
    assert v.count() == 65536 + 32768 + 4
    assert v.calc_stat() == {"bit_blocks": 1, "gap_blocks": 1, "full_blocks": 1}


def test_from_positions_and_positions():
    pos = [0, 1, 65535, 65536, 200000, 3 * 65536 - 1]
    v = bm.BVector.from_positions(pos, 4)
    assert list(v.positions()) == sorted(pos)
    assert v.count() == len(pos)
