"""The oracle against the committed golden vectors (outputs of the unmodified reference)."""
import numpy as np
import pytest

import bitmagic_b200 as bm
import golden_util as gu
import orclib


@pytest.mark.parametrize("name", ["agg_mixed", "agg_edge", "agg_zipf"])
def test_oracle_aggregate_vs_golden(name):
    ps, cases = gu.load_agg(name)
    for case in cases:
        kind, pop, dig, nr, blk, gaps = orclib.oracle_aggregate(ps, case["op"], case["g0"], case["g1"], case["flags"])
        glen = np.where(kind == bm.BLK_GAP, (gaps[:, 0] >> 3) + 1, 0)
        flat = np.concatenate([gaps[c, :glen[c]] for c in range(len(kind))]) if glen.sum() else np.zeros(0, np.uint16)
        # XOR fixtures come from chained bit_xor, whose block kinds follow a different storage rule
        gu.check_agg_case(case, kind, pop, blk, flat if case["op"] != bm.OP_XOR else None, check_kind=case["op"] != bm.OP_XOR)
        assert case["any"] == bool(pop.sum())
        # digest / run-count are consistent with the result bits
        for c in range(ps.n_blocks):
            w = np.ascontiguousarray(blk[c])
            assert dig[c] == sum(1 << i for i in range(64) if w[32 * i:32 * i + 32].any())
            assert nr[c] == bm.hostfmt.calc_change(w)


def test_oracle_rs_vs_golden():
    ps, vs = gu.load_rs("rs_mixed")
    for v, g in enumerate(vs):
        bc, sc, sb = orclib.oracle_rs_build(ps, v)
        if int(g["total"]) == 0:
            assert bc.sum() == 0
            continue
        assert np.array_equal(bc, g["bcount"])
        nz = g["bcount"] > 0
        assert np.array_equal(sc[nz], g["sub"][nz])
        assert np.array_equal(sb, g["sb"])
        assert np.array_equal(orclib.oracle_rank(ps, v, g["pos"]), g["rank_out"])
        pos, found = orclib.oracle_select(ps, v, g["rank"])
        assert np.array_equal(found, g["sel_found"])
        assert np.array_equal(pos[found], g["sel_pos"][g["sel_found"]])


@pytest.mark.parametrize("name", ["scan_plain", "scan_nullable"])
def test_oracle_scan_vs_golden(name):
    """orc_scan on the reference's own planes == the committed answers of bm::sparse_vector_scanner<>."""
    ps, vals, nulls, cases = gu.load_scan(name)
    npl = ps.n_vec - 1
    for case in cases:
        kind, pop, dig, nr, blk, gaps = orclib.oracle_scan(ps, case["pred"], case["search"], 0, npl, npl, bm.F_OPT_COMPRESS)
        assert np.array_equal(blk, case["blk"]) and np.array_equal(pop, case["pop"])
        assert np.array_equal(pop.reshape(len(case["search"]), -1).sum(1), case["counts"])
    # the planes are the bit-transposed values
    live = np.ones(vals.size, bool) if nulls is None else nulls == 0
    kind, pop, dig, nr, blk, gaps = orclib.oracle_scan(ps, bm.SCAN_EQ, [55], 0, npl, npl, 0)
    assert int(pop.sum()) == int(((vals == 55) & live).sum())


@pytest.mark.parametrize("name", ["blobs", "blobs_entropy"])
def test_oracle_deserialize_vs_golden(name):
    """orc_deserialize on the committed serializer BLOBs (levels 0..2: explicit-length encodings; levels 3..6: gamma /
    interpolative / super-block encodings) == the committed bm::deserialize output."""
    nv, nb, blobs, kinds, blks, gapsf = gu.load_blobs(name)
    for level, bl in blobs.items():
        for v in range(nv):
            rc, kind, blk, gaps = orclib.oracle_deserialize(bl[v], nb)
            assert rc == 0
            assert np.array_equal(kind, kinds[level][v]) and np.array_equal(blk, blks[v])
            glen = np.where(kind == bm.BLK_GAP, (gaps[:, 0] >> 3) + 1, 0)
            flat = np.concatenate([gaps[c, :glen[c]] for c in range(nb)]) if glen.sum() else np.zeros(0, np.uint16)
            assert np.array_equal(flat, gapsf[level][v])


def test_device_decoder_host_build_vs_golden():
    """Host build of the product's BLOB walker / entropy decoder (oracle/blob_host_check.cpp over bitmagic_b200/csrc/blob_entropy.cuh)
    on the committed entropy-coded BLOBs == the committed bm::deserialize output (kinds of all blocks; bits / GAP words of the
    blocks decoded from entropy-coded tokens)."""
    nv, nb, blobs, kinds, blks, gapsf = gu.load_blobs("blobs_entropy")
    n_ent = 0
    for level, bl in blobs.items():
        for v in range(nv):
            rc, kind, dec, gw, blk, gaps, n = orclib.blob_host_check(bl[v], nb)
            assert rc == 0, f"level {level} vector {v}: rc={rc}"
            n_ent += n
            assert np.array_equal(kind, kinds[level][v])
            want, cur = {}, 0                                   # split the flat golden GAP words by their headers
            for c in np.flatnonzero(kind == bm.BLK_GAP):
                n = (int(gapsf[level][v][cur]) >> 3) + 1
                want[int(c)] = gapsf[level][v][cur:cur + n]; cur += n
            for c in np.flatnonzero(dec):
                if kind[c] == bm.BLK_BIT:
                    assert np.array_equal(blk[c], blks[v][c])
                else:
                    assert int(gw[c]) == want[int(c)].size and np.array_equal(gaps[c, :gw[c]], want[int(c)])
    assert n_ent > 300
