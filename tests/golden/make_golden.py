"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libbmref.so, built from
/root/reference/src by oracle/Makefile).  Run in the build container:  python tests/golden/make_golden.py
Each fixture = a packed input set + the reference's outputs for several aggregator / rs_index calls."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import bitmagic_b200 as bm   # noqa: E402
import gen                   # noqa: E402
import orclib                # noqa: E402

OUT = Path(__file__).resolve().parent


def pack_fields(ps):
    return dict(n_vec=ps.n_vec, n_blocks=ps.n_blocks, desc=ps.desc, bit_base=ps.bit_base, gap_base=ps.gap_base,
                bit_pool=ps.bit_pool, gap_pool=ps.gap_pool)


def aggregate_fixture(name, vecs, cases):
    ps = bm.PackedSet.pack(vecs)
    d = pack_fields(ps)
    d["n_cases"] = len(cases)
    for i, (op, g0, g1, flags) in enumerate(cases):
        kind, pop, blk, gaps, any_ = orclib.ref_aggregate(ps, op, g0, g1, flags)
        d[f"c{i}_op"] = op; d[f"c{i}_g0"] = np.asarray(g0, np.uint32)
        d[f"c{i}_g1"] = np.asarray(g1 if g1 is not None else [], np.uint32); d[f"c{i}_flags"] = flags
        d[f"c{i}_kind"] = kind; d[f"c{i}_pop"] = pop
        d[f"c{i}_blk"] = blk    # logical result bits [n_blocks][2048] (npz-compressed)
        glen = np.where(kind == bm.BLK_GAP, (gaps[:, 0] >> 3) + 1, 0)
        d[f"c{i}_gaps"] = np.concatenate([gaps[c, :glen[c]] for c in range(len(kind))]) if glen.sum() else np.zeros(0, np.uint16)
        d[f"c{i}_any"] = any_
    np.savez_compressed(OUT / f"{name}.npz", **d)
    print(name, "cases", len(cases), "bytes", (OUT / f"{name}.npz").stat().st_size)


def rs_fixture(name, vecs, seed):
    rng = np.random.default_rng(seed)
    ps = bm.PackedSet.pack(vecs)
    d = pack_fields(ps)
    for v in range(ps.n_vec):
        bc, sc, sb, tot = orclib.ref_rs_build(ps, v)
        pos = rng.integers(0, ps.n_blocks * 65536, 2000).astype(np.uint64)
        rank = rng.integers(0, tot + 3, 2000).astype(np.uint64)
        rr, rp, rf, _ = orclib.ref_rank_select(ps, v, pos, rank)
        d[f"v{v}_bcount"] = bc; d[f"v{v}_sub"] = sc; d[f"v{v}_sb"] = sb; d[f"v{v}_total"] = tot
        d[f"v{v}_pos"] = pos; d[f"v{v}_rank_out"] = rr; d[f"v{v}_rank"] = rank; d[f"v{v}_sel_pos"] = rp; d[f"v{v}_sel_found"] = rf
    np.savez_compressed(OUT / f"{name}.npz", **d)
    print(name, "bytes", (OUT / f"{name}.npz").stat().st_size)


def scan_fixture(name, seed, nullable):
    """A real bm::sparse_vector<unsigned>: its own optimize()d planes (+ universe) and bm::sparse_vector_scanner<> answers."""
    import test_oracle_vs_reference as tor
    vals, nulls = tor.scan_inputs(seed, n=100000, nullable=nullable)
    planes = orclib.ref_sv_planes(vals, nulls)
    ps = bm.PackedSet.pack(planes)
    d = pack_fields(ps)
    d["values"] = vals
    d["nulls"] = nulls if nulls is not None else np.zeros(0, np.uint8)
    d["n_cases"] = len(tor.SCAN_CASES)
    for i, (pred, search) in enumerate(tor.SCAN_CASES):
        counts, kind, pop, blk = orclib.ref_sv_scan(vals, nulls, pred, search)
        d[f"c{i}_pred"] = pred; d[f"c{i}_search"] = np.asarray(search, np.uint64); d[f"c{i}_counts"] = counts
        d[f"c{i}_pop"] = pop; d[f"c{i}_blk"] = blk
    np.savez_compressed(OUT / f"{name}.npz", **d)
    print(name, "cases", len(tor.SCAN_CASES), "bytes", (OUT / f"{name}.npz").stat().st_size)


def blob_fixture(name, vecs=None, levels=(0, 1, 2)):
    """BLOBs written by bm::serializer<> at the given compression levels + what bm::deserialize makes of them (kinds, bits, GAP
    words).  "blobs": explicit-length encodings (levels 0..2); "blobs_entropy": gamma / interpolative / super-block encodings
    (levels 3..6) over tests/gen.entropy_vectors."""
    import test_oracle_vs_reference as tor
    vecs = tor.blob_inputs() if vecs is None else vecs
    ps = bm.PackedSet.pack(vecs)
    d = dict(n_vec=ps.n_vec, n_blocks=ps.n_blocks, levels=np.array(levels))
    for level in levels:
        for v in range(ps.n_vec):
            # pseudo-levels 100 + l: level l written with serializer::set_bookmarks(true, 4) (skip marks every 4 blocks)
            blob = orclib.ref_serialize_bookmarks(ps, v, level - 100, 4) if level >= 100 else orclib.ref_serialize(ps, v, level)
            kind, pop, blk, gaps = orclib.ref_deserialize(blob, ps.n_blocks)
            d[f"l{level}_v{v}_blob"] = blob; d[f"l{level}_v{v}_kind"] = kind
            glen = np.where(kind == bm.BLK_GAP, (gaps[:, 0] >> 3) + 1, 0)
            d[f"l{level}_v{v}_gaps"] = np.concatenate([gaps[c, :glen[c]] for c in range(len(kind))]) if glen.sum() else np.zeros(0, np.uint16)
            if level == levels[-1]:
                d[f"v{v}_blk"] = blk
    np.savez_compressed(OUT / f"{name}.npz", **d)
    print(name, "bytes", (OUT / f"{name}.npz").stat().st_size)


if __name__ == "__main__":
    assert orclib.have_ref(), "build oracle/_ref first (make -C oracle)"
    if "blob" in sys.argv[1:] or len(sys.argv) == 1:
        blob_fixture("blobs")
        import test_oracle_vs_reference as tor
        blob_fixture("blobs_entropy", tor.entropy_inputs(), (3, 4, 5, 106, 6))
        if "blob" in sys.argv[1:]:
            sys.exit(0)
    if "scan" in sys.argv[1:] or len(sys.argv) == 1:
        scan_fixture("scan_plain", 31, False)
        scan_fixture("scan_nullable", 32, True)
        if "scan" in sys.argv[1:]:
            sys.exit(0)
    C = bm.F_OPT_COMPRESS
    rng = np.random.default_rng(20260923)
    vecs = gen.mixed_vectors(rng, 12, 6)
    aggregate_fixture("agg_mixed", vecs, [
        (bm.OP_OR, list(range(12)), None, 0), (bm.OP_OR, list(range(12)), None, C), (bm.OP_OR, [3, 7], None, 0),
        (bm.OP_AND, [0, 1], None, 0), (bm.OP_AND, [2, 5, 9], None, C), (bm.OP_AND_SUB, [0], list(range(1, 12)), C),
        (bm.OP_AND_SUB, [4, 6], [1, 2, 3], C), (bm.OP_AND_SUB, [8, 9, 10], [], C), (bm.OP_XOR, [0, 1], None, 0),
        (bm.OP_XOR, [2, 3, 4], None, 0)])
    aggregate_fixture("agg_edge", gen.edge_vectors(4), [
        (bm.OP_OR, [0, 1, 2, 3, 4, 5], None, 0), (bm.OP_OR, [0, 1], None, C), (bm.OP_OR, [5], None, 0), (bm.OP_OR, [3, 4], None, C),
        (bm.OP_AND, [2, 3], None, 0), (bm.OP_AND, [0, 2], None, C), (bm.OP_AND, [2, 2], None, 0),
        (bm.OP_AND_SUB, [2], [0], C), (bm.OP_AND_SUB, [2], [5], C), (bm.OP_AND_SUB, [2, 3], [1, 4], C), (bm.OP_AND_SUB, [1], [], C)])
    # Zipf-like mix, the C3 recipe scaled down: d_k = 0.5/k, optimize()d, AND {1,2} SUB {3..}
    rng = np.random.default_rng(7)
    zv = [bm.BVector.random(3, 0.5 / (k + 1), rng).optimize() for k in range(40)]
    aggregate_fixture("agg_zipf", zv, [(bm.OP_AND_SUB, [0, 1], list(range(2, 40)), C), (bm.OP_OR, list(range(20, 40)), None, C),
                                       (bm.OP_AND, [0, 1, 2], None, C)])
    rng = np.random.default_rng(99)
    rs_fixture("rs_mixed", gen.mixed_vectors(rng, 3, 520, p_null=0.2, p_full=0.1, p_gap=0.4) + gen.edge_vectors(520)[:5], 5)
