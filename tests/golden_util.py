"""Loaders for the committed reference fixtures in tests/golden/ (made by tests/golden/make_golden.py)."""
from pathlib import Path

import numpy as np

import bitmagic_b200 as bm

GOLDEN = Path(__file__).resolve().parent / "golden"


def load_set(z) -> bm.PackedSet:
    return bm.PackedSet(int(z["n_vec"]), int(z["n_blocks"]), np.ascontiguousarray(z["desc"]),
                        np.ascontiguousarray(z["bit_base"]), np.ascontiguousarray(z["gap_base"]),
                        np.ascontiguousarray(z["bit_pool"]), np.ascontiguousarray(z["gap_pool"]))


def load_agg(name):
    z = np.load(GOLDEN / f"{name}.npz")
    ps = load_set(z)
    cases = []
    for i in range(int(z["n_cases"])):
        kind = z[f"c{i}_kind"]
        blk = z[f"c{i}_blk"].reshape(ps.n_blocks, 2048)
        cases.append(dict(op=int(z[f"c{i}_op"]), g0=z[f"c{i}_g0"], g1=z[f"c{i}_g1"], flags=int(z[f"c{i}_flags"]),
                          kind=kind, pop=z[f"c{i}_pop"], blk=blk, gaps=z[f"c{i}_gaps"], any=bool(z[f"c{i}_any"])))
    return ps, cases


def load_rs(name):
    z = np.load(GOLDEN / f"{name}.npz")
    ps = load_set(z)
    vs = []
    for v in range(ps.n_vec):
        vs.append({k: z[f"v{v}_{k}"] for k in ("bcount", "sub", "sb", "total", "pos", "rank_out", "rank", "sel_pos", "sel_found")})
    return ps, vs


def load_scan(name):
    z = np.load(GOLDEN / f"{name}.npz")
    ps = load_set(z)
    cases = [dict(pred=int(z[f"c{i}_pred"]), search=z[f"c{i}_search"], counts=z[f"c{i}_counts"], pop=z[f"c{i}_pop"], blk=z[f"c{i}_blk"])
             for i in range(int(z["n_cases"]))]
    return ps, z["values"], (z["nulls"] if z["nulls"].size else None), cases


def load_blobs(name="blobs"):
    """-> n_vec, n_blocks, {level: [blob per vector]}, {level: [kind per vector]}, [bits per vector], {level: [flat GAP words per vector]}"""
    z = np.load(GOLDEN / f"{name}.npz")
    nv, nb = int(z["n_vec"]), int(z["n_blocks"])
    blobs = {int(l): [z[f"l{int(l)}_v{v}_blob"] for v in range(nv)] for l in z["levels"]}
    kinds = {int(l): [z[f"l{int(l)}_v{v}_kind"] for v in range(nv)] for l in z["levels"]}
    gaps = {int(l): [z[f"l{int(l)}_v{v}_gaps"] for v in range(nv)] for l in z["levels"]}
    return nv, nb, blobs, kinds, [z[f"v{v}_blk"] for v in range(nv)], gaps


def check_agg_case(case, kind, pop, blocks, gaps_flat=None, check_kind=True):
    """Compare one aggregate output (kind[n], pop[n], blocks[n][2048], optional concatenated GAP words) with the fixture."""
    assert np.array_equal(blocks, case["blk"]), "result bits differ from the reference"
    assert np.array_equal(pop, case["pop"]), "popcounts differ from the reference"
    if check_kind:
        assert np.array_equal(kind, case["kind"]), f"block kinds differ: {kind} vs {case['kind']}"
    if gaps_flat is not None:
        assert np.array_equal(gaps_flat, case["gaps"]), "GAP encodings differ from the reference"
