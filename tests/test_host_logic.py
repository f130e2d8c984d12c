"""Host-side containers (numpy only): packing, GAP encode/decode, compare; checked against the oracle."""
import numpy as np
import pytest

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


def _vec_blocks_c(vectors, n_blocks):
    """bmb200_vec_blocks array for hostfmt.BVectors (what DeviceSet.upload_vectors builds)."""
    import ctypes as C
    from bitmagic_b200.capi import VecBlocksC, ptr
    arr = (VecBlocksC * len(vectors))()
    keep = []
    for i, v in enumerate(vectors):
        kind = np.ascontiguousarray(v.kind, dtype=np.uint8)
        ptrs = np.zeros(v.n_blocks, dtype=np.uint64)
        for nb in range(v.n_blocks):
            if kind[nb] in (bm.BLK_BIT, bm.BLK_GAP):
                blk = v.blocks[nb]; keep.append(blk); ptrs[nb] = blk.ctypes.data
        keep += [kind, ptrs]
        arr[i] = VecBlocksC(v.n_blocks, ptr(kind), ptr(ptrs))
    return arr, keep


def test_slab_upload_source_table_host_build():
    """bmb200_set_upload_slabs, host half (csrc/host_pack.hpp pack_sources via oracle/host_pack_check.cpp): blocks laid into a few
    64-byte aligned host slabs; the source table must point at every block's bytes inside the mirror (slabs sorted by address,
    256-byte steps), and a block outside the slabs / a misaligned one must be reported (the product then packs on the host)."""
    import ctypes as C
    import subprocess
    from bitmagic_b200.capi import HostSlabC, VecBlocksC, ptr
    so = orclib.ORACLE_DIR / "libhostpack.so"
    subprocess.run(["make", "-C", str(orclib.ORACLE_DIR), str(so)], check=True, capture_output=True)
    lib = C.CDLL(str(so))
    rng = np.random.default_rng(3)
    vecs = gen.mixed_vectors(rng, 11, 17, p_null=0.1)
    nv, nb = len(vecs), 17
    raw = [np.zeros(200_000 + 64, np.uint8) for _ in range(40)]
    slabs = []
    for r in raw:
        off = (-r.ctypes.data) % 64
        slabs.append([r.ctypes.data + off, r[off:off + 200_000], 0])
    cur = 0
    arr = (VecBlocksC * nv)(); keep = []; where = {}
    for i, v in enumerate(vecs):
        kind = np.ascontiguousarray(v.kind, dtype=np.uint8); ptrs = np.zeros(nb, np.uint64)
        for c in range(v.n_blocks):
            if kind[c] in (bm.BLK_BIT, bm.BLK_GAP):
                b = np.ascontiguousarray(v.blocks[c]).view(np.uint8); need = (b.size + 63) & ~63
                if slabs[cur][1].size - slabs[cur][2] < need:
                    cur += 1
                sl = slabs[cur]; sl[1][sl[2]:sl[2] + b.size] = b; ptrs[c] = sl[0] + sl[2]; where[(c, i)] = b.copy(); sl[2] += need
        keep += [kind, ptrs]; arr[i] = VecBlocksC(nb, ptr(kind), ptr(ptrs))
    used = [sl for sl in slabs if sl[2]]
    rng.shuffle(used)                                       # the caller's slab order is arbitrary
    carr = (HostSlabC * len(used))()
    for k, sl in enumerate(used):
        carr[k].base = sl[0]; carr[k].bytes = sl[2]
    for threads in (1, 4):
        src = np.zeros(nv * nb, np.uint32); dev_off = np.zeros(len(used), np.uint64); total = C.c_uint64(0)
        assert lib.host_pack_sources(nv, nb, arr, threads, carr, len(used), ptr(src), ptr(dev_off), C.byref(total)) == 0
        mirror = np.zeros(total.value, np.uint8)
        for k, sl in enumerate(used):
            assert int(dev_off[k]) % 256 == 0
            mirror[int(dev_off[k]):int(dev_off[k]) + sl[2]] = sl[1][:sl[2]]
        order = np.argsort([sl[0] for sl in used])
        assert all(dev_off[order[k]] < dev_off[order[k + 1]] for k in range(len(used) - 1)), "mirror follows address order"
        for (c, i), b in where.items():
            o = int(src[c * nv + i]) * 32
            assert np.array_equal(mirror[o:o + b.size], b), f"block ({c},{i})"
    # one block outside every slab / one misaligned pointer: reported, not mis-copied
    kind0 = keep[0]; ptrs0 = keep[1]
    c0 = int(np.flatnonzero((kind0 == bm.BLK_BIT) | (kind0 == bm.BLK_GAP))[0])
    saved = int(ptrs0[c0])
    stray = np.ascontiguousarray(vecs[0].blocks[c0]); ptrs0[c0] = stray.ctypes.data
    src = np.zeros(nv * nb, np.uint32); dev_off = np.zeros(len(used), np.uint64); total = C.c_uint64(0)
    assert lib.host_pack_sources(nv, nb, arr, 2, carr, len(used), ptr(src), ptr(dev_off), C.byref(total)) == 1
    if kind0[c0] == bm.BLK_GAP:
        ptrs0[c0] = saved + 2
        assert lib.host_pack_sources(nv, nb, arr, 2, carr, len(used), ptr(src), ptr(dev_off), C.byref(total)) == 1
    ptrs0[c0] = saved


def test_slab_allocator_under_bvector_traffic():
    """bm::b200::slab_bvector (bmb200_alloc.hpp) next to bm::bvector<> on the same operations, 8 threads: equal contents and
    block kinds, every block inside the heap's slabs and 64-byte aligned, freed blocks reused (oracle/slab_heap_check.cpp)."""
    import subprocess
    exe = orclib.ORACLE_DIR / "_ref" / "slab_heap_check"
    if not exe.exists():
        pytest.skip("oracle/_ref/slab_heap_check not built (needs /root/reference at build time)")
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr


def test_upload_packer_host_build_matches_packedset():
    """The product's upload packer (csrc/host_pack.hpp: threaded layout + staging-ring pipeline), built for the host by
    oracle/host_pack_check.cpp, against hostfmt.PackedSet.pack -- several thread counts and slot sizes (1 chunk .. 1 column per chunk)."""
    import ctypes as C
    import subprocess
    so = orclib.ORACLE_DIR / "libhostpack.so"
    if not so.exists():
        subprocess.run(["make", "-C", str(orclib.ORACLE_DIR), str(so)], check=True, capture_output=True)
    lib = C.CDLL(str(so))
    rng = np.random.default_rng(11)
    vecs = gen.mixed_vectors(rng, 9, 23) + gen.edge_vectors(23)[:3]
    short = bm.BVector(5); short.set_full(1); short.set_bits(4, np.full(2048, 0x0F0F0F0F, np.uint32)); vecs.append(short)   # fewer blocks than the set
    nv, nb = len(vecs), 23
    ref = bm.PackedSet.pack(vecs, nb)
    arr, keep = _vec_blocks_c(vecs, nb)
    for threads, slot in ((1, 1 << 30), (3, 1 << 16), (8, 1), (5, 40000)):
        n_bit, n_gap = C.c_uint64(0), C.c_uint64(0)
        assert lib.host_pack_sizes(nv, nb, arr, threads, C.byref(n_bit), C.byref(n_gap)) == 0
        assert n_bit.value == int(ref.bit_base[-1]) and n_gap.value == int(ref.gap_base[-1])
        desc = np.zeros(nv * nb, np.uint32); bb = np.zeros(nb + 1, np.uint64); gb = np.zeros(nb + 1, np.uint64)
        bits = np.full(n_bit.value * 2048 + 4, 0x77777777, np.uint32); gaps = np.full(n_gap.value * 8 + 8, 0x7777, np.uint16)
        nch = C.c_uint32(0)
        rc = lib.host_pack_check(nv, nb, arr, threads, C.c_uint64(slot), orclib.ptr(desc), orclib.ptr(bb), orclib.ptr(gb),
                                 orclib.ptr(bits), orclib.ptr(gaps), C.byref(nch))
        assert rc == 0
        assert np.array_equal(desc, ref.desc) and np.array_equal(bb, ref.bit_base) and np.array_equal(gb, ref.gap_base)
        assert np.array_equal(bits[:n_bit.value * 2048], ref.bit_pool), f"bit pool differs (threads={threads}, slot={slot})"
        assert np.array_equal(gaps[:n_gap.value * 8], ref.gap_pool), f"gap pool differs (threads={threads}, slot={slot})"
        if slot == 1:
            assert nch.value > 2 * 4          # slot = the widest column: many chunks, the 4-slot ring wraps several times
    # a GAP block whose header claims more than 1280 words is refused, like before
    bad = bm.BVector(1); g = gen.gap_from_runs([65535], 0).copy(); g[0] = (1281 << 3); bad.set_gap(0, g)
    arr2, keep2 = _vec_blocks_c([bad], 1)
    n_bit, n_gap = C.c_uint64(0), C.c_uint64(0)
    assert lib.host_pack_sizes(1, 1, arr2, 2, C.byref(n_bit), C.byref(n_gap)) == bm.capi.ERR_BADARG


def test_binding_result_store_recycles_blocks_in_place():
    """bm::b200::detail::store_result (C++ binding): results of changing shapes stored into ONE target bvector (blocks recycled in
    place, threaded over top-level sub-trees) == the same results stored into fresh bvectors (compare() == 0, calc_stat kinds, size)."""
    import subprocess
    exe = orclib.ORACLE_DIR / "_ref" / "store_result_check"
    if not exe.exists():
        import pytest
        pytest.skip("oracle/_ref/store_result_check not built (needs /root/reference at build time)")
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK:" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
