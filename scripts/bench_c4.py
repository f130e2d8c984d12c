"""BASELINE config 4: rs_index build + 10M rank (count_to) + 10M select on ONE 2^32-bit vector, 1% density,
un-optimized (65536 bit-blocks, 512 MiB) and optimize()d (mixed bit/GAP).  GPU numbers = CUDA events with the
vector, the index and the query arrays resident in HBM; CPU numbers = the unmodified reference (BM64ADDR build,
oracle/_ref/libbmref64.so), 1 thread, same vector, same queries; answers compared for ALL queries."""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
import bitmagic_b200 as bm
import orclib

NQ = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
NB = 65536
torch.cuda.set_device(0)
ctx = bm.Context(0)
stream = torch.cuda.current_stream()
ctx.set_stream(stream.cuda_stream)
out = {}
for label, optimize in (("bit_blocks", False), ("optimized", True)):
    dset = bm.DeviceSet.synth(ctx, 1, NB, np.array([0.01]), np.array([7], np.uint64), optimize)
    ps = None
    # --- build ---
    rs = bm.DeviceRS(ctx, dset, 0); ctx.sync()
    for _ in range(3):
        rs.rebuild()
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    reps = 10
    evs[0].record(stream)
    for _ in range(reps):
        rs.rebuild()                       # kernels only (rs_block + two scans); buffers already allocated
    evs[1].record(stream); torch.cuda.synchronize()
    build_ms = evs[0].elapsed_time(evs[1]) / reps
    total = rs.total()
    rng = np.random.default_rng(8)
    pos = rng.integers(0, NB * 65536, NQ, dtype=np.uint64)
    rank = rng.integers(1, total + 1, NQ, dtype=np.uint64)
    d_pos = torch.from_numpy(pos.view(np.int64)).cuda(); d_rank = torch.from_numpy(rank.view(np.int64)).cuda()
    d_out = torch.empty(NQ, dtype=torch.int64, device="cuda"); d_sel = torch.empty(NQ, dtype=torch.int64, device="cuda")
    d_found = torch.empty(NQ, dtype=torch.uint8, device="cuda")
    def timed(fn, reps=3):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps): fn()
        b.record(stream); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    rank_ms = timed(lambda: rs.rank_dev(d_pos.data_ptr(), NQ, d_out.data_ptr()))
    sel_ms = timed(lambda: rs.select_dev(d_rank.data_ptr(), NQ, d_sel.data_ptr(), d_found.data_ptr()))
    g_rank = d_out.cpu().numpy().view(np.uint64); g_sel = d_sel.cpu().numpy().view(np.uint64); g_found = d_found.cpu().numpy().astype(bool)
    kinds = {"bit": dset.n_bit_blocks, "gap_units16B": dset.n_gap_units}
    stored = dset.stored_bytes()
    res = {"blocks": NB, "kinds": kinds, "stored_bytes": stored, "bits_set": int(total),
           "gpu": {"rs_build_ms": build_ms, "rs_build_blocks_per_s": NB / (build_ms * 1e-3), "rs_build_GBps": stored / (build_ms * 1e-3) / 1e9,
                   "rank_ms": rank_ms, "rank_Mq_per_s": NQ / rank_ms / 1e3, "select_ms": sel_ms, "select_Mq_per_s": NQ / sel_ms / 1e3}}
    if orclib.have_ref(True):
        ps = dset.download()
        t0 = time.time()
        r_rank, r_sel, r_found, (tb, tr, ts) = orclib.ref_rank_select(ps, 0, pos, rank, addr64=True)
        res["cpu_reference_1thread"] = {"rs_build_ms": tb * 1e3, "rank_Mq_per_s": NQ / tr / 1e6, "select_Mq_per_s": NQ / ts / 1e6,
                                        "simd": "avx2", "addr": "BM64ADDR"}
        res["parity"] = {"rank_equal": bool(np.array_equal(g_rank, r_rank)), "select_found_equal": bool(np.array_equal(g_found, r_found)),
                         "select_pos_equal": bool(np.array_equal(g_sel[g_found], r_sel[r_found])), "queries": NQ}
        bc, sc, sb = rs.export()
        rbc, rsc, rsb, rtot = orclib.ref_rs_build(ps, 0, addr64=True)
        nz = rbc > 0
        res["parity"]["index_fields_equal"] = bool(np.array_equal(bc, rbc) and np.array_equal(sc[nz], rsc[nz]) and np.array_equal(sb, rsb) and rtot == total)
        res["speedup_vs_1thread"] = {"build": tb * 1e3 / build_ms, "rank": (NQ / rank_ms / 1e3) / (NQ / tr / 1e6), "select": (NQ / sel_ms / 1e3) / (NQ / ts / 1e6)}
    # sector accounting: 32-byte sectors a query touches by construction of the index (sb_cum is a 2 KB table shared by all
    # queries = cache-resident, not counted): rank = row_cum + descriptor + fine entry + <= 2 sectors of the block's 64-byte window;
    # select = 2 (row pivots) + 2 (row entries) + descriptor + 1 (fine pivots) + 2 (fine entries) + 2 (window)
    res["sectors_per_query_by_construction"] = {"rank": 5, "select": 10, "round1_anchor_scheme": {"rank": "~15 (3 index + ~11 of a <= 170-word scan)", "select": "~29"}}
    out[label] = res
    rs.free(); dset.free()
    print(label, json.dumps(res), flush=True)
# the random-access bound of this GPU: independent random 32-byte sector reads over a 512 MiB footprint (scripts/microbench_sector.cu)
bound = None
try:
    import subprocess
    exe = ROOT / "scripts" / "_bin" / "microbench_sector"
    if exe.exists():
        lines = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120).stdout.strip().splitlines()
        bound = [json.loads(x) for x in lines]
        gs = bound[0]["Gsectors_per_s"]
        for label in out:
            g = out[label]["gpu"]
            out[label]["fraction_of_sector_bound"] = {"rank": g["rank_Mq_per_s"] / 1e3 * 5 / gs, "select": g["select_Mq_per_s"] / 1e3 * 10 / gs,
                                                      "bound_Gsectors_per_s": gs}
except Exception as e:       # the microbenchmark is optional evidence
    bound = str(e)
print(json.dumps({"workload": "c4: rs_index build + rank/select on one 2^32-bit vector, 1% density", "n_queries": NQ, "results": out,
                  "random_sector_bound": bound}))
