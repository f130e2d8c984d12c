#!/usr/bin/env python
"""Scanner row: bmb200_scan vs bm::sparse_vector_scanner<> on the same sparse vector.
  python scripts/bench_scan.py [n_elements]     -> one JSON line
Workloads: "dense" = uniform 20-bit values (bit-block planes), "sparse" = 99.6 % zeros (GAP planes).
The GPU answers all search values in ONE launch (values of a column adjacent -> planes from L2); the reference answers
them one by one on one host core.  Algorithmic bytes of a scan = stored bytes of the planes + universe, once per value."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch                 # noqa: E402
import bitmagic_b200 as bm   # noqa: E402
import orclib                # noqa: E402


def run(name, vals, ctx, preds, n_search=64):
    planes = orclib.ref_sv_planes(vals, None)
    ps = bm.PackedSet.pack(planes)
    npl = len(planes) - 1
    dset = bm.DeviceSet.upload(ctx, ps)
    rng = np.random.default_rng(3)
    out = {"elements": int(vals.size), "planes": npl, "block_columns": ps.n_blocks, "stored_MiB": ps.stored_bytes() / 2**20,
           "kinds": {k: int((ps.kinds() == v).sum()) for k, v in (("null", 0), ("full", 1), ("bit", 2), ("gap", 3))}}
    for pname, pred in preds:
        search = rng.choice(vals, n_search).astype(np.uint64)
        if pred == bm.SCAN_RANGE:
            search = np.stack([search, search + 1000], 1)
        res = bm.scan(ctx, dset, pred, search, 0, npl, npl, bm.F_OPT_COMPRESS)
        ctx.sync()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        ev0.record(torch.cuda.current_stream())
        for _ in range(reps):
            bm.scan(ctx, dset, pred, search, 0, npl, npl, bm.F_OPT_COMPRESS, result=res)
        ev1.record(torch.cuda.current_stream())
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / reps
        tot_gpu = int(res.group_totals(n_search).sum())
        res.free()
        ref_s = search.astype(np.uint32)[:16]
        sec, tot_ref = orclib.ref_sv_time_scan(vals, None, pred, ref_s, repeats=1)
        r16 = bm.scan(ctx, dset, pred, search[:16], 0, npl, npl, bm.F_COUNT_ONLY)
        same = int(r16.group_totals(16).sum()) == tot_ref
        r16.free()
        out[pname] = {"gpu_ms_per_launch": ms, "searches_per_launch": n_search, "gpu_searches_per_s": n_search / (ms * 1e-3),
                      "gpu_plane_GBps": ps.stored_bytes() * n_search / (ms * 1e-3) / 1e9, "result_bits": tot_gpu,
                      "ref_ms_per_search_1core": sec * 1e3 / 16, "ref_searches_per_s_1core": 16 / sec, "counts_equal_first16": bool(same),
                      "speedup_vs_1core": (n_search / (ms * 1e-3)) / (16 / sec)}
    dset.free()
    return out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
    ctx = bm.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    rng = np.random.default_rng(11)
    preds = [("find_eq", bm.SCAN_EQ), ("find_gt", bm.SCAN_GT), ("find_range", bm.SCAN_RANGE)]
    dense = rng.integers(0, 1 << 20, n).astype(np.uint32)
    sparse = np.where(rng.random(n) < 0.004, rng.integers(1, 1 << 16, n), 0).astype(np.uint32)
    t0 = time.time()
    res = {"dense20": run("dense20", dense, ctx, preds), "sparse16": run("sparse16", sparse, ctx, preds)}
    print(json.dumps({"bench": "scan", "n": n, "wall_s": round(time.time() - t0, 1), "results": res}))


if __name__ == "__main__":
    main()
