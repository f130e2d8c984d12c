#!/usr/bin/env python
"""Deserialize-to-device row: bmb200_set_upload_blobs vs bm::deserialize on the same BLOBs.
  python scripts/bench_blob.py [n_vec] [n_blocks] [level] [bookmark_interval]   -> one JSON line
The set is the C3 recipe scaled down (Zipf densities, optimize()d); every vector is serialized by the reference's own
bm::serializer<> at the given compression level (default 2 = explicit-length encodings, host token walk; 6 = the serializer's
default: gamma / interpolative encodings, token walk + entropy decode on the GPU).  GPU time = token walk + H2D of the BLOB
bytes + decode kernels (one call);
the baseline is bm::deserialize of the same BLOBs on one host core + what it would still have to upload (raw blocks)."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import bitmagic_b200 as bm   # noqa: E402
import orclib                # noqa: E402


def main():
    nv = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    nbk = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    level = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    bmi = int(sys.argv[4]) if len(sys.argv) > 4 else 0          # > 0: serializer::set_bookmarks(true, bmi) -- the walk of a vector splits there
    ctx = bm.Context(0)
    dens = np.array([0.5 / (k + 1) for k in range(nv)])
    seed = np.arange(1000, 1000 + nv, dtype=np.uint64)
    dset = bm.DeviceSet.synth(ctx, nv, nbk, dens, seed, True)
    ps = dset.download()
    blobs = [orclib.ref_serialize_bookmarks(ps, v, level, bmi) if bmi else orclib.ref_serialize(ps, v, level) for v in range(nv)]
    blob_bytes = int(sum(b.size for b in blobs))
    t0 = time.perf_counter()
    for b in blobs:
        orclib.ref_deserialize(b, nbk)           # includes the wrapper's export walk; lower bound printed separately below
    t_ref = time.perf_counter() - t0
    d2 = bm.DeviceSet.upload_blobs(ctx, blobs, nbk); ctx.sync(); d2.free()      # warm-up
    reps = 7; t_gpu = 1e30; t_raw = 1e30        # best of `reps`: cudaMalloc / cudaFree of the arena is erratic on a shared box
    for _ in range(reps):
        t0 = time.perf_counter()
        d2 = bm.DeviceSet.upload_blobs(ctx, blobs, nbk); ctx.sync()
        t_gpu = min(t_gpu, time.perf_counter() - t0)
        if _ < reps - 1:
            d2.free()
    for _ in range(reps):
        t0 = time.perf_counter()
        d3 = bm.DeviceSet.upload(ctx, ps); ctx.sync()
        t_raw = min(t_raw, time.perf_counter() - t0)
        d3.free()
    back = d2.download()
    same = all(np.array_equal(back.vector(v).block_words(c), ps.vector(v).block_words(c)) for v in range(0, nv, max(1, nv // 16)) for c in range(nbk))
    g = list(range(nv))
    r1 = bm.aggregate(ctx, d2, bm.OP_AND_SUB, [0, 1], g[2:], bm.F_OPT_COMPRESS); r2 = bm.aggregate(ctx, dset, bm.OP_AND_SUB, [0, 1], g[2:], bm.F_OPT_COMPRESS)
    agg_same = r1.total() == r2.total()
    print(json.dumps({"bench": "blob", "level": level, "bookmark_interval": bmi, "n_vec": nv, "n_blocks": nbk, "stored_bytes": int(ps.stored_bytes()), "blob_bytes": blob_bytes,
                      "upload_blobs_ms": t_gpu * 1e3, "upload_blobs_GBps_of_blob": blob_bytes / t_gpu / 1e9,
                      "upload_blobs_GBps_of_blocks": ps.stored_bytes() / t_gpu / 1e9, "upload_raw_packed_ms": t_raw * 1e3,
                      "ref_deserialize_1core_ms": t_ref * 1e3, "ref_deserialize_GBps_of_blocks": ps.stored_bytes() / t_ref / 1e9,
                      "bits_equal_sampled": bool(same), "aggregate_equal": bool(agg_same)}))


if __name__ == "__main__":
    main()
