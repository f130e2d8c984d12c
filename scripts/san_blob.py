#!/usr/bin/env python
"""Small deserialize-to-device workload for compute-sanitizer: the committed entropy-coded BLOBs (levels 4, 6 and 6 with bookmarks)
decoded on the GPU and compared with the committed bm::deserialize output."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import bitmagic_b200 as bm   # noqa: E402
import golden_util as gu     # noqa: E402

ctx = bm.default_context(0)
nv, nb, blobs, kinds, blks, gapsf = gu.load_blobs("blobs_entropy")
for level in (4, 6, 106):
    dset = bm.DeviceSet.upload_blobs(ctx, blobs[level], nb)
    ps = dset.download()
    for v in range(nv):
        bv = ps.vector(v)
        assert np.array_equal(bv.kind, kinds[level][v])
        assert np.array_equal(np.stack([bv.block_words(c) for c in range(nb)]), blks[v])
    dset.free()
print("san_blob ok")
