"""Times the C3 set with bit-only / GAP-only SUB groups to see how the two phases compose."""
import sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
import bitmagic_b200 as bm

def timeit(ctx, dset, op, g0, g1, flags, steps=10, warm=3):
    stream = torch.cuda.current_stream()
    res = bm.aggregate(ctx, dset, op, g0, g1, flags)
    for _ in range(warm):
        bm.aggregate(ctx, dset, op, g0, g1, flags, result=res)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(steps):
        bm.aggregate(ctx, dset, op, g0, g1, flags, result=res)
    b.record(stream); torch.cuda.synchronize()
    t = a.elapsed_time(b) / steps
    tot = res.total()[0]
    res.free()
    return t, tot

torch.cuda.set_device(0)
ctx = bm.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
nv, nb = 1024, 16384
dens = np.array([0.5 / (k + 1) for k in range(nv)]); seed = np.arange(1000, 1000 + nv, dtype=np.uint64)
dset = bm.DeviceSet.synth(ctx, nv, nb, dens, seed, True)
C = bm.F_OPT_COMPRESS
kinds = None
cases = {
  "full":      ([0, 1], list(range(2, nv))),
  "bit_only":  ([0, 1], list(range(2, 51))),
  "gap_only":  ([0, 1], list(range(60, nv))),
  "gap_dense": ([0, 1], list(range(60, 200))),
  "gap_sparse":([0, 1], list(range(500, nv))),
}
for ctas in (2, 1):
    ctx.set_tuning(1, ctas)
    for name, (g0, g1) in cases.items():
        t, tot = timeit(ctx, dset, bm.OP_AND_SUB, g0, g1, C)
        print(f"ctas/SM={ctas} {name:10s} {t:7.3f} ms  bits={tot}", flush=True)
ctx.set_tuning(1, 2)
t, tot = timeit(ctx, dset, bm.OP_OR, list(range(60, nv)), None, 0)
print(f"OR over GAP vectors 60..1023: {t:7.3f} ms bits={tot}")
t, tot = timeit(ctx, dset, bm.OP_AND_SUB, [0, 1], list(range(2, nv)), bm.F_COUNT_ONLY)
print(f"full, count-only: {t:7.3f} ms")
