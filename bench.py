#!/usr/bin/env python
"""bench.py -- headline benchmark for the block-level aggregation hot path.

A "step" is ONE pass of the aggregator over one synthetic vector set:
  workload c3 (default; the config BASELINE.json's target is quoted on, configs[2]):
      bm::aggregator::combine_and_sub over 1024 vectors x 2^30 bits, Zipf density mix d_k = 0.5/k,
      every vector optimize()d (k <~ 51 bit-blocks, the rest GAP), AND = {1,2}, SUB = {3..1024}, opt_compress
  workload c2 (configs[1]): combine_or over 256 vectors x 2^28 bits, 5 % density, bit-blocks only
Multi-GPU (torchrun, one rank per GPU): the block range is sharded -- every rank owns a contiguous range of
block columns of every vector (weak scaling: a full-size shard per rank), aggregates it locally and the ranks
exchange per-block popcounts with one NCCL all_gather + one all_reduce of the cardinality on the same stream.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c3|c2]

Prints ONE JSON line (see README / DESIGN.md section "Measurement").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

WORKLOADS = {
    "c3": dict(n_vec=1024, n_blocks=16384, op="and_sub", desc="combine_and_sub 1024 x 2^30 bits, Zipf d_k=0.5/k, optimize()d (bit + GAP), AND={1,2} SUB={3..1024}, opt_compress"),
    "c2": dict(n_vec=256, n_blocks=4096, op="or", desc="combine_or 256 x 2^28 bits, iid 5% density, bit-blocks"),
    # configs[4]: 4096 x 2^32 bits block-range sharded over 8 GPUs = 8192 block columns per GPU; density is not fixed by
    # BASELINE (as bit-blocks it would be 2 TiB), SURVEY 8d proposes iid p=0.0025 + optimize() => GAP blocks, ~22 GB per GPU
    "c5": dict(n_vec=4096, n_blocks=8192, op="or", desc="combine_or 4096 vectors, 8192 block columns per GPU (2^32 bits over 8 GPUs), iid 0.25% density, optimize()d (GAP blocks)"),
}


def workload_inputs(name: str, rank: int):
    w = WORKLOADS[name]
    nv = w["n_vec"]
    if name == "c3":
        dens = np.array([0.5 / (k + 1) for k in range(nv)])
        seed = np.arange(1000, 1000 + nv, dtype=np.uint64) + np.uint64(1_000_003 * rank)
        optimize = True
    elif name == "c5":
        dens = np.full(nv, 0.0025)
        seed = np.arange(5000, 5000 + nv, dtype=np.uint64) + np.uint64(1_000_003 * rank)
        optimize = True
    else:
        dens = np.full(nv, 0.05)
        seed = np.arange(100, 100 + nv, dtype=np.uint64) + np.uint64(1_000_003 * rank)
        optimize = False
    return dens, seed, optimize


def workload_groups(name: str):
    import bitmagic_b200 as bm
    nv = WORKLOADS[name]["n_vec"]
    if name == "c3":
        return bm.OP_AND_SUB, np.array([0, 1], np.uint32), np.arange(2, nv, dtype=np.uint32), bm.F_OPT_COMPRESS
    if name == "c5":
        return bm.OP_OR, np.arange(nv, dtype=np.uint32), None, bm.F_OPT_COMPRESS
    return bm.OP_OR, np.arange(nv, dtype=np.uint32), None, bm.F_OPT_NONE


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines: list[str] = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def profiled_traffic(workload: str):
    """dram__bytes_read+write per launch of the dominant kernel from the committed ncu capture (profiles/), or None."""
    f = ROOT / "profiles" / "r01" / "ncu_agg_kernel.json"
    key = {"c3": "c3_agg_kernel_and_sub", "c2": "c2_agg_kernel_or", "c5": "c5_agg_kernel_or"}.get(workload)
    try:
        d = json.loads(f.read_text())[key]
        def gb(x):
            return float(x["value"]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[x["unit"]]
        return int(gb(d["dram__bytes_read.sum"]) + gb(d["dram__bytes_write.sum"]))
    except Exception:
        return None


def set_stats(ps_kinds_counts, gap_words_exact, result_bytes, n_cols):
    """Algorithmic bytes: stored size of every source block (bit 8192 B, GAP 2*(len+1) B,
    FULL/NULL 0) + the result blocks actually written (8192 B per bit-block, 2*(len+1) B per GAP block -- the
    bit->GAP step is fused into the kernel) + 12 B of popcount/digest per block column."""
    n_bit = ps_kinds_counts["bit"]
    return n_bit * 8192 + gap_words_exact * 2 + result_bytes + n_cols * 12


def device_set_stats(ctx, dset, torch):
    """Count source blocks by kind and the exact GAP payload with a few torch ops over the device arrays."""
    ptrs = dset.device_ptrs()
    n = dset.n_vec * dset.n_blocks

    class Wrap:
        def __init__(self, addr, nbytes, typestr, shape):
            self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (addr, False), "version": 2}
    desc = torch.as_tensor(Wrap(ptrs.desc, n * 4, "<i4", (n,)), device=f"cuda:{ctx.device}").long() & 0xFFFFFFFF
    kind = desc & 3
    counts = {"null": int((kind == 0).sum()), "full": int((kind == 1).sum()), "bit": int((kind == 2).sum()), "gap": int((kind == 3).sum())}
    gap_words = 0
    if counts["gap"]:
        gp = torch.as_tensor(Wrap(ptrs.gap_pool, dset.n_gap_units * 16, "<i2", (dset.n_gap_units * 8,)), device=f"cuda:{ctx.device}")
        gb = torch.as_tensor(Wrap(ptrs.gap_base, (dset.n_blocks + 1) * 8, "<i8", (dset.n_blocks + 1,)), device=f"cuda:{ctx.device}")
        col = torch.arange(dset.n_blocks, device=desc.device).repeat_interleave(dset.n_vec)
        isgap = kind == 3
        dg = desc[isgap]
        off = (gb[col[isgap]] + ((dg >> 2) & 0x0FFFFFFF)) * 8 + (dg >> 31)      # + lead pad (BMB200_DESC_GAP_PAD)
        hdr = gp[off].long() & 0xFFFF
        gap_words = int(((hdr >> 3) + 1).sum())
    return counts, gap_words


def run_reference(args):
    """--impl reference: the unmodified reference (oracle/_ref/libbmref.so; else the C oracle port) on the host
    cores, all threads, on a bounded column sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    import torch
    import bitmagic_b200 as bm
    import orclib
    w = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, args.ref_threads or cores))
    # every worker gets whole 256-block superblocks (the reference walks all 256 sub-blocks of a top block
    # once a group has > 32 vectors, src/bmaggregator.h:1565-1566), so threads <= n_blocks / 256
    threads = max(1, min(threads, w["n_blocks"] // 256))
    sample_cols = min(w["n_blocks"], args.ref_cols or 256 * threads)
    dens, seed, optimize = workload_inputs(args.workload, 0)
    op, g0, g1, flags = workload_groups(args.workload)
    ctx = bm.default_context(0)
    dset = bm.DeviceSet.synth(ctx, w["n_vec"], sample_cols, dens, seed, optimize)     # same generator, first columns
    ps = dset.download()
    kinds = ps.kinds()
    n_src_blocks = int((kinds != 0).sum())
    dset.free()
    kind = "reference" if orclib.have_ref() else "port"
    times = []
    for it in range(args.warmup + args.steps):
        if kind == "reference":
            sec, tot = orclib.ref_time_aggregate(ps, op, g0, g1, flags, threads=threads, repeats=1)
        else:
            t0 = time.perf_counter(); orclib.oracle_aggregate(ps, op, g0, g1, flags); sec = time.perf_counter() - t0
            threads = 1
        if it >= args.warmup:
            times.append(sec)
    ms = 1e3 * float(np.mean(times))
    value = n_src_blocks / (ms * 1e-3)
    sample = f"{sample_cols} of {w['n_blocks']} block columns x {w['n_vec']} vectors per step ({ps.stored_bytes() / 2**20:.0f} MiB)"
    line = {"impl": "reference", "metric": "aggregator input 64Kbit-blocks/s", "value": value, "unit": "blocks/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": args.workload + ": " + w["desc"], "sample": sample},
            "cpu_baseline": {"value": value, "unit": "blocks/s", "cores": threads, "kind": kind, "sample": sample,
                             "simd": orclib.ref().ref_simd().decode() if kind == "reference" else "scalar"},
            "e2e": {"value": value, "unit": "blocks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    return line


class _StdoutToStderr:
    """Everything libraries print while the bench runs (NCCL's version banner, warnings ...) goes to stderr, so that stdout
    carries exactly ONE line: the JSON result."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        os.close(self._saved)
        return False


def main():
    with _StdoutToStderr():
        line = _main()
    if line is not None:
        print(json.dumps(line), flush=True)


def _main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=list(WORKLOADS))
    ap.add_argument("--cols", type=int, default=0, help="override block columns per GPU (reduced runs are flagged)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-cols", type=int, default=256, help="block columns in the cpu_baseline sample")
    ap.add_argument("--ref-cols", type=int, default=0, help="block columns per step for --impl reference (0 = 256 per thread)")
    ap.add_argument("--ref-threads", type=int, default=0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import bitmagic_b200 as bm

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = torch.device(f"cuda:{local}")

    w = WORKLOADS[args.workload]
    n_cols = args.cols or w["n_blocks"]
    ctx = bm.Context(local)
    stream = torch.cuda.current_stream(dev)
    ctx.set_stream(stream.cuda_stream)

    dens, seed, optimize = workload_inputs(args.workload, rank)
    op, g0, g1, flags = workload_groups(args.workload)
    t0 = time.time()
    dset = bm.DeviceSet.synth(ctx, w["n_vec"], n_cols, dens, seed, optimize)
    ctx.sync()
    t_synth = time.time() - t0
    counts, gap_words = device_set_stats(ctx, dset, torch)
    n_src_blocks = counts["bit"] + counts["gap"] + counts["full"]

    res = bm.aggregate(ctx, dset, op, g0, g1, flags)     # allocates the result buffers once
    ctx.sync()
    rp = res.device_ptrs()

    class Wrap:
        def __init__(self, addr, typestr, shape):
            self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (addr, False), "version": 2}
    pop_t = torch.as_tensor(Wrap(rp["popcnt"], "<i4", (n_cols,)), device=dev)
    gathered = torch.empty(world * n_cols, dtype=torch.int32, device=dev) if world > 1 else None
    card = torch.zeros(1, dtype=torch.int64, device=dev)

    from bitmagic_b200.sharding import exchange_popcounts

    def step():
        bm.aggregate(ctx, dset, op, g0, g1, flags, result=res)
        if world > 1:                                          # per-block popcounts of every shard + global cardinality
            exchange_popcounts(pop_t, world * n_cols, dist, out=gathered)

    l0 = ctx.launch_count()
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize(dev)
    launches_per_step = (ctx.launch_count() - l0) // args.warmup

    # ---- timed region: device-resident inputs, CUDA events on the launching stream ----
    sampler = ClockSampler(local) if rank == 0 else None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    if sampler:
        sampler.start()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    # a dedicated event pair around the dominant kernel of every step (aggregate launch only)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    evs[0].record(stream)
    for i in range(args.steps):
        kev[i][0].record(stream)
        bm.aggregate(ctx, dset, op, g0, g1, flags, result=res)
        kev[i][1].record(stream)
        if world > 1:
            _, card = exchange_popcounts(pop_t, world * n_cols, dist, out=gathered)
        evs[i + 1].record(stream)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    total_ms = evs[0].elapsed_time(evs[-1])
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps

    total_bits, any_ = res.total()
    kind_r, pop_r, dig_r, nr_r = res.meta()
    res_bytes = int((kind_r == bm.BLK_BIT).sum()) * 8192 + int(2 * (nr_r[kind_r == bm.BLK_GAP].astype(np.int64) + 1).sum())
    alg_bytes = set_stats(counts, gap_words, res_bytes, n_cols)

    # whole-job numbers
    src_blocks_all = torch.tensor([n_src_blocks], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(src_blocks_all)
    value = float(src_blocks_all.item()) / (ms_per_step * 1e-3)

    # ---- e2e through the host C-ABI call: pinned HOST buffers -> H2D -> kernel -> D2H of the metadata ----
    e2e = None
    if not args.no_e2e:
        ps = dset.download()
        for a in (ps.desc, ps.bit_base, ps.gap_base, ps.bit_pool, ps.gap_pool):
            if a.size:
                torch.cuda.cudart().cudaHostRegister(a.ctypes.data, a.nbytes, 0)
        h2d = ps.desc.nbytes + ps.bit_base.nbytes + ps.gap_base.nbytes + ps.bit_pool.nbytes + ps.gap_pool.nbytes + (g0.size + (g1.size if g1 is not None else 0)) * 4
        d2h = n_cols * (1 + 4 + 8 + 4) + 8
        bm.aggregate_host(ctx, ps, op, g0, g1, flags)        # warm-up
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            k2, p2, d2, n2, tot2 = bm.aggregate_host(ctx, ps, op, g0, g1, flags)
        torch.cuda.synchronize(dev)
        e_ms = (time.perf_counter() - t0) * 1e3 / args.e2e_steps
        te = torch.tensor([e_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e_ms = float(te.item())
        assert tot2 == total_bits and np.array_equal(p2, pop_r), "e2e result differs from the device-resident run"
        e2e = {"value": float(src_blocks_all.item()) / (e_ms * 1e-3), "unit": "blocks/s", "ms_per_step": e_ms,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "h2d_gbs": h2d / e_ms / 1e6}
        for a in (ps.desc, ps.bit_base, ps.gap_base, ps.bit_pool, ps.gap_pool):
            if a.size:
                torch.cuda.cudart().cudaHostUnregister(a.ctypes.data)
        del ps

    # ---- CPU baseline (rank 0, N=1 only): the reference on a bounded sample of the same workload ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import orclib
        ncs = min(args.cpu_cols, n_cols)
        ps_s = dset.download(0, ncs)
        src_s = int((ps_s.kinds() != 0).sum())
        if orclib.have_ref():
            sec, tot = orclib.ref_time_aggregate(ps_s, op, g0, g1, flags, threads=1, repeats=2)
            kind_c, simd = "reference", orclib.ref().ref_simd().decode()
        else:
            t0 = time.perf_counter(); o = orclib.oracle_aggregate(ps_s, op, g0, g1, flags); sec = time.perf_counter() - t0
            tot = int(o[1].sum()); kind_c, simd = "port", "scalar"
        assert tot == int(pop_r[:ncs].sum()), "CPU baseline and GPU disagree on the sampled columns"
        cpu = {"value": src_s / sec, "unit": "blocks/s", "cores": 1, "kind": kind_c, "simd": simd,
               "sample": f"first {ncs} of {n_cols} block columns x {w['n_vec']} vectors ({ps_s.stored_bytes() / 2**20:.0f} MiB), {sec * 1e3:.0f} ms",
               "gbs": ps_s.stored_bytes() / sec / 1e9, "checked_equal_popcount": True}

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
        line = {
            "metric": "aggregator input 64Kbit-blocks/s", "value": value, "unit": "blocks/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": args.workload + ": " + w["desc"], "n_vec": w["n_vec"], "block_columns_per_gpu": n_cols,
                       "reduced": bool(args.cols and args.cols != w["n_blocks"]),
                       "l2": f"inputs ({alg_bytes / 2**30:.2f} GiB per GPU) larger than the 126 MB L2; no flush needed",
                       "parallelism": f"block-range sharded x{world}", "source_blocks": counts, "synth_s": round(t_synth, 2)},
            "gbs_per_gpu": alg_bytes / (ms_per_step * 1e-3) / 1e9,
            "clocks": clocks, "gpu_launches": int(launches_per_step * args.steps),
            "e2e": e2e,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None if (args.cols and args.cols != w["n_blocks"]) else profiled_traffic(args.workload),
                         "traffic_source": "profiles/r01/ncu_agg_kernel.json (ncu --set full, same command, full-size shard)",
                         "kernel": "agg_kernel<%s>" % w["op"], "kernel_ms": kern_ms,
                         "algorithmic_bytes_per_launch": int(alg_bytes), "peak_source": peak_src},
            "cpu_baseline": cpu,
            "result_bits": int(total_bits),
        }
    else:
        line = None
    if world > 1:
        dist.destroy_process_group()
    return line


if __name__ == "__main__":
    main()
