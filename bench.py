#!/usr/bin/env python
"""bench.py -- headline benchmark of the hydro-tools hot path on B200.

Metric (BASELINE.json): D8 flow direction + cell-count accumulation, Gcells/s.
A "step" is one pass of the hot path (ht_d8_f32 + ht_acc_local + ht_acc_resolve
+ ht_acc_final) over one synthetic conditioned fractal DEM.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  torchrun --nproc-per-node N ... bench.py --gpus N ...

N = 1 : BASELINE.json configs[1], 10 000 x 10 000 fp32, whole raster on one GPU.
N > 1 : the raster is row-band sharded, 10 000 rows per GPU (weak scaling), halo
        rows and band-boundary flows exchanged through torch.distributed (NCCL).
`value`  : device-resident (DEM already in HBM), CUDA events, max over ranks.
`e2e`    : the same work through hydrotools.cuda.flow_direction_accumulation_arrays
           with pinned HOST buffers, H2D and D2H copies inside the timed region.
`--impl reference` times the CPU oracle (restatement of GRASS r.watershed, which
the reference shells out to; single-threaded like GRASS) on a bounded sample.
"""
import argparse
import json
import os
import sys
import threading
import time

# NCCL prints its version banner on stdout at NCCL_DEBUG=VERSION and WARN; the contract is
# ONE JSON line there.  INFO / TRACE, if somebody asks for them, are left alone.
if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "WARN"):
    os.environ["NCCL_DEBUG"] = "NONE"

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "hydro-tools_b200"))

METRIC = "D8 flow-dir+accumulation Gcells/s"
UNIT = "Gcells/s"
SEED = 20261019
ROWS_PER_GPU = 10000
COLS = 10000
CSX = CSY = 10.0
TERRAIN_N = 32768  # configs[2]: fused slope/aspect/TRI


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """Polls NVML while the timed region runs (nvidia-smi -lms 200 is too coarse
    for a region that lasts tens of milliseconds)."""

    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap",
        0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
        0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
        0x100: "display_clock_setting",
    }

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def __enter__(self):
        if self.nv and not os.environ.get("HT_BENCH_NO_CLOCKS"):
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# --------------------------------------------------------------------------- reference arm
def oracle_sample_gcells(sample_rows, cols, total_rows):
    import oracle
    dem = oracle.synth_dem(SEED, sample_rows, cols, row0=0, total_rows=total_rows)
    t0 = time.perf_counter()
    oracle.rwatershed(dem, None, CSX, CSY, positive_only=True)
    dt = time.perf_counter() - t0
    return sample_rows * cols / dt / 1e9, dt


def run_reference(args):
    """CPU arm: the oracle restatement of r.watershed -s -a (GRASS is
    single-threaded; so is this).  `--steps 1 --warmup 0` times ONE pass over the whole
    configs[1] raster (about a minute, 4 GB); otherwise every step is a bounded row
    sample of it, and the line says so (`extrapolated`, `sample_rows`, config.rows)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total_rows = ROWS_PER_GPU * args.gpus
    full = args.steps == 1 and args.warmup == 0 and args.gpus == 1
    if full:
        rows = total_rows
    else:
        # size the sample so that (steps + warmup) samples take about two minutes
        g, dt = oracle_sample_gcells(500, COLS, total_rows)
        budget = 120.0 / max(args.steps + args.warmup, 1)
        rows = int(max(200, min(2500, 500 * budget / max(dt, 1e-3) * 0.8)))
    for _ in range(args.warmup):
        oracle_sample_gcells(rows, COLS, total_rows)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_sample_gcells(rows, COLS, total_rows)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    value = rows * COLS / dt / 1e9
    sample = (f"the whole {total_rows} x {COLS} synthetic DEM" if full else
              f"first {rows} rows x {COLS} cols of the {total_rows} x {COLS} synthetic DEM per step")
    cfg = workload_config(args.gpus)
    cfg["rows_timed_per_step"] = rows
    cfg["extrapolated"] = not full
    if not full:
        cfg["note"] = (f"CPU arm times a {rows}-row sample per step and reports its per-cell rate; "
                       "A* is O(N log N) with a growing working set, so the sample flatters the CPU")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": cfg, "extrapolated": not full, "sample_rows": rows,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(n):
    return {
        "workload": f"synthetic conditioned fractal DEM {ROWS_PER_GPU * n}x{COLS} fp32, "
                    f"D8 + cell-count accumulation (BASELINE.json configs[1] per GPU)",
        "rows": ROWS_PER_GPU * n, "cols": COLS, "seed": SEED, "cell_size_m": CSX,
        "sharding": "whole raster on one GPU" if n == 1 else f"{n} row bands of {ROWS_PER_GPU} rows",
        "l2": "inputs (400 MB DEM per GPU) exceed the 126 MB L2; no flush between iterations",
        "step": "every kernel, memset and collective of BandEngine.flow_direction_accumulation, "
                "captured once (after the warm-up) and replayed as one CUDA graph launch per step",
    }



# --------------------------------------------------------------------------- correctness inside the run
_DR = (0, -1, -1, -1, 0, 1, 1, 1, 0)
_DC = (0, 1, 0, -1, -1, -1, 0, 1, 1)


def mass_balance(fd, fa, rank, world, total_rows, weight=None, watershed_edges=True):
    """Size-independent check of a sharded accumulation, on the device:
      * acc(c) == w(c) + sum of acc over c's donors for EVERY cell of every band, the donors
        in the neighbour band included (edge rows of direction and accumulation travel);
      * the accumulation of the cells that do not pass flow on adds up to the total weight
        (the cell count when `weight` is None).
    `watershed_edges`: r.watershed semantics (border cells keep their flow); otherwise
    r.flowaccumulation's (a cell passes flow on unless its receiver is outside).
    Returns (ok_everywhere, terminal_sum, expected_sum)."""
    import torch
    import torch.distributed as dist
    dev = fd.device
    rows, cols = fd.shape
    g0 = rank * rows
    dr = torch.tensor(_DR, device=dev)
    dc = torch.tensor(_DC, device=dev)
    cc1 = torch.arange(cols, device=dev).view(1, -1)
    if world > 1:
        edge_fd = torch.stack([fd[0], fd[-1]]).contiguous()
        edge_fa = torch.stack([fa[0], fa[-1]]).contiguous()
        all_fd = torch.empty((world * 2, cols), dtype=fd.dtype, device=dev)
        all_fa = torch.empty((world * 2, cols), dtype=fa.dtype, device=dev)
        dist.all_gather_into_tensor(all_fd, edge_fd)
        dist.all_gather_into_tensor(all_fa, edge_fa)
    inflow = torch.zeros((rows, cols), dtype=torch.float64, device=dev)

    def passes(f, grow):
        f = f.to(torch.int64)
        rr = grow.expand(-1, cols)
        cc = cc1.expand(f.shape[0], -1)
        code = f.clamp(min=0, max=8)
        tr, tc = rr + dr[code], cc + dc[code]
        ok = (f > 0) & (f <= 8) & (tr >= 0) & (tr < total_rows) & (tc >= 0) & (tc < cols)
        if watershed_edges:
            ok &= ~((rr == 0) | (cc == 0) | (rr == total_rows - 1) | (cc == cols - 1))
        return ok, tr - g0, tc

    def scatter(f, a, grow):
        ok, tr, tc = passes(f, grow)
        ok = ok & (tr >= 0) & (tr < rows)
        inflow.view(-1).index_add_(0, (tr * cols + tc)[ok], a[ok])

    term = torch.zeros((), dtype=torch.float64, device=dev)
    CH = 500
    for r0 in range(0, rows, CH):
        r1 = min(rows, r0 + CH)
        grow = torch.arange(g0 + r0, g0 + r1, device=dev).view(-1, 1)
        ok, _, _ = passes(fd[r0:r1], grow)
        term += fa[r0:r1][~ok].sum()
        scatter(fd[r0:r1], fa[r0:r1], grow)
    if rank > 0:
        k = 2 * (rank - 1) + 1
        scatter(all_fd[k:k + 1], all_fa[k:k + 1], torch.tensor([[g0 - 1]], device=dev))
    if rank < world - 1:
        k = 2 * (rank + 1)
        scatter(all_fd[k:k + 1], all_fa[k:k + 1], torch.tensor([[g0 + rows]], device=dev))
    own = 1.0 if weight is None else weight.double()
    expect = inflow + own
    if weight is None:
        good = torch.equal(fa, expect)
    else:  # fp64 report of an exact integer sum: one rounding per value
        good = bool(((fa - expect).abs() <= 1e-9 * expect.abs().clamp(min=1e-300)).all())
    ok_t = torch.tensor([1.0 if good else 0.0], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(rows * cols) if weight is None else float(weight.double().sum())],
                       dtype=torch.float64, device=dev)
    term = term.reshape(1)
    if world > 1:
        dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
        dist.all_reduce(term)
        dist.all_reduce(tot)
    return bool(ok_t.item()), float(term.item()), float(tot.item())


def oracle_parity_small(rank, world, dev):
    """Oracle comparison inside the run (over NCCL when world > 1): D8 + accumulation
    (unsigned and signed), weighted accumulation over the stored direction grid, on a side
    raster small enough for the single-threaded oracle: 700 rows per GPU x 1500 columns,
    with NULL blobs.  Returns mismatch counts (all must be 0)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from hydrotools.cuda import distributed as HD
    rows, cols = 700, 1500
    total = rows * world
    nd = -32767.0
    eng = HD.BandEngine(rank, world, rows, cols, total, CSX, CSY, nd)
    eng.generate_synthetic(5, with_nulls=True)
    eng.flow_direction_accumulation()
    fd, fa = (t.clone() for t in eng.results())
    eng.flow_direction_accumulation(positive_only=False)
    _, fs = (t.clone() for t in eng.results())
    import oracle
    w_all = oracle.synth_weight(5, total, cols)
    eng.load_direction(fd)
    wsum = eng.flow_accumulation(torch.from_numpy(w_all[rank * rows:(rank + 1) * rows]).to(dev)).clone()
    cnt = eng.flow_accumulation(None)

    def gather(t):
        t = t.contiguous()
        if world == 1:
            return t.cpu().numpy()
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        return torch.cat(parts).cpu().numpy()

    got = {k: gather(v) for k, v in (("fd", fd), ("fa", fa), ("fs", fs), ("w", wsum), ("cnt", cnt))}
    out = None
    if rank == 0:
        dem = oracle.synth_dem(5, total, cols, with_nulls=True)
        efd, efa = oracle.rwatershed(dem, nd, CSX, CSY)
        _, efs = oracle.rwatershed(dem, nd, CSX, CSY, positive_only=False)
        ew = oracle.flowacc(efd, w_all)
        ec = oracle.flowacc(efd)

        def neq(a, b):
            return int((~((a == b) | (np.isnan(a) & np.isnan(b)))).sum())
        m = ~np.isnan(ew)
        rel = float(np.max(np.abs(got["w"][m] - ew[m]) / np.maximum(np.abs(ew[m]), 1e-30)))
        out = {"raster": f"{total}x{cols} with NULL blobs, {world} band(s) of {rows} rows",
               "direction_mismatches": int((got["fd"] != efd).sum()),
               "accumulation_mismatches": neq(got["fa"], efa),
               "signed_accumulation_mismatches": neq(got["fs"], efs),
               "stored_direction_count_mismatches": neq(got["cnt"], ec),
               "weighted_max_rel_err": rel, "weighted_tolerance": 1e-5,
               "weighted_nan_mask_equal": bool(np.array_equal(np.isnan(got["w"]), np.isnan(ew))),
               "oracle": "oracle.rwatershed / oracle.flowacc on the whole raster, rank 0"}
        out["ok"] = (out["direction_mismatches"] == 0 and out["accumulation_mismatches"] == 0 and
                     out["signed_accumulation_mismatches"] == 0 and
                     out["stored_direction_count_mismatches"] == 0 and rel <= 1e-5 and
                     out["weighted_nan_mask_equal"])
    del eng
    torch.cuda.empty_cache()
    return out


def weighted_run(rank, world, rows, cols, peak, barrier, dev, steps=3):
    """Row a3 (FlowAccumulation.calculate "sum", watershed.py:357-382): weighted accumulation
    over a STORED direction grid (r.flowaccumulation semantics), weights ~ U(1650, 1700) mm.
    Timed: weight maximum (+ 4-byte all-reduce), stages A / B / C and the band exchange."""
    import torch
    import torch.distributed as dist
    from hydrotools.cuda import distributed as HD, _native as N, _device as D
    total = rows * world
    eng = HD.BandEngine(rank, world, rows, cols, total, CSX, CSY)
    eng.generate_synthetic(SEED)
    eng.flow_direction_accumulation()
    fd = eng.results()[0].clone()
    eng.load_direction(fd)
    w, wld = D.empty_padded(rows, cols, torch.float32, dev, 4)
    N.call("ht_synth_weight_f32", w.data_ptr(), rows, cols, wld, SEED + 1, rank * rows, D.stream_ptr())
    wv = w[:, :cols]
    for _ in range(2):
        acc = eng.flow_accumulation(wv)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        acc = eng.flow_accumulation(wv)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    acc = acc.clone()  # the engine returns a view of its own buffer
    again = eng.flow_accumulation(wv)
    rep = torch.tensor([1.0 if torch.equal(again, acc) else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(rep, op=dist.ReduceOp.MIN)
    ok, term, tot = mass_balance(fd, acc, rank, world, total, weight=wv, watershed_edges=False)
    del eng, acc, again, w
    torch.cuda.empty_cache()
    cells = total * cols
    gbs = rows * cols * 13 / (ms * 1e-3) / 1e9
    return {"workload": f"{total}x{cols} fp32 weights over a stored int8 direction grid, {world} band(s)",
            "ms": ms, "gcells_s": cells / (ms * 1e-3) / 1e9, "algorithmic_bytes_per_cell": 13,
            "achieved": gbs, "peak": peak, "unit": "GB/s per GPU", "frac": gbs / peak,
            "lower_bound_ms": rows * cols * 13 / (peak * 1e9) * 1e3,
            "bit_reproducible": bool(rep.item()),
            "parity": {"acc_equals_weight_plus_donors_everywhere": ok,
                       "terminal_weight": term, "total_weight": tot,
                       "rel_diff": abs(term - tot) / max(abs(tot), 1e-300)}}


def general_path_run(peak, dev, n=10000, sigma=0.3, steps=3):
    """D8 direction of a RAW DEM through the general kernel chain
    (ht_d8_general_f32: priority-flood levels, parallel rounds, depression episodes) --
    the bench's fluvial terrain plus N(0, sigma) metres of noise, rounded to centimetres:
    millions of pits and tie cells.  Followed by the usual accumulation; checked by mass
    balance (the oracle comparison of this path is in the tests, up to 4000 x 4000)."""
    import torch
    from hydrotools.cuda import synth, watershed as W, _native as N
    t, ld = synth.synth_dem(SEED, n, n)
    g = torch.Generator(device=dev)
    g.manual_seed(SEED)
    for r0 in range(0, n, 2000):  # synthetic input, made on the device in slabs
        v = t[r0:r0 + 2000, :n]
        v.add_(torch.randn(v.shape, generator=g, device=dev) * sigma)
        v.mul_(100.0).round_().div_(100.0)
    pits, ties = W._validate(t, ld, n, n, None, dev)
    ms_dir, info = [], None
    for it in range(steps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        packed, pld, info = W.d8_general(t, ld, n, n, None, CSX, CSY)
        torch.cuda.synchronize()
        if it:
            ms_dir.append((time.perf_counter() - t0) * 1e3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    acc, ald, dout, dld = W.accumulate(packed, pld, n, n, N.MODE_COUNT, want_dir=True)
    e1.record()
    torch.cuda.synchronize()
    ok, term, tot = mass_balance(dout[:, :n], acc[:, :n], 0, 1, n)
    ms = sum(ms_dir) / len(ms_dir)
    del t, packed, acc, dout
    torch.cuda.empty_cache()
    return {"workload": f"{n}x{n} fp32 fluvial terrain + N(0, {sigma} m) noise, rounded to cm (raw DEM)",
            "pits": pits, "tie_cells": ties, "direction_ms": ms, "gcells_s": n * n / ms / 1e6,
            "accumulation_ms": e0.elapsed_time(e1), "stats": info,
            "algorithmic_bytes_per_cell": 5, "achieved": n * n * 5 / ms / 1e6, "peak": peak,
            "unit": "GB/s", "frac": n * n * 5 / ms / 1e6 / peak,
            "timing": "host wall clock around the synchronous entry point (its host loops read one "
                      "device counter per round)",
            "parity": {"acc_equals_one_plus_donors_everywhere": ok, "terminal_cells": term, "cells": tot}}


def e2e_cog_run(host_dem, rows, cols, steps=3):
    """The end-to-end step with the on-disk format produced on the device:
    pinned host DEM -> H2D -> validate + D8 + accumulation -> 512 x 512 tiles of direction and
    accumulation deflated in HBM -> D2H of the COMPRESSED tile streams (what to_raster would
    have to compress on the host after 9 B/cell crossed PCIe).  Returns Gcells/s and the bytes
    that travelled.  The container write itself (hydrotools.cuda._cog.write_cog) is a file
    write of those bytes and is left out, like the reference's disk."""
    import torch
    import hydrotools.cuda as hc
    from hydrotools.cuda import _cog
    dev = torch.device("cuda", torch.cuda.current_device())
    sizes = {}

    def step():
        t = host_dem.to(dev, non_blocking=True)
        fd, fa = hc.flow_direction_accumulation_arrays(t, nodata=None, csx=CSX, csy=CSY)
        for name, a, pad in (("direction", fd, -128), ("accumulation", fa, float("nan"))):
            data, off, size = _cog.encode_tiles(a, pad, copy=False)  # lands in pinned host memory
            sizes[name] = int(len(data))
        torch.cuda.synchronize()

    step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    raw = rows * cols * 9
    return {"value": rows * cols / dt / 1e9, "unit": UNIT, "ms_per_step": dt * 1e3,
            "h2d_bytes_per_step": rows * cols * 4, "d2h_bytes_per_step": sum(sizes.values()),
            "compressed_bytes": sizes, "uncompressed_bytes": raw,
            "compression_ratio": raw / max(sum(sizes.values()), 1),
            "api": "flow_direction_accumulation_arrays (CUDA tensors) + _cog.encode_tiles: the compute and "
                   "compression of hydrotools.cuda.flow_direction_accumulation(dem, fd.tif, fa.tif) with "
                   "HYDROTOOLS_CUDA_IO=device, file write excluded"}


def count_run(rank, world, rows, cols, peak, barrier, dev, steps=3):
    """D8 + cell-count accumulation on rows x cols per GPU (BASELINE.json configs[4] when
    8 x 12 500 x 100 000), checked by mass balance on every cell."""
    import torch
    import torch.distributed as dist
    from hydrotools.cuda import distributed as HD
    total = rows * world
    eng = HD.BandEngine(rank, world, rows, cols, total, CSX, CSY)
    eng.generate_synthetic(SEED)
    pits, ties = eng.validate()
    for _ in range(2):
        eng.flow_direction_accumulation()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.flow_direction_accumulation()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    fd, fa = eng.results()
    ok, term, tot = mass_balance(fd, fa, rank, world, total)
    mem = torch.tensor([torch.cuda.max_memory_allocated() / 2**30], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(mem, op=dist.ReduceOp.MAX)
    del eng, fd, fa
    torch.cuda.empty_cache()
    cells = total * cols
    gbs = rows * cols * 14 / (ms * 1e-3) / 1e9
    return {"workload": f"{total}x{cols} fp32 DEM, D8 + cell-count accumulation, {world} band(s) of {rows} rows",
            "validate": {"pits": pits, "ties": ties},
            "ms": ms, "gcells_s": cells / (ms * 1e-3) / 1e9, "algorithmic_bytes_per_cell": 14,
            "achieved": gbs, "peak": peak, "unit": "GB/s per GPU", "frac": gbs / peak,
            "lower_bound_ms": rows * cols * 14 / (peak * 1e9) * 1e3,
            "peak_gib_per_gpu": float(mem.item()),
            "parity": {"acc_equals_one_plus_donors_everywhere": ok, "terminal_cells": term,
                       "cells": tot, "exact": ok and term == tot}}


# --------------------------------------------------------------------------- product arm
def run_product(args):
    import torch
    import torch.distributed as dist
    import numpy as np

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import hydrotools.cuda as hc
    from hydrotools.cuda import _native as N, _device as D, synth
    from hydrotools.cuda import distributed as HD

    peak, peak_kind = measured_peak()
    dev = torch.device("cuda", local_rank)
    rows, cols = ROWS_PER_GPU, COLS
    total_rows = rows * world
    row0 = rank * rows

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    engine = HD.BandEngine(rank, world, rows, cols, total_rows, CSX, CSY, nodata=None)
    engine.generate_synthetic(SEED)
    pits, ties = engine.validate()
    if pits or ties:
        raise SystemExit(f"synthetic DEM is not conditioned: {pits} pits, {ties} ties")

    if world > 1:
        # NCCL sets up its channels lazily on the first collectives of each size; keep that
        # out of the W warm-up steps the caller asked for
        for _ in range(5):
            engine.flow_direction_accumulation()
    # One step = every kernel, memset and collective of engine.flow_direction_accumulation(),
    # captured once and replayed as ONE CUDA graph launch per step (BandEngine.graph): a
    # sharded step is ~40 host calls, and 8 ranks share the box's host cores.
    try:
        step = engine.graph(engine.flow_direction_accumulation)
        step_mode = "cuda graph replay (one launch per step)"
    except Exception as exc:  # capture refused on this driver / NCCL: enqueue launch by launch
        print(f"[bench] graph capture failed ({type(exc).__name__}: {exc}); launching step by step",
              file=sys.stderr)
        def step():
            engine.flow_direction_accumulation()
        step.release = lambda: None
        step_mode = "launch by launch"
    for _ in range(args.warmup):
        step()
    launches0 = engine.launches
    # the NVML poller starts before the barrier: its first queries (slow when 8 ranks ask at
    # once) are not part of the timed region, the later ones sample it every 4 ms
    with ClockSampler(local_rank) as clocks:
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1) / max(args.steps, 1)
    launches = engine.launches - launches0
    step.release()  # the graph holds NCCL state; the results stay in the engine's buffers
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = total_rows * cols / (ms * 1e-3) / 1e9

    # correctness of THIS run's result: acc == 1 + donors on every cell of every band
    # (neighbour-band donors included) and terminal mass == cells
    fd_dev, fa_dev = engine.results()
    mb_ok, mb_term, mb_tot = mass_balance(fd_dev, fa_dev, rank, world, total_rows)
    parity = {"bench_raster": {"acc_equals_one_plus_donors_everywhere": mb_ok,
                               "terminal_cells": mb_term, "cells": mb_tot,
                               "exact": mb_ok and mb_term == mb_tot}}
    del fd_dev, fa_dev

    # ---- end to end through the public API: pinned host in, host out
    numa = D.bind_to_gpu_numa_node(local_rank)  # staging buffers on the GPU's own NUMA node
    host_dem = torch.empty((rows, cols), dtype=torch.float32, pin_memory=True)
    host_dem.copy_(engine.dem_view())
    out_fd = torch.empty((rows, cols), dtype=torch.int8, pin_memory=True)
    out_fa = torch.empty((rows, cols), dtype=torch.float64, pin_memory=True)
    torch.cuda.synchronize()
    if world == 1:
        api = ("hydrotools.cuda.flow_direction_accumulation_arrays (validate + D8 + "
               "accumulation), pinned host buffers")

        def e2e_step():
            hc.flow_direction_accumulation_arrays(
                host_dem.numpy(), nodata=None, csx=CSX, csy=CSY, positive_only=True,
                out_direction=out_fd.numpy(), out_accumulation=out_fa.numpy())
    else:
        api = ("hydrotools.cuda.distributed.BandEngine: load (H2D of the rank's band) + validate "
               "+ flow_direction_accumulation + results to pinned host buffers")

        def e2e_step():
            engine.load(host_dem)
            pits, ties = engine.validate()
            if pits or ties:
                raise SystemExit("DEM is not conditioned")
            engine.flow_direction_accumulation()
            fd_dev, fa_dev = engine.results()
            out_fd.copy_(fd_dev, non_blocking=True)
            out_fa.copy_(fa_dev, non_blocking=True)
            torch.cuda.synchronize()

    n_e2e = max(2, min(args.steps, 5))
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        e2e_step()
    barrier()
    dt = (time.perf_counter() - t0) / n_e2e
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    single_call = None
    if world == 1:
        # The same call for a STREAM of rasters (hydrotools.cuda.FlowPipeline): upload of
        # raster i + 1, kernels of raster i and download of raster i - 1 overlap, every raster
        # still crosses the host link in full (400 MB in, 900 MB out).  The one-call figure
        # above is the latency of a single raster and is kept beside it.
        single_call = {"ms_per_step": dt * 1e3, "value": total_rows * cols / dt / 1e9, "api": api}
        pipe = hc.FlowPipeline(rows, cols, nodata=None, csx=CSX, csy=CSY)
        n_e2e = max(6, min(args.steps, 10))
        for _ in pipe.run([host_dem] * 3):
            pass
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        last = None
        for last in pipe.run([host_dem] * n_e2e):
            pass
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n_e2e
        out_fd.copy_(torch.from_numpy(last[0]))
        out_fa.copy_(torch.from_numpy(last[1]))
        api = ("hydrotools.cuda.FlowPipeline.run over a stream of rasters (validate + D8 + accumulation "
               "per raster; pinned host buffers; transfers of neighbouring rasters overlapped)")
        del pipe, last
    e2e_cog = e2e_cog_run(host_dem, rows, cols) if world == 1 else None
    # what the host link can do with exactly these buffers: H2D of the DEM and D2H of both
    # results at the same time (full duplex), every rank at once -- the floor of a step that
    # has to move 13 B/cell over PCIe
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    tmp_dem = torch.empty((rows, cols), dtype=torch.float32, device=dev)
    fd_dev, fa_dev = engine.results()

    def link_step():
        with torch.cuda.stream(s_in):
            tmp_dem.copy_(host_dem, non_blocking=True)
        with torch.cuda.stream(s_out):
            out_fd.copy_(fd_dev, non_blocking=True)
            out_fa.copy_(fa_dev, non_blocking=True)
        torch.cuda.synchronize()
    link_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        link_step()
    barrier()
    link = (time.perf_counter() - t0) / 3
    t = torch.tensor([link], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    link = float(t.item())
    del tmp_dem
    e2e = {"value": total_rows * cols / dt / 1e9, "unit": UNIT,
           "h2d_bytes_per_step": rows * cols * 4 * world, "d2h_bytes_per_step": rows * cols * 9 * world,
           "ms_per_step": dt * 1e3, "steps": n_e2e, "api": api,
           "note": "bound by PCIe: 13 B/cell cross the host link every step",
           "hostlink_floor_ms": link * 1e3,
           "hostlink_gbs_aggregate": rows * cols * 13 * world / link / 1e9,
           "frac_of_hostlink": link / dt, "numa": numa, "single_call": single_call,
           "device_compressed": e2e_cog}
    # the e2e result must be the same answer
    fd_dev, fa_dev = engine.results()
    assert torch.equal(out_fd.to(dev), fd_dev) and torch.equal(out_fa.to(dev), fa_dev)

    # ---- per-kernel timing (CUDA events on the launching stream) for the roofline
    kernels = engine.time_kernels(iters=max(3, min(args.steps, 10)))
    resolve_rounds = engine.band.resolve_rounds()
    cells = rows * cols
    # D8 4 B read + 1 B written; accumulation 1 B read + 8 B written (+ the
    # final int8 direction codes, 1 B), split over stage A (reads the packed byte) and
    # stage C (reads it again, writes fp64 + int8); the u16 tile-local counts that travel
    # from A to C and the boundary graph are workspace traffic, not algorithmic bytes
    alg_bytes = {"d8_kernel": 5, "acc_local_cnt_kernel": 1, "coarse_resolve_kernel": 0,
                 "acc_final_cnt_kernel": 10}
    for k in kernels:
        b = alg_bytes.get(k["name"], 0) * cells
        k["algorithmic_bytes"] = b
        k["gbs"] = b / (k["ms"] * 1e-3) / 1e9 if k["ms"] > 0 else 0.0
        k["frac"] = k["gbs"] / peak
    dom = max(kernels, key=lambda k: k["ms"])
    traffic = load_traffic(dom["name"])
    roofline = {"bound": "hbm", "kernel": dom["name"], "achieved": dom["gbs"], "peak": peak,
                "peak_kind": f"of {peak_kind} (MEASURED_PEAKS.json hbm_gbs)", "unit": "GB/s",
                "frac": dom["frac"], "traffic": traffic,
                "share_of_step": dom["ms"] / sum(k["ms"] for k in kernels),
                "note": "the accumulation kernels are bound on chip (stage A: shared-memory "
                        "wavefronts of the pointer doubling, 85 % of the pipe), not by HBM; see DESIGN.md 3.3"}
    # the whole step against the same roofline: 14 B/cell (5 D8 + 9 accumulation)
    path_gbs = rows * cols * 14 / (ms * 1e-3) / 1e9
    path_roofline = {"algorithmic_bytes_per_cell": 14, "achieved": path_gbs, "peak": peak,
                     "unit": "GB/s", "frac": path_gbs / peak}

    cpu = None
    # configs[2]: fused slope + aspect + TRI on 32768^2; with N > 1 the same raster in N
    # row bands (strong scaling), halo rows exchanged inside the timed region
    del engine
    torch.cuda.empty_cache()
    parity["oracle"] = oracle_parity_small(rank, world, dev)
    # row a3 on the bench raster (single GPU) / per band
    weighted = weighted_run(rank, world, rows, cols, peak, barrier, dev)
    targets = None
    if world == 8:
        # BASELINE.json configs[4] (north_star target) and configs[3]
        targets = {"configs[4]": count_run(rank, world, 12500, 100000, peak, barrier, dev),
                   "configs[3]": weighted_run(rank, world, 5000, 40000, peak, barrier, dev)}
    stencil = terrain_roofline(peak) if world == 1 else terrain_bands(peak, rank, world, barrier, dev)
    wfilter = window_filter_bench(peak) if world == 1 else None
    general = general_path_run(peak, dev) if world == 1 else None
    if rank == 0:
        stencil["cpu_baseline"] = terrain_cpu_baseline(TERRAIN_N)
        g, dt = oracle_sample_gcells(2500, COLS, total_rows)
        cpu = {"value": g, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"first 2500 rows x {COLS} cols of the same DEM, oracle/rwatershed.c "
                         f"(restatement of single-threaded GRASS r.watershed -s -a), {dt:.1f} s",
               "sample_rows": 2500, "extrapolated": True,
               "host_cores_available": os.cpu_count()}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic", "config": workload_config(world),
            "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": launches, "step_mode": step_mode,
            "roofline": roofline, "path_roofline": path_roofline, "kernels": kernels,
            "boundary_graph": {"resolve_rounds": resolve_rounds,
                               "note": "pointer-doubling rounds of stage B until no node is active"},
            "parity": parity, "weighted": weighted, "target_configs": targets,
            "general_path": general,
            "stencil_roofline": stencil, "window_filter": wfilter,
            "cpu_baseline": cpu,
        }))
    if world > 1:
        dist.destroy_process_group()


def load_traffic(kernel_name):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture
    (profiles/traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel_name)
    except Exception:
        return None


def terrain_roofline(peak):
    """configs[2] on one GPU: fused slope+aspect+TRI on 32768^2, 16 B/cell."""
    import torch
    from hydrotools.cuda import _native as N, _device as D, synth
    n = TERRAIN_N
    dem, ld = synth.synth_dem(SEED, n, n)
    outs = [torch.empty((n, ld), dtype=torch.float32, device="cuda") for _ in range(3)]
    st = D.stream_ptr()

    def run():
        N.call("ht_terrain_fused_f32", dem.data_ptr(), n, n, ld, 0, 0.0, CSX, CSY, 1.0, 0,
               outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), ld, 0, 0, st)

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gbs = n * n * 16 / (ms * 1e-3) / 1e9
    del outs, dem
    torch.cuda.empty_cache()
    return {"kernel": "terrain_kernel (fused slope+aspect+TRI)", "workload": f"{n}x{n} fp32",
            "ms": ms, "gcells_s": n * n / (ms * 1e-3) / 1e9, "achieved": gbs, "peak": peak,
            "unit": "GB/s", "frac": gbs / peak, "algorithmic_bytes_per_cell": 16}


def window_filter_bench(peak):
    """SURVEY.md 8(f) N3 (interpolate.raster_filter): 3 x 3 max (streaming: 4 B read +
    4 B written per cell) and a disc of radius 10 cells, mean (317 samples per cell, fp64
    like the reference) on 16384^2; CPU baseline: the numpy restatement of the reference's
    numba loop on a bounded sample, one core."""
    import ctypes as C
    import numpy as np
    import torch
    from hydrotools.cuda import _native as N, _device as D, synth
    from oracle import rasterfilter as orf
    n = 16384
    dem, ld = synth.synth_dem(SEED, n, n)
    out = torch.empty((n, ld), dtype=torch.float32, device="cuda")
    st = D.stream_ptr()
    y, x = np.mgrid[-10:11, -10:11]
    res = []
    for name, k, method in (("3x3 max", np.ones((3, 3), np.uint8), "max"),
                            ("disc r=10 mean (run differences of row prefix sums built in shared memory)", ((x * x + y * y) <= 100).astype(np.uint8), "mean")):
        k = np.ascontiguousarray(k)
        code = {"max": 2, "mean": 4}[method]

        def run():
            N.call("ht_window_filter_f32", dem.data_ptr(), n, n, ld, 0, 0.0,
                   k.ctypes.data_as(C.c_void_p), k.shape[0], k.shape[1], code, out.data_ptr(), ld,
                   0, 0, st)
        if method == "mean":  # many samples: the public entry picks the row-prefix-sum kernel
            import hydrotools.cuda as hc
            view = dem[:, :n]

            def run():
                hc.window_filter(view, k.astype(bool), "mean")
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        sample = dem[:256, :2048].cpu().numpy()
        t0 = time.perf_counter()
        orf.raster_filter(sample, k.astype(bool), method)
        dt = time.perf_counter() - t0
        gbs = n * n * 8 / (ms * 1e-3) / 1e9
        res.append({"kernel": name, "samples_per_cell": int(k.sum()), "ms": ms,
                    "gcells_s": n * n / (ms * 1e-3) / 1e9, "achieved": gbs, "peak": peak,
                    "unit": "GB/s", "frac": gbs / peak, "algorithmic_bytes_per_cell": 8,
                    "cpu_baseline": {"value": sample.size / dt / 1e9, "unit": UNIT, "cores": 1,
                                     "kind": "port", "sample": f"256 x 2048 cells, oracle/rasterfilter.py, {dt:.2f} s"}})
    del dem, out
    torch.cuda.empty_cache()
    return {"workload": f"{n}x{n} fp32 (next row N3: interpolate.raster_filter)", "cases": res}


def terrain_cpu_baseline(n):
    """gdaldem is single-threaded and the reference runs it three times (slope, aspect,
    TRI: elevation.py:41-54, 70-78, 118-126); the oracle restatement computes the three
    in one pass, on a bounded row sample of the same DEM: one core (what gdaldem uses) and
    all host cores (rows split over pthreads)."""
    import oracle
    rows = 2048
    dem = oracle.synth_dem(SEED, rows, n, row0=0, total_rows=n)
    oracle.gdaldem(dem[:64], None, CSX, CSY)
    t0 = time.perf_counter()
    oracle.gdaldem(dem, None, CSX, CSY, threads=1)
    dt1 = time.perf_counter() - t0
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    oracle.gdaldem(dem, None, CSX, CSY, threads=cores)
    dtn = time.perf_counter() - t0
    return {"value": rows * n / dt1 / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"first {rows} rows x {n} cols, oracle/gdaldem.c (slope + aspect + TRI in one "
                      f"pass; the reference spawns gdaldem once per output), {dt1:.1f} s",
            "sample_rows": rows, "extrapolated": True,
            "all_cores": {"value": rows * n / dtn / 1e9, "unit": UNIT, "cores": cores,
                          "kind": "port", "sample": f"same sample, rows split over {cores} threads, {dtn:.2f} s"}}


def terrain_bands(peak, rank, world, barrier, dev):
    import torch
    import torch.distributed as dist
    from hydrotools.cuda import distributed as HD
    n = TERRAIN_N
    rows = n // world
    eng = HD.BandEngine(rank, world, rows, n, rows * world, CSX, CSY, nodata=None)
    eng.generate_synthetic(SEED)
    for _ in range(3):
        out = eng.terrain()
    del out
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = eng.terrain()
        del out
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / 10], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    cells = rows * world * n
    gbs_per_gpu = rows * n * 16 / (ms * 1e-3) / 1e9
    return {"kernel": "terrain_kernel (fused slope+aspect+TRI)", "workload": f"{rows * world}x{n} fp32 in "
            f"{world} row bands, 1 halo row exchanged per shared side and step", "ms": ms,
            "gcells_s": cells / (ms * 1e-3) / 1e9, "achieved": gbs_per_gpu, "peak": peak,
            "unit": "GB/s per GPU", "frac": gbs_per_gpu / peak, "algorithmic_bytes_per_cell": 16,
            "scaling": "strong"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
