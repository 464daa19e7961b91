"""CPU (gloo, world_size 2 and 3): the row-band orchestration of
hydrotools.cuda.distributed.BandEngine -- halo send/recv, all-gather of the band
tables, band-level forest, inflow hand-back -- with a CPU stand-in for the
kernels built on the oracle.  The product backend (CudaBand) is exercised on the
GPU by tests/test_bands_gpu.py with the same table functions."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hydro-tools_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import oracle  # noqa: E402

ND = -32767.0
DR = np.array([0, -1, -1, -1, 0, 1, 1, 1, 0])
DC = np.array([0, 1, 0, -1, -1, -1, 0, 1, 1])


class OracleBand:
    """Band backend on CPU tensors: same interface as CudaBand, arithmetic from the
    oracle (test infrastructure only)."""

    def __init__(self, dem_whole, lo, hi, nodata):
        self.whole, self.lo, self.hi, self.nodata = dem_whole, lo, hi, nodata
        self.rows, self.cols = hi - lo, dem_whole.shape[1]
        self.total = dem_whole.shape[0]
        self.dem = np.full((self.rows + 4, self.cols), np.nan, np.float32)
        self.dem[2:2 + self.rows] = dem_whole[lo:hi]
        self.extra = np.zeros((self.rows, self.cols))
        self.launches = 0

    # ---- data / halos
    def edge_rows(self, top, depth):
        r0 = 2 if top else 2 + self.rows - depth
        return torch.from_numpy(self.dem[r0:r0 + depth].copy())

    def set_halo_rows(self, top, rows):
        depth = rows.shape[0]
        r0 = 2 - depth if top else 2 + self.rows
        self.dem[r0:r0 + depth] = rows.numpy()

    def take_launches(self):
        return 0

    # ---- "kernels"
    def direction(self):
        # the halo rows must have arrived: the band + halos equals the whole raster's rows
        a, b = max(self.lo - 2, 0), min(self.hi + 2, self.total)
        got = self.dem[2 - (self.lo - a):2 + self.rows + (b - self.hi)]
        np.testing.assert_array_equal(got, self.whole[a:b])
        fd, _ = oracle.rwatershed(self.whole, self.nodata, 10.0, 10.0)
        null = fd == oracle.DIR_NULL
        edge = np.zeros(fd.shape, bool)
        edge[[0, -1], :] = True
        edge[:, [0, -1]] = True
        pad = np.pad(null, 1, constant_values=False)
        for dr in range(3):
            for dc in range(3):
                edge |= pad[dr:dr + fd.shape[0], dc:dc + fd.shape[1]]
        eff = np.where((fd > 0) & ~edge, fd, 0).astype(np.int8)
        eff[null] = oracle.DIR_NULL
        self.fd_out = fd[self.lo:self.hi]
        self.eff_whole = eff
        # band-local graph: links that leave the band are cut
        d = eff[self.lo:self.hi].copy()
        r = np.arange(self.rows)[:, None] + DR[np.clip(d, 0, 8)]
        self.exits = (d > 0) & ((r < 0) | (r >= self.rows))
        self.d_local = np.where(self.exits, 0, d).astype(np.int8)

    def _accumulate(self, extra):
        w = (1.0 + extra).astype(np.float32)
        return oracle.flowacc(self.d_local, w)

    def local(self, mode=0):
        self.acc_local = self._accumulate(np.zeros((self.rows, self.cols)))

    def resolve(self, want_roots):
        pass

    def band_tables(self, mode=0):
        d = self.eff_whole[self.lo:self.hi]
        delivered = np.zeros((2, self.cols), np.int64)
        rr, cc = np.nonzero(self.exits)
        for r, c in zip(rr, cc):
            side = 0 if r + DR[d[r, c]] < 0 else 1
            delivered[side, c + DC[d[r, c]]] += int(self.acc_local[r, c])
        link = np.full((2, self.cols), -1, np.int32)
        for s, r0 in ((0, 0), (1, self.rows - 1)):
            for c0 in range(self.cols):
                r, c = r0, c0
                if d[r, c] == oracle.DIR_NULL:
                    continue
                while True:
                    k = d[r, c]
                    if k <= 0:
                        break
                    nr, nc = r + DR[k], c + DC[k]
                    if nr < 0 or nr >= self.rows:
                        link[s, c0] = (0 if nr < 0 else 1) * self.cols + nc
                        break
                    r, c = nr, nc
        return torch.from_numpy(delivered), torch.from_numpy(link)

    def forest_resolve(self, base, nxt, mode=0):
        base, nxt = base.numpy(), nxt.numpy()
        agg = base.copy()
        for u in np.nonzero(base)[0]:
            v = nxt[u]
            while v >= 0:
                agg[v] += base[u]
                v = nxt[v]
        return torch.from_numpy(agg)

    def add_inflow(self, top, bot, mode=0):
        self.extra[0] += top.numpy()
        self.extra[-1] += bot.numpy()

    def final(self, mode=0):
        self.acc = self._accumulate(self.extra)

    def results(self):
        return self.fd_out, self.acc

    def dem_view(self):
        return torch.from_numpy(self.dem[2:2 + self.rows].copy())


def _filter_worker(rank, world, port, cuts, kernel, method, q):
    """BandEngine.window_filter over gloo: the halo rows a window needs come from the
    neighbour bands; the per-band filter itself is the numpy restatement here."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hydrotools.cuda import distributed as HD
    from hydrotools.cuda import interpolate as I
    from oracle import rasterfilter as orf

    def cpu_window_filter(a, k, m, nodata=None, halo_top=0, halo_bot=0):
        a = a.numpy().astype(np.float32).copy()
        if nodata is not None:
            a[a == nodata] = np.nan
        out = orf.raster_filter(a, k, m)
        return torch.from_numpy(out[halo_top:a.shape[0] - halo_bot])

    I.window_filter = cpu_window_filter
    dem = oracle.synth_dem(3, cuts[-1], 97, with_nulls=True)
    lo, hi = cuts[rank], cuts[rank + 1]
    band = OracleBand(dem, lo, hi, ND)
    eng = HD.BandEngine(rank, world, hi - lo, 97, cuts[-1], 10.0, 10.0, ND, band=band)
    out = eng.window_filter(kernel, method)
    q.put((rank, out.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cuts,kernel,method", [([0, 60, 130], (5, 3), "mean"),
                                                ([0, 45, 90, 130], (8, 8), "std"),
                                                ([0, 30, 130], (1, 9), "max")])
def test_band_filter_over_gloo(cuts, kernel, method):
    from oracle import rasterfilter as orf
    world = len(cuts) - 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30100 + (os.getpid() % 500) + world
    procs = [ctx.Process(target=_filter_worker, args=(r, world, port, cuts, kernel, method, q))
             for r in range(world)]
    for p in procs:
        p.start()
    parts = dict()
    for _ in range(world):
        r, o = q.get(timeout=180)
        parts[r] = o
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    dem = oracle.synth_dem(3, cuts[-1], 97, with_nulls=True).astype(np.float32)
    dem[dem == ND] = np.nan
    exp = orf.raster_filter(dem, kernel, method)
    got = np.vstack([parts[r] for r in range(world)])
    np.testing.assert_array_equal(np.isnan(got), np.isnan(exp))
    m = ~np.isnan(exp)
    np.testing.assert_array_equal(got[m], exp[m])


def _worker(rank, world, port, cuts, seed, flip, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hydrotools.cuda import distributed as HD
    dem = oracle.synth_dem(seed, cuts[-1], 97, with_nulls=True)
    if flip:
        dem = np.ascontiguousarray(dem[::-1])
    lo, hi = cuts[rank], cuts[rank + 1]
    eng = HD.BandEngine(rank, world, hi - lo, 97, cuts[-1], 10.0, 10.0, ND,
                        band=OracleBand(dem, lo, hi, ND))
    eng.flow_direction_accumulation()
    fd, fa = eng.results()
    q.put((rank, fd, fa))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cuts,flip", [([0, 60, 130], False), ([0, 45, 90, 130], False),
                                       ([0, 70, 130], True)])
def test_band_engine_over_gloo(cuts, flip):
    world = len(cuts) - 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, cuts, 11, flip, q))
             for r in range(world)]
    for p in procs:
        p.start()
    parts = dict()
    for _ in range(world):
        r, fd, fa = q.get(timeout=180)
        parts[r] = (fd, fa)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    dem = oracle.synth_dem(11, cuts[-1], 97, with_nulls=True)
    if flip:
        dem = np.ascontiguousarray(dem[::-1])
    efd, efa = oracle.rwatershed(dem, ND, 10.0, 10.0)
    np.testing.assert_array_equal(np.vstack([parts[r][0] for r in range(world)]), efd)
    np.testing.assert_array_equal(np.vstack([parts[r][1] for r in range(world)]), efa)


def test_build_band_graph_indexing():
    from hydrotools.cuda import distributed as HD
    cols = 4
    delivered = torch.zeros((3, 2, cols), dtype=torch.int64)
    delivered[0, 1, 2] = 7      # band 0 hands 7 to column 2 of band 1's first row
    delivered[2, 0, 1] = 5      # band 2 hands 5 up to column 1 of band 1's last row
    link = torch.full((3, 2, cols), -1, dtype=torch.int32)
    link[1, 0, 2] = 1 * cols + 3  # enters band 1 at (top, 2), leaves downwards into column 3
    base, nxt = HD.build_band_graph(delivered, link)
    node = lambda b, s, c: (b * 2 + s) * cols + c  # noqa: E731
    assert base[node(1, 0, 2)] == 7 and base[node(1, 1, 1)] == 5 and base.sum() == 12
    assert nxt[node(1, 0, 2)] == node(2, 0, 3)
    assert (nxt >= 0).sum() == 1


def test_band_edge_slots_match_kernel_numbering():
    from hydrotools.cuda import distributed as HD
    top, bot = HD.band_edge_slots(130, 70, "cpu")
    # tile (0, 1), local column 5 -> slot 256 + 5; last row 129 is local row 1 of tile row 2
    assert top[69] == 1 * 256 + 5
    assert bot[69] == (2 * 2 + 1) * 256 + 64 + 5
    top, bot = HD.band_edge_slots(129, 10, "cpu")  # last tile row is one cell thick
    assert bot[3] == 2 * 256 + 3
