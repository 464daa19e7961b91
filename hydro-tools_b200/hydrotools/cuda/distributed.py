"""Row-band sharding of the hot path across the GPUs of one box
(BASELINE.json north_star: "The DEM is row-band sharded
... Stencil halos and accumulation boundary flows are exchanged with NCCL over
NVLink, and the boundary graph is resolved iteratively until no flow changes").

One process per GPU (torchrun); ``torch.distributed`` is the plumbing:
  * DEM halo rows (2 per shared side for D8, 1 for the terrain stencils): every band
    contributes its first / last rows to one all-gather and picks its neighbours' rows;
  * accumulation: every band resolves its own tile boundary graph first, then
    publishes two small tables per shared side -- the flow it delivers into each
    cell of the neighbour's edge row, and, for each of its own edge-row cells,
    through which edge cell the path that enters there leaves the band again.
    The tables (2 x cols x 12 B per band) are all-gathered, every rank resolves
    the band-level forest (same pointer-doubling kernel, <= 2*P*cols nodes), adds
    the inflow it receives to its own boundary graph and resolves that once
    more.  Exact in one exchange; no per-iteration collective.

``BandEngine`` holds the orchestration and is backend-agnostic so that the
exchange logic is testable on CPU (gloo, world_size 2) with a CPU stand-in for
the kernels; ``CudaBand`` is the product backend.
"""
from typing import Optional, Tuple

import torch
import torch.distributed as dist

T = 64          # tile edge of the accumulation kernels (csrc/ht_acc.cu)
SLOTS = 4 * T


def band_edge_slots(rows: int, cols: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """Boundary-graph slot of every cell of the band's first / last row
    (mirrors slot_of() in csrc/ht_acc.cu)."""
    tiles_x = (cols + T - 1) // T
    c = torch.arange(cols, device=device, dtype=torch.int64)
    tx, lx = c // T, c % T
    top = tx * SLOTS + lx
    ty = (rows - 1) // T
    h = rows - ty * T
    bot = (ty * tiles_x + tx) * SLOTS + (lx if h == 1 else T + lx)
    return top, bot


def build_band_graph(delivered: torch.Tensor, link: torch.Tensor):
    """Band-level forest from the all-gathered tables.

    delivered[b, s, c]: flow band b hands to the cell in column c of the edge row
    of the band above (s = 0) / below (s = 1).  link[b, s, c]: for band b's own
    cell in column c of its first (s = 0) / last (s = 1) row, -1 if a path that
    enters there ends inside the band, else side*cols + c' of the foreign cell it
    reaches (side 0 = band above).  Node (b, s, c) = that own cell; returns
    (base, next) flattened.
    """
    P, _, cols = delivered.shape
    base = torch.zeros_like(delivered)
    base[1:, 0] = delivered[:-1, 1]   # my first row receives what the band above sends down
    base[:-1, 1] = delivered[1:, 0]   # my last row receives what the band below sends up
    b = torch.arange(P, device=link.device, dtype=torch.int64).view(P, 1, 1)
    L = link.to(torch.int64)
    side = torch.div(L, cols, rounding_mode="floor")
    c2 = L - side * cols
    up = ((b - 1) * 2 + 1) * cols + c2    # last row of the band above
    down = ((b + 1) * 2 + 0) * cols + c2  # first row of the band below
    nxt = torch.where(L < 0, torch.full_like(L, -1), torch.where(side == 0, up, down))
    return base.reshape(-1), nxt.reshape(-1).to(torch.int32)


class BandEngine:
    """Runs flow_direction_accumulation / terrain on one row band per rank."""

    def __init__(self, rank: int, world: int, rows: int, cols: int, total_rows: int,
                 csx: float, csy: float, nodata: Optional[float] = None, band=None,
                 group=None, row0: Optional[int] = None):
        self.rank, self.world = rank, world
        self.rows, self.cols, self.total_rows = rows, cols, total_rows
        self.group = group
        self.launches = 0
        self._halo_buf = {}
        if band is None:
            if row0 is None:
                row0 = rank * rows if world > 1 else 0
            band = CudaBand(rows, cols, row0, total_rows, csx, csy, nodata,
                            has_above=rank > 0, has_below=rank < world - 1)
        self.band = band

    # ----------------------------------------------------------------- data
    def generate_synthetic(self, seed: int, with_nulls: bool = False):
        """Own rows only; halos arrive by exchange like for any other DEM."""
        self.band.generate_synthetic(seed, with_nulls)

    def load(self, dem):
        self.band.load(dem)

    def dem_view(self):
        return self.band.dem_view()

    def validate(self):
        if self.world > 1:
            self.exchange_halos(2)
        c = self.band.validate()
        if self.world > 1:
            dist.all_reduce(c, group=self.group)
        c = c.cpu()
        if int(c[2]):
            raise ValueError(f"{int(c[2])} cells have |elevation| >= 2**24 mm (16 777 m); not supported")
        return int(c[0]), int(c[1])

    # ----------------------------------------------------------------- comm
    def exchange_halos(self, depth: int, async_op: bool = False):
        """`depth` DEM rows to / from each band neighbour.  One all-gather of every
        band's first and last `depth` rows (2 * depth * cols * 4 B per band): a single
        collective costs less than two send/recv pairs, and over NVSwitch the extra
        rows are free.  With ``async_op`` the collective is only enqueued (on NCCL's
        stream); the returned callable waits for it and places the rows."""
        band, P = self.band, self.world
        if hasattr(band, "pack_edge_rows"):
            send = band.pack_edge_rows(depth)  # preallocated [2, depth, cols]
        else:
            send = torch.stack([band.edge_rows(top=True, depth=depth),
                                band.edge_rows(top=False, depth=depth)])  # [2, depth, cols]
        key = (depth, send.dtype)
        every = self._halo_buf.get(key)
        if every is None:
            every = torch.empty((P,) + tuple(send.shape), dtype=send.dtype, device=send.device)
            self._halo_buf[key] = every
        work = dist.all_gather_into_tensor(every.view(P * 2, depth, self.cols), send, group=self.group,
                                           async_op=async_op)

        def finish():
            if async_op:
                work.wait()
            if self.rank > 0:
                band.set_halo_rows(top=True, rows=every[self.rank - 1, 1])
            if self.rank < P - 1:
                band.set_halo_rows(top=False, rows=every[self.rank + 1, 0])
        if async_op:
            return finish
        finish()

    # ----------------------------------------------------------------- hot path
    def flow_direction_accumulation(self, mode: int = 0, positive_only: bool = True):
        """D8 + accumulation of this rank's band (results(): direction int8, accumulation
        fp64).  ``positive_only=False`` reproduces r.watershed without ``-a``
        (/root/reference/src/hydrotools/watershed.py:52-54): accumulation is negative
        at edge cells and below cells whose sign r.watershed flipped -- a second
        sharded accumulation over the taint flags decides that."""
        band = self.band
        if self.world > 1 and hasattr(band, "direction_part") and self.rows >= 4 * band.D8_TILE_ROWS:
            # the interior rows need no halo: their D8 tiles run while the two halo rows of
            # each neighbour are in flight; the first and last tile rows follow the exchange
            work = self.exchange_halos(2, async_op=True)
            e = band.D8_TILE_ROWS
            band.direction_part(e, self.rows - e)
            work()
            band.direction_part(0, e)
            band.direction_part(self.rows - e, self.rows)
        else:
            if self.world > 1:
                self.exchange_halos(2)
            band.direction()
        self.accumulate(mode)
        if not positive_only:
            band.stash_accumulation()
            self.accumulate(band.N.MODE_TAINT)
            band.apply_sign()
            self.launches += band.take_launches()

    def graph(self, fn, warmup: int = 3):
        """Capture one call of ``fn`` (a method of this engine working on its resident
        buffers, e.g. ``engine.flow_direction_accumulation``) into a CUDA graph and return
        ``replay()``.  A sharded step is ~40 launches, memsets and two collectives; at 2 ms
        per step the host cannot enqueue them as fast as the GPU retires them (8 ranks share
        the box's cores), one graph launch can.  The NCCL collectives and the side-stream
        fork / join of the step are captured with it (every rank must capture and replay in
        step).  Buffers are the engine's own: results() after replay() reads the last step."""
        cap = torch.cuda.Stream(device=self.band.dev)
        cap.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cap):
            for _ in range(max(warmup, 1)):  # lazy state: TMA encoder, NCCL channels, side stream
                fn()
        torch.cuda.current_stream().wait_stream(cap)
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier(group=self.group)
        l0 = self.launches
        g = torch.cuda.CUDAGraph()
        # thread_local: NCCL's watchdog thread polls its events while this thread captures
        with torch.cuda.graph(g, stream=cap, capture_error_mode="thread_local"):
            fn()
        per_replay = self.launches - l0
        self.launches = l0

        def replay():
            g.replay()
            self.launches += per_replay
        def release():
            # a process group cannot be torn down while a graph still holds its collectives
            torch.cuda.synchronize()
            g.reset()
        replay.graph = g
        replay.release = release
        return replay

    def load_direction(self, direction):
        """A STORED direction raster (this rank's rows; int8 GRASS codes, -128 NULL) instead
        of a DEM: FlowAccumulation(direction), /root/reference/src/hydrotools/watershed.py:
        329-339, sharded.  Edge rows travel to the band neighbours (one all-gather of two int8
        rows per band), then the band is packed with r.flowaccumulation's semantics -- a code
        whose receiver is NULL or outside the raster ends the path, every other cell passes
        its flow on -- exactly as hydrotools.cuda.flow_accumulation_arrays does unsharded."""
        band, P = self.band, self.world
        band.set_direction(direction)
        if P > 1:
            send = band.dir_edge_rows()  # [2, cols] int8: first, last row
            every = torch.empty((P * 2, self.cols), dtype=send.dtype, device=send.device)
            dist.all_gather_into_tensor(every, send, group=self.group)
            every = every.view(P, 2, self.cols)
            band.set_dir_halo(every[self.rank - 1, 1] if self.rank > 0 else None,
                              every[self.rank + 1, 0] if self.rank < P - 1 else None)
        band.pack_direction()
        self.launches += band.take_launches()

    def flow_accumulation(self, weight, weight_nodata: Optional[float] = None):
        """Weighted accumulation (FlowAccumulation.calculate(..., "sum"),
        /root/reference/src/hydrotools/watershed.py:357-382) over the direction grid the
        last flow_direction_accumulation() left on the device: ``weight`` = this rank's rows
        of the weight raster.  Returns this rank's rows of the fp64 sums."""
        band = self.band
        # the result lands in the band's second accumulation buffer and is returned as a view of
        # it: valid until the next flow_accumulation / signed run of this engine (no copy of
        # 8 B/cell inside the hot path)
        band.stash_accumulation()
        if weight is None:  # cell counts ("modals", watershed.py:341-355)
            self.accumulate(band.N.MODE_COUNT, want_dir=False)
        else:
            band.set_weight(weight, weight_nodata)
            self.accumulate(band.N.MODE_WEIGHTED)
        out = band.acc[:, :self.cols]
        band.unstash_accumulation()
        return out

    def accumulate(self, mode: int = 0, want_dir: bool = True):
        band = self.band
        if mode == 2 and hasattr(band, "weight_max"):
            # weighted sums are exact fixed point; the unit follows from the largest
            # |weight| of the WHOLE raster: one 4-byte MAX all-reduce
            m = band.weight_max()
            if self.world > 1:
                dist.all_reduce(m, op=dist.ReduceOp.MAX, group=self.group)
            band.set_weight_max(m)
        band.local(mode)
        if self.world == 1:
            band.resolve(False)
            band.final(mode, want_dir) if want_dir is False else band.final(mode)
            self.launches += band.take_launches()
            return
        P = self.world
        if hasattr(band, "publish_tables"):
            # product backend: tables, band-level forest and the hand-back of the imported
            # flow are kernels behind the C-ABI; torch only moves the all-gather
            delta = band.supports_delta(mode)
            if delta:
                # count modes: band-edge slots carry a mark through the first resolve, so
                # that the second one only visits the nodes downstream of the band edge
                band.mark_edges(mode)
            band.resolve(True)
            if delta and hasattr(band, "final_phase"):
                # Stage C in two passes.  Only the tiles with a marked entry slot (the ones on
                # or downstream of the band's first / last row) depend on what the neighbour
                # bands deliver: the others are written on the main stream while the tables
                # travel and the band-level forest and the delta resolve run on a side stream.
                band.tile_phases(mode)
                mine = band.publish_tables(mode)  # [4 or 6, cols] int64
                main = torch.cuda.current_stream()
                side = band.side_stream()
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    every = torch.empty((P * mine.shape[0], self.cols), dtype=torch.int64,
                                        device=mine.device)
                    dist.all_gather_into_tensor(every, mine, group=self.group)
                    band.import_flows(every, P, self.rank, mode, delta)
                    every.record_stream(side)
                mine.record_stream(side)
                band.final_phase(0, mode, want_dir)
                main.wait_stream(side)
                band.final_phase(1, mode, want_dir)
                self.launches += band.take_launches()
                return
            mine = band.publish_tables(mode)  # [4 or 6, cols] int64
            every = torch.empty((P * mine.shape[0], self.cols), dtype=torch.int64, device=mine.device)
            dist.all_gather_into_tensor(every, mine, group=self.group)
            band.import_flows(every, P, self.rank, mode, delta)
        else:
            band.resolve(True)
            delivered, link = band.band_tables(mode)
            # (P*k, cols): concatenation along dim 0 is the layout every backend accepts
            if delivered.dtype == torch.int64:
                both = torch.cat([delivered, link.to(torch.int64)])  # [4, cols]
                every = torch.empty((P * 4, self.cols), dtype=torch.int64, device=both.device)
                dist.all_gather_into_tensor(every, both, group=self.group)
                every = every.view(P, 4, self.cols)
                all_d, all_l = every[:, :2], every[:, 2:]
            else:
                all_d = torch.empty((P * 2, self.cols), dtype=delivered.dtype, device=delivered.device)
                all_l = torch.empty((P * 2, self.cols), dtype=link.dtype, device=link.device)
                dist.all_gather_into_tensor(all_d, delivered.contiguous(), group=self.group)
                dist.all_gather_into_tensor(all_l, link.contiguous(), group=self.group)
            base, nxt = build_band_graph(all_d.reshape(P, 2, self.cols), all_l.reshape(P, 2, self.cols))
            inflow = band.forest_resolve(base, nxt, mode).view(P, 2, self.cols)
            band.add_inflow(inflow[self.rank, 0], inflow[self.rank, 1], mode)
            band.resolve(False)
        band.final(mode, want_dir) if want_dir is False else band.final(mode)
        self.launches += band.take_launches()

    def terrain(self, scale: float = 1.0, percent: bool = False):
        if self.world > 1:
            self.exchange_halos(1)
        out = self.band.terrain(scale, percent)
        self.launches += self.band.take_launches()
        return out

    def window_filter(self, kernel, method: str):
        """Moving-window reducer (hydrotools.cuda.window_filter,
        /root/reference/src/hydrotools/interpolate.py:110-247) of this rank's rows of the
        DEM; the rows of the neighbour bands the window reaches into travel in one
        all-gather, like the stencil halos."""
        from .interpolate import _as_kernel, window_filter
        band, P = self.band, self.world
        k = _as_kernel(kernel)
        i0 = k.shape[0] // 2
        i1 = k.shape[0] - 1 - i0
        own = band.dem_view()
        top = bot = None
        d = max(i0, i1)
        if P > 1 and d > 0:
            if d > self.rows:
                raise ValueError("the kernel reaches beyond the neighbouring band")
            send = torch.stack([band.edge_rows(top=True, depth=d), band.edge_rows(top=False, depth=d)])
            every = torch.empty((P,) + tuple(send.shape), dtype=send.dtype, device=send.device)
            dist.all_gather_into_tensor(every.view(P * 2, d, self.cols), send, group=self.group)
            if self.rank > 0 and i0:
                top = every[self.rank - 1, 1][d - i0:]      # last i0 rows of the band above
            if self.rank < P - 1 and i1:
                bot = every[self.rank + 1, 0][:i1]          # first i1 rows of the band below
        parts = [x for x in (top, own, bot) if x is not None]
        a = torch.cat(parts) if len(parts) > 1 else own
        return window_filter(a, k, method, nodata=getattr(band, "nodata", None),
                             halo_top=0 if top is None else int(top.shape[0]),
                             halo_bot=0 if bot is None else int(bot.shape[0]))

    def results(self):
        return self.band.results()

    def time_kernels(self, iters: int = 5):
        return self.band.time_kernels(iters)


class CudaBand:
    """One rank's row band in HBM and the sm_100a kernels on it (product backend)."""

    def __init__(self, rows, cols, row0, total_rows, csx, csy, nodata, has_above, has_below):
        from . import _native as N, _device as D
        self.N, self.D = N, D
        self.dev = D.require_cuda()
        if (has_above or has_below) and rows < 2:
            raise ValueError("row bands need at least 2 rows")
        self.rows, self.cols, self.row0, self.total_rows = rows, cols, row0, total_rows
        self.csx, self.csy, self.nodata = float(abs(csx)), float(abs(csy)), nodata
        self.htop, self.hbot = (2 if has_above else 0), (2 if has_below else 0)
        self.dem, self.ld = D.empty_padded(rows, cols, torch.float32, self.dev, 4, 2, 2)
        self.packed, self.pld = D.empty_padded(rows, cols, torch.uint8, self.dev, 16, 1, 1)
        self.acc, self.ald = D.empty_padded(rows, cols, torch.float64, self.dev, 2)
        self.dir, self.dld = D.empty_padded(rows, cols, torch.int8, self.dev, 16)
        self.ws = torch.empty(int(N.lib.ht_acc_workspace_bytes(rows, cols)), dtype=torch.uint8,
                              device=self.dev)
        self._ws_w = None  # weighted mode: 128-bit sums and 8 B/cell of tile-local sums
        import ctypes
        lay = (ctypes.c_int64 * 11)()
        N.check(N.lib.ht_acc_workspace_layout(rows, cols, self.ws.data_ptr(), self.ws.numel(), lay),
                "ht_acc_workspace_layout")
        self.lay = list(lay)
        self.nslots, self.vfirst = self.lay[9], self.lay[10]
        self.idx_top, self.idx_bot = band_edge_slots(rows, cols, self.dev)
        self.weight, self.wld, self.wnd = None, 0, None
        self._launches = 0
        self._tmp = None
        self._acc2 = None
        self._mode = 0
        self._edge_buf = {}

    def _wsm(self, mode):
        """Workspace of `mode` (the weighted one is allocated on first use)."""
        if mode != self.N.MODE_WEIGHTED:
            return self.ws
        if self._ws_w is None:
            n = int(self.N.lib.ht_acc_workspace_bytes_mode(self.rows, self.cols, mode))
            self._ws_w = torch.empty(n, dtype=torch.uint8, device=self.dev)
        return self._ws_w

    # ---- pointers
    def _dem0(self):
        return self.dem.data_ptr() + 2 * self.ld * 4

    def _packed0(self):
        return self.packed.data_ptr() + self.pld

    def _ws_view(self, k, dtype):
        off, n = self.lay[k], self.nslots
        size = n * (8 if dtype in (torch.int64, torch.float64) else 4)
        return self.ws[off:off + size].view(dtype)

    def _vdtype(self, mode):
        return torch.float64 if mode == self.N.MODE_WEIGHTED else torch.int64

    def take_launches(self):
        n, self._launches = self._launches, 0
        return n

    # ---- data
    def generate_synthetic(self, seed, with_nulls=False):
        nd = -32767.0 if self.nodata is None else self.nodata
        self.N.call("ht_synth_dem_f32", self._dem0(), self.rows, self.cols, self.ld, seed,
                    self.row0, self.total_rows, int(with_nulls), float(nd), self.D.stream_ptr())

    def load(self, dem):
        a = self.D.as_2d(dem)
        view = self.dem[2:2 + self.rows, :self.cols]
        if isinstance(a, torch.Tensor):
            view.copy_(a.to(torch.float32), non_blocking=True)
        else:
            import numpy as np
            view.copy_(torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)),
                       non_blocking=False)

    def set_weight(self, w, nodata=None):
        """A float32 CUDA tensor whose rows are 16-byte aligned is used in place (it must stay
        unchanged until the accumulation has run); anything else is copied to the device."""
        a = self.D.as_2d(w)
        if (isinstance(a, torch.Tensor) and a.is_cuda and a.dtype == torch.float32 and a.stride(1) == 1
                and a.stride(0) % 4 == 0 and a.data_ptr() % 16 == 0
                and tuple(a.shape) == (self.rows, self.cols)):
            self.weight, self.wld = a, int(a.stride(0))
        else:
            self.weight, self.wld = self.D.upload(a, torch.float32, 4, self.dev)
        self.wnd = nodata

    def dem_view(self):
        return self.dem[2:2 + self.rows, :self.cols]

    # ---- a stored direction raster instead of a DEM
    def set_direction(self, d):
        if getattr(self, "dir_in", None) is None:
            self.dir_in, self.dil = self.D.empty_padded(self.rows, self.cols, torch.int8, self.dev,
                                                        16, 1, 1)
        a = self.D.as_2d(d)
        view = self.dir_in[1:1 + self.rows, :self.cols]
        if isinstance(a, torch.Tensor):
            view.copy_(a.to(torch.int8), non_blocking=True)
        else:
            import numpy as np
            view.copy_(torch.from_numpy(np.ascontiguousarray(a, dtype=np.int8)), non_blocking=False)

    def dir_edge_rows(self):
        return torch.stack([self.dir_in[1, :self.cols], self.dir_in[self.rows, :self.cols]])

    def set_dir_halo(self, above, below):
        if above is not None:
            self.dir_in[0, :self.cols].copy_(above)
        if below is not None:
            self.dir_in[self.rows + 1, :self.cols].copy_(below)

    def pack_direction(self):
        self.N.call("ht_dir_pack_band_i8", self.dir_in.data_ptr() + self.dil, self.dil, self.rows,
                    self.cols, self.row0, self.total_rows, int(self.htop > 0), int(self.hbot > 0),
                    self._packed0(), self.pld, self.D.stream_ptr())
        self._launches += 1

    def edge_rows(self, top, depth):
        r0 = 2 if top else 2 + self.rows - depth
        return self.dem[r0:r0 + depth, :self.cols].contiguous()

    def pack_edge_rows(self, depth):
        """First and last `depth` rows in one preallocated [2, depth, cols] buffer."""
        buf = self._edge_buf.get(depth)
        if buf is None:
            buf = torch.empty((2, depth, self.cols), dtype=torch.float32, device=self.dev)
            self._edge_buf[depth] = buf
        buf[0].copy_(self.dem[2:2 + depth, :self.cols])
        buf[1].copy_(self.dem[2 + self.rows - depth:2 + self.rows, :self.cols])
        return buf

    def set_halo_rows(self, top, rows):
        depth = rows.shape[0]
        r0 = 2 - depth if top else 2 + self.rows
        self.dem[r0:r0 + depth, :self.cols].copy_(rows)

    # ---- kernels
    def validate(self):
        counts = torch.zeros(3, dtype=torch.int64, device=self.dev)
        has, nd = self.D.nodata_args(self.nodata)
        self.N.call("ht_d8_validate_conditioned_f32", self._dem0(), self.rows, self.cols, self.ld,
                    has, nd, self.row0, self.total_rows, self.htop, self.hbot, counts.data_ptr(),
                    self.D.stream_ptr())
        return counts

    def direction(self):
        has, nd = self.D.nodata_args(self.nodata)
        self.N.call("ht_d8_f32", self._dem0(), self.rows, self.cols, self.ld, has, nd, self.csx,
                    self.csy, self.row0, self.total_rows, self.htop, self.hbot, self._packed0(),
                    self.pld, self.D.stream_ptr())
        self._launches += 1

    D8_TILE_ROWS = 32  # tile height of d8_kernel (csrc/ht_d8.cu): part boundaries fall on tile rows

    def direction_part(self, r_lo, r_hi):
        """D8 of rows [r_lo, r_hi) of the band (multiples of D8_TILE_ROWS, or the band's ends).
        Rows inside the band find their halo in the band itself; only the parts that touch a
        neighbour band need its halo rows and store its edge row's packed byte."""
        has, nd = self.D.nodata_args(self.nodata)
        top = 2 if (r_lo > 0 or self.htop) else 0
        bot = 2 if (r_hi < self.rows or self.hbot) else 0
        self.N.call("ht_d8_part_f32", self._dem0() + r_lo * self.ld * 4, r_hi - r_lo, self.cols,
                    self.ld, has, nd, self.csx, self.csy, self.row0 + r_lo, self.total_rows, top, bot,
                    int(r_lo == 0 and self.htop > 0), int(r_hi == self.rows and self.hbot > 0),
                    self._packed0() + r_lo * self.pld, self.pld, self.D.stream_ptr())
        self._launches += 1

    def _wargs(self, mode):
        if mode == self.N.MODE_WEIGHTED:
            has, nd = self.D.nodata_args(self.wnd)
            return self.weight.data_ptr(), self.wld, has, nd
        return None, 0, 0, 0.0

    def weight_max(self):
        """Bits of the largest finite |weight| of this band (int32 tensor of one element)."""
        w, wld, has, nd = self._wargs(self.N.MODE_WEIGHTED)
        m = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.N.call("ht_acc_weight_max", w, wld, self.rows, self.cols, has, nd, m.data_ptr(),
                    self.D.stream_ptr())
        self._launches += 1
        return m

    def set_weight_max(self, m):
        ws = self._wsm(self.N.MODE_WEIGHTED)
        self.N.call("ht_acc_set_weight_max", self.rows, self.cols, ws.data_ptr(), ws.numel(),
                    m.data_ptr(), self.D.stream_ptr())

    def local(self, mode=0):
        w, wld, has, nd = self._wargs(mode)
        ws = self._wsm(mode)
        self.N.call("ht_acc_local", self._packed0(), self.pld, self.rows, self.cols, self.row0,
                    self.total_rows, int(self.htop > 0), int(self.hbot > 0), mode, w, wld, has, nd,
                    ws.data_ptr(), ws.numel(), self.D.stream_ptr())
        self._launches += 1
        self._mode = mode

    def resolve(self, want_roots):
        ws = self._wsm(self._mode)
        self.N.call("ht_acc_resolve", self.rows, self.cols, self._mode, int(want_roots),
                    ws.data_ptr(), ws.numel(), None, self.D.stream_ptr())
        # count modes from 128 tiles on: the two-level resolve (csrc/ht_acc.cu resolve_two_level:
        # super_tile_kernel<1>, virtual_nodes, outer_nodes, coarse_resolve, super_tile_kernel<3>,
        # + edge_roots with roots); else the one cooperative kernel
        tiles = ((self.rows + T - 1) // T) * ((self.cols + T - 1) // T)
        force = self.N.lib.ht_acc_stage_b_mode(-1)  # query only: 0 automatic, 1 flat, 2 two levels
        two = self._mode != self.N.MODE_WEIGHTED and (force == 2 or (force == 0 and tiles >= 128))
        self._launches += (5 + int(bool(want_roots))) if two else 1

    def resolve_rounds(self) -> int:
        """Pointer-doubling rounds stage B needs on the current boundary graph."""
        r = torch.zeros(1, dtype=torch.int32, device=self.dev)
        ws = self._wsm(self._mode)
        self.N.call("ht_acc_resolve", self.rows, self.cols, self._mode, 0, ws.data_ptr(),
                    ws.numel(), r.data_ptr(), self.D.stream_ptr())
        return int(r.item())

    def final(self, mode=0, want_dir=True):
        w, wld, has, nd = self._wargs(mode)
        ws = self._wsm(mode)
        self.N.call("ht_acc_final", self._packed0(), self.pld, self.rows, self.cols, self.row0,
                    self.total_rows, int(self.htop > 0), int(self.hbot > 0), mode, w, wld, has, nd,
                    self.acc.data_ptr(), self.ald,
                    self.dir.data_ptr() if (want_dir and mode == 0) else None, self.dld,
                    ws.data_ptr(), ws.numel(), self.D.stream_ptr())
        self._launches += 1

    def tile_phases(self, mode=0):
        """After resolve(True) on a marked graph: one byte per tile, 1 = the tile's inflow
        depends on the neighbour bands' deliveries (an entry slot carries a mark)."""
        ws = self._wsm(mode)
        self.N.call("ht_acc_tile_phases", self.rows, self.cols, mode, ws.data_ptr(), ws.numel(),
                    self.D.stream_ptr())
        self._launches += 1

    def final_phase(self, phase, mode=0, want_dir=True):
        """Stage C of the tiles of one phase (see tile_phases)."""
        w, wld, has, nd = self._wargs(mode)
        ws = self._wsm(mode)
        self.N.call("ht_acc_final_phase", self._packed0(), self.pld, self.rows, self.cols, self.row0,
                    self.total_rows, int(self.htop > 0), int(self.hbot > 0), mode, w, wld, has, nd,
                    self.acc.data_ptr(), self.ald,
                    self.dir.data_ptr() if (want_dir and mode == 0) else None, self.dld,
                    ws.data_ptr(), ws.numel(), int(phase), self.D.stream_ptr())
        self._launches += 1

    def side_stream(self):
        """High-priority stream for the table exchange chain of a sharded accumulation."""
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self.dev, priority=-1)
        return self._side

    MARK = 1 << 40  # band-edge mark in the u64 sums of the boundary graph (csrc/ht_acc.cu)

    def supports_delta(self, mode):
        return True

    def mark_edges(self, mode=0):
        """After stage A: tag the slots of the band's first / last row (where flow from a
        neighbour band can come in) so that the resolve leaves a mark on every node
        downstream of them."""
        ws = self._wsm(mode)
        self.N.call("ht_acc_mark_edges", self.rows, self.cols, mode, int(self.htop > 0),
                    int(self.hbot > 0), ws.data_ptr(), ws.numel(), self.D.stream_ptr())
        self._launches += 1

    def resolve_delta(self, top, bot, mode=0):
        """Second resolve of a sharded run, over the marked nodes only."""
        top = top.contiguous() if self.htop else None
        bot = bot.contiguous() if self.hbot else None
        ws = self._wsm(mode)
        self.N.call("ht_acc_resolve_delta", self.rows, self.cols, mode,
                    top.data_ptr() if top is not None else None,
                    bot.data_ptr() if bot is not None else None,
                    ws.data_ptr(), ws.numel(), self.D.stream_ptr())
        self._launches += 2

    def publish_tables(self, mode=0):
        """The tables this band publishes after resolve(want_roots=True), one kernel:
        count modes [delivered up, delivered down, link top, link bottom] x cols; weighted
        mode the two deliveries as 128-bit integers (low, high words): 6 x cols (int64)."""
        k = 6 if mode == self.N.MODE_WEIGHTED else 4
        out = torch.empty((k, self.cols), dtype=torch.int64, device=self.dev)
        ws = self._wsm(mode)
        self.N.call("ht_band_tables", self.rows, self.cols, mode, ws.data_ptr(), ws.numel(),
                    out.data_ptr(), self.D.stream_ptr())
        self._launches += 1
        return out

    def band_tables4(self):
        return self.publish_tables(0)

    def band_graph(self, every, bands, mode=0):
        """Band-level forest (base: u64 as int64, or [n, 2] int64 = 128-bit values in
        weighted mode; next int32) from the all-gathered tables."""
        n = bands * 2 * self.cols
        wide = mode == self.N.MODE_WEIGHTED
        base = torch.empty((n, 2) if wide else (n,), dtype=torch.int64, device=self.dev)
        nxt = torch.empty(n, dtype=torch.int32, device=self.dev)
        self.N.call("ht_band_graph", every.data_ptr(), bands, self.cols, mode, base.data_ptr(),
                    nxt.data_ptr(), self.D.stream_ptr())
        self._launches += 1
        return base, nxt

    def import_flows(self, every, bands, rank, mode, delta):
        """Band-level forest from the all-gathered tables, its resolve, and the hand-back of
        the flow this band imports: a resolve over the marked nodes (count modes) or the
        imports added to the boundary graph and a second full resolve (weighted)."""
        base, nxt = self.band_graph(every, bands, mode)
        inflow = self.forest_resolve(base, nxt, mode)
        wide = mode == self.N.MODE_WEIGHTED
        inflow = inflow.view(bands, 2, self.cols, 2) if wide else inflow.view(bands, 2, self.cols)
        top, bot = inflow[rank, 0], inflow[rank, 1]  # contiguous slices
        if delta:
            self.resolve_delta(top, bot, mode)
            return
        ws = self._wsm(mode)
        self.N.call("ht_acc_add_imports", self.rows, self.cols, mode,
                    top.data_ptr() if self.htop else None, bot.data_ptr() if self.hbot else None,
                    ws.data_ptr(), ws.numel(), self.D.stream_ptr())
        self._launches += 1
        self.resolve(False)

    def band_tables(self, mode=0):
        """Count modes, torch formulation of publish_tables (tests compare the two):
        (delivered [2, cols], link [2, cols]) after resolve(want_roots=True)."""
        if mode == self.N.MODE_WEIGHTED:
            raise ValueError("weighted tables are 128-bit: use publish_tables")
        inflow = self._ws_view(2, torch.int64)       # aggB
        root = self._ws_view(7, torch.int32)  # rootB
        delivered = inflow[self.vfirst:self.vfirst + 2 * self.cols].view(2, self.cols).clone()
        delivered &= self.MARK - 1
        link = torch.stack([root[self.idx_top], root[self.idx_bot]]).to(torch.int64) - self.vfirst
        link = torch.where(link >= 0, link, torch.full_like(link, -1)).to(torch.int32)
        return delivered, link

    def forest_resolve(self, base, nxt, mode=0):
        n = nxt.numel()
        out = torch.empty_like(base)
        wide = mode == self.N.MODE_WEIGHTED
        need = (56 if wide else 32) * n + 8192
        if self._tmp is None or self._tmp.numel() < need:
            self._tmp = torch.empty(need, dtype=torch.uint8, device=self.dev)
        self.N.call("ht_forest_resolve", base.data_ptr(), nxt.data_ptr(), n, 2 if wide else 0,
                    out.data_ptr(), self._tmp.data_ptr(), self._tmp.numel(), self.D.stream_ptr())
        self._launches += 1
        return out

    def add_inflow(self, top, bot, mode=0):
        """Count modes: imports added to the boundary graph before a second full resolve."""
        self.N.call("ht_acc_add_imports", self.rows, self.cols, mode,
                    top.contiguous().data_ptr() if self.htop else None,
                    bot.contiguous().data_ptr() if self.hbot else None,
                    self.ws.data_ptr(), self.ws.numel(), self.D.stream_ptr())
        self._launches += 1

    def terrain(self, scale=1.0, percent=False):
        outs = [self.D.empty_padded(self.rows, self.cols, torch.float32, self.dev, 4) for _ in range(3)]
        has, nd = self.D.nodata_args(self.nodata)
        self.N.call("ht_terrain_fused_f32", self._dem0(), self.rows, self.cols, self.ld, has, nd,
                    self.csx, self.csy, float(scale), int(percent), outs[0][0].data_ptr(),
                    outs[1][0].data_ptr(), outs[2][0].data_ptr(), outs[0][1],
                    int(self.htop > 0), int(self.hbot > 0), self.D.stream_ptr())
        self._launches += 2
        return tuple(o[0][:, :self.cols] for o in outs)

    def results(self):
        return self.dir[:, :self.cols], self.acc[:, :self.cols]

    # a second accumulation (taint counts, weights) lands in self.acc: keep the first
    def stash_accumulation(self):
        if self._acc2 is None:
            self._acc2 = torch.empty_like(self.acc)
        self.acc, self._acc2 = self._acc2, self.acc

    def unstash_accumulation(self):
        self.acc, self._acc2 = self._acc2, self.acc

    def apply_sign(self):
        """After stash_accumulation() + a MODE_TAINT run: self.acc holds the taint counts,
        the stash the cell counts; negate the counts where r.watershed would."""
        taint = self.acc
        self.unstash_accumulation()
        self.N.call("ht_acc_apply_sign", self.acc.data_ptr(), self.ald, taint.data_ptr(), self.ald,
                    self._packed0(), self.pld, self.rows, self.cols, self.D.stream_ptr())
        self._launches += 1

    def time_kernels(self, iters=5):
        """Average device time of each stage (CUDA events on the launching stream)."""
        stages = [("d8_kernel", self.direction),
                  ("acc_local_cnt_kernel", lambda: self.local(0)),
                  ("coarse_resolve_kernel", lambda: self.resolve(False)),
                  ("acc_final_cnt_kernel", lambda: self.final(0))]
        out = []
        for name, fn in stages:
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            row = {"name": name, "ms": e0.elapsed_time(e1) / iters}
            if name == "coarse_resolve_kernel":
                row["launches"] = ("stage B = ht_acc_resolve: super_tile_kernel<1>, virtual_nodes_kernel, "
                                   "outer_nodes_kernel, coarse_resolve_kernel on the outer nodes, "
                                   "super_tile_kernel<3> (two-level resolve, DESIGN.md 3.3)")
            out.append(row)
        self.take_launches()
        return out
