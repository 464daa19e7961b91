"""A stream of rasters through ``flow_direction_accumulation_arrays`` with the host link
busy in both directions at once.

One call of ``flow_direction_accumulation_arrays`` is upload -> kernels -> download, one
after the other: 400 MB in and 900 MB out for a 10 000 x 10 000 DEM cost 24 ms of PCIe time
around 2 ms of kernels.  The callers of the reference's function
(/root/reference/src/hydrotools/watershed.py:30-75) work through lists of DEMs
(``stream_analysis`` per region, /root/reference/src/hydrotools/streams.py:537-544): with
several rasters in flight the upload of raster i + 1, the kernels of raster i and the download
of raster i - 1 overlap on three CUDA streams (PCIe is full duplex), and the steady state is
bound by the larger of the two transfers -- the 9 B/cell of results -- instead of their sum.

Same results, bit for bit, as the one-raster call (same kernels, same order per raster).
"""
from typing import Iterable, Iterator, Optional, Tuple

import numpy as np
import torch

from . import _native as N
from . import _device as D


class _Slot:
    """Device and pinned host buffers of one raster in flight."""

    def __init__(self, rows, cols, dev, own_outputs):
        self.dem, self.ld = D.empty_padded(rows, cols, torch.float32, dev, 4)
        self.packed, self.pld = D.empty_padded(rows, cols, torch.uint8, dev, 16)
        self.acc, self.ald = D.empty_padded(rows, cols, torch.float64, dev, 2)
        self.dir, self.dld = D.empty_padded(rows, cols, torch.int8, dev, 16)
        self.counts = torch.zeros(3, dtype=torch.int64, device=dev)
        self.counts_host = torch.zeros(3, dtype=torch.int64).pin_memory()
        self.stage = None  # pinned staging of a pageable input
        self.out_fd = self.out_fa = None
        if own_outputs:
            self.out_fd = torch.empty((rows, cols), dtype=torch.int8, pin_memory=True)
            self.out_fa = torch.empty((rows, cols), dtype=torch.float64, pin_memory=True)
        self.uploaded = torch.cuda.Event()
        self.validated = torch.cuda.Event()
        self.computed = torch.cuda.Event()
        self.downloaded = torch.cuda.Event()
        self.busy = False
        self.used = False


class FlowPipeline:
    """``flow_direction_accumulation_arrays`` over many rasters of ONE shape.

    >>> pipe = FlowPipeline(rows, cols, nodata=None, csx=10, csy=10)
    >>> for fd, fa in pipe.run(dems):        # dems: iterable of host arrays
    ...     use(fd, fa)                      # valid until two more rasters were taken

    ``depth`` rasters are in flight (2 is enough to keep both directions of the link busy;
    every slot holds 14 B/cell in HBM and, unless ``outs`` is given to run(), 9 B/cell of
    page-locked host memory).  Host inputs that are page-locked (``torch.Tensor.pin_memory``)
    are copied straight from where they are; pageable ones are staged.
    """

    def __init__(self, rows: int, cols: int, nodata: Optional[float] = None, csx: float = 1.0,
                 csy: float = 1.0, positive_only: bool = True, depth: int = 2,
                 own_outputs: bool = True):
        if rows <= 0 or cols <= 0:
            raise ValueError("empty raster")
        self.dev = D.require_cuda()
        self.rows, self.cols = int(rows), int(cols)
        self.nodata, self.csx, self.csy = nodata, float(abs(csx)), float(abs(csy))
        self.positive_only = positive_only
        self.slots = [_Slot(self.rows, self.cols, self.dev, own_outputs) for _ in range(max(2, depth))]
        self.ws = torch.empty(int(N.lib.ht_acc_workspace_bytes_mode(self.rows, self.cols, 0)),
                              dtype=torch.uint8, device=self.dev)
        self.taint = None
        if not positive_only:
            self.taint, self.tld = D.empty_padded(self.rows, self.cols, torch.float64, self.dev, 2)
        self.s_up = torch.cuda.Stream(device=self.dev)
        self.s_val = torch.cuda.Stream(device=self.dev)
        self.s_run = torch.cuda.Stream(device=self.dev)
        self.s_down = torch.cuda.Stream(device=self.dev)
        self.launches = 0

    # ---- stages ------------------------------------------------------------------
    def _upload(self, slot: _Slot, dem) -> None:
        a = D.as_2d(dem)
        if tuple(a.shape) != (self.rows, self.cols):
            raise ValueError(f"raster of shape {tuple(a.shape)} in a pipeline of {(self.rows, self.cols)}")
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(np.ma.getdata(a), dtype=np.float32))
        if t.dtype != torch.float32:
            t = t.to(torch.float32)
        if not t.is_cuda and not t.is_pinned():
            if slot.stage is None:
                slot.stage = torch.empty((self.rows, self.cols), dtype=torch.float32, pin_memory=True)
            slot.stage.copy_(t)  # the slot is free: its previous upload was consumed
            t = slot.stage
        view = slot.dem[:, :self.cols]
        with torch.cuda.stream(self.s_up):
            if slot.used:
                self.s_up.wait_event(slot.computed)  # the kernels of the slot's last raster read it
            view.copy_(t, non_blocking=True)
            slot.uploaded.record()
        # pits / ties / out-of-range cells decide which kernel chain the raster takes: counted
        # on a stream of its own so that the count does not queue behind the previous raster
        has, nd = D.nodata_args(self.nodata)
        with torch.cuda.stream(self.s_val):
            self.s_val.wait_event(slot.uploaded)
            N.call("ht_d8_validate_conditioned_f32", slot.dem.data_ptr(), self.rows, self.cols,
                   slot.ld, has, nd, 0, self.rows, 0, 0, slot.counts.data_ptr(), D.stream_ptr())
            slot.counts_host.copy_(slot.counts, non_blocking=True)
            slot.validated.record()
        self.launches += 1
        slot.busy = True

    def _compute(self, slot: _Slot) -> None:
        slot.validated.synchronize()
        pits, ties, out_of_range = (int(v) for v in slot.counts_host)
        if out_of_range:
            raise ValueError(f"{out_of_range} cells have |elevation| >= 2**24 mm (16 777 m); not supported")
        has, nd = D.nodata_args(self.nodata)
        with torch.cuda.stream(self.s_run):
            self.s_run.wait_event(slot.validated)
            if slot.used:
                self.s_run.wait_event(slot.downloaded)  # the outputs of the slot's last raster
            slot.used = True
            if pits or ties:
                # the general chain (r.watershed on any DEM) drives its rounds from the host
                from .watershed import d8_general
                packed, pld, _ = d8_general(slot.dem, slot.ld, self.rows, self.cols, self.nodata,
                                            self.csx, self.csy)
            else:
                packed, pld = slot.packed, slot.pld
                N.call("ht_d8_f32", slot.dem.data_ptr(), self.rows, self.cols, slot.ld, has, nd,
                       self.csx, self.csy, 0, self.rows, 0, 0, packed.data_ptr(), pld, D.stream_ptr())
            N.call("ht_acc_f64", packed.data_ptr(), pld, self.rows, self.cols, N.MODE_COUNT, None, 0,
                   0, 0.0, slot.acc.data_ptr(), slot.ald, slot.dir.data_ptr(), slot.dld,
                   self.ws.data_ptr(), self.ws.numel(), D.stream_ptr())
            self.launches += 4
            if not self.positive_only:
                N.call("ht_acc_f64", packed.data_ptr(), pld, self.rows, self.cols, N.MODE_TAINT, None,
                       0, 0, 0.0, self.taint.data_ptr(), self.tld, None, 0, self.ws.data_ptr(),
                       self.ws.numel(), D.stream_ptr())
                N.call("ht_acc_apply_sign", slot.acc.data_ptr(), slot.ald, self.taint.data_ptr(),
                       self.tld, packed.data_ptr(), pld, self.rows, self.cols, D.stream_ptr())
                self.launches += 4
            slot.computed.record()

    def _download(self, slot: _Slot, out_fd, out_fa) -> Tuple[torch.Tensor, torch.Tensor]:
        fd = slot.out_fd if out_fd is None else out_fd
        fa = slot.out_fa if out_fa is None else out_fa
        if fd is None or fa is None:
            raise ValueError("a pipeline built with own_outputs=False needs `outs`")
        with torch.cuda.stream(self.s_down):
            self.s_down.wait_event(slot.computed)
            fd.copy_(slot.dir[:, :self.cols], non_blocking=True)
            fa.copy_(slot.acc[:, :self.cols], non_blocking=True)
            slot.downloaded.record()
        return fd, fa

    # ---- driver ------------------------------------------------------------------
    def run(self, dems: Iterable, outs: Optional[Iterable] = None) -> Iterator[Tuple[np.ndarray, np.ndarray]]:
        """Yields ``(direction int8, accumulation float64)`` per raster, in order.  ``outs``:
        one ``(direction, accumulation)`` pair of page-locked torch tensors per raster to
        receive the results; without it the arrays are views of the pipeline's own
        page-locked buffers and stay valid until ``depth`` more rasters have been taken."""
        dems = iter(dems)
        outs = iter(outs) if outs is not None else None
        n = len(self.slots)
        pending = []  # (slot, fd, fa) whose download is in flight, oldest first
        cur = next(dems, None)
        i = 0
        if cur is not None:
            self._upload(self.slots[0], cur)
        while cur is not None:
            slot = self.slots[i % n]
            # kernels and download of raster i are queued BEFORE the host waits for an older
            # download: the device never idles on the host
            self._compute(slot)
            o = next(outs) if outs is not None else (None, None)
            fd, fa = self._download(slot, o[0], o[1])
            pending.append((slot, fd, fa))
            cur = next(dems, None)
            if cur is not None:
                # the slot raster i + 1 goes to must have given up its previous raster
                follow = self.slots[(i + 1) % n]
                while any(item[0] is follow for item in pending):
                    yield self._finish(pending.pop(0))
                self._upload(follow, cur)
            i += 1
        while pending:
            yield self._finish(pending.pop(0))

    def _finish(self, item):
        slot, fd, fa = item
        slot.downloaded.synchronize()
        slot.busy = False
        return fd.numpy(), fa.numpy()


def flow_direction_accumulation_many(dems: Iterable, nodata: Optional[float] = None, csx: float = 1.0,
                                     csy: float = 1.0, positive_only: bool = True,
                                     copy: bool = True) -> Iterator[Tuple[np.ndarray, np.ndarray]]:
    """``flow_direction_accumulation_arrays`` for every raster of ``dems`` (host arrays of one
    shape), transfers and kernels of neighbouring rasters overlapped (FlowPipeline).  With
    ``copy=False`` the yielded arrays are views of page-locked buffers that are overwritten
    two rasters later."""
    dems = iter(dems)
    first = next(dems, None)
    if first is None:
        return
    a = D.as_2d(first)
    pipe = FlowPipeline(int(a.shape[0]), int(a.shape[1]), nodata, csx, csy, positive_only)

    def chain():
        yield first
        yield from dems
    for fd, fa in pipe.run(chain()):
        yield (fd.copy(), fa.copy()) if copy else (fd, fa)
