"""B200 replacements for the downstream consumers of the direction grid:
the reference walks the grid cell by cell in numba / GRASS; here the device does it with
pointer jumping and path walkers (csrc/ht_streams.cu).

* ``reach_labels``   -- ``hydrotools.streams._fa_label_task``
  (/root/reference/src/hydrotools/streams.py:320-388)
* ``upstream_mean`` / ``upstream_stats`` -- ``hydrotools.morphology._upstream_compute`` and
  its path-level caller (/root/reference/src/hydrotools/morphology.py:50-153, 292-360)
* ``basin_mask`` / ``basin`` -- GRASS ``r.water.outlet``
  (/root/reference/src/hydrotools/watershed.py:292-318)
"""
import ctypes as C
import math
from typing import Optional

import numpy as np
import torch

from . import _native as N
from . import _device as D
from . import _io
from .watershed import _dir_to_int8


def _ws(rows, cols, dev, max_window=0, max_inlets=0):
    n = int(N.lib.ht_streams_workspace_bytes(rows, cols, max_window, max_inlets))
    return torch.empty(n, dtype=torch.uint8, device=dev)


def reach_labels(reaches: D.ArrayLike, direction: D.ArrayLike, nodata: int = 0):
    """Label the connected runs of equal reach value along the flow directions.

    ``reaches`` int32 (rows, cols), ``direction`` int8 GRASS codes.  Returns uint32 labels,
    0 where ``reaches == nodata``, numbered 1.. in row-major order of each run's first cell
    -- the numbering of the reference's scan."""
    dev = D.require_cuda()
    on_host = not isinstance(reaches, torch.Tensor)
    r = D.as_2d(reaches)
    rows, cols = int(r.shape[0]), int(r.shape[1])
    d = D.as_2d(direction)
    if tuple(d.shape) != (rows, cols):
        raise ValueError("reaches and direction grids do not align")
    rbuf, rld = D.upload(r, torch.int32, 4, dev)
    dbuf, dld = D.upload(d, torch.int8, 16, dev)
    out, old = D.empty_padded(rows, cols, torch.int32, dev, 4)
    ws = _ws(rows, cols, dev)
    N.call("ht_reach_labels_i32", rbuf.data_ptr(), rld, dbuf.data_ptr(), dld, rows, cols, int(nodata),
           out.data_ptr(), old, ws.data_ptr(), ws.numel(), D.stream_ptr())
    view = out[:, :cols]
    return D.download(view).view(np.uint32) if on_host else view


def upstream_mean(values: D.ArrayLike, streams: D.ArrayLike, direction: D.ArrayLike, csx: float,
                  csy: float, distance: float, out_fill: float = float("nan")):
    """Moving mean of ``values`` over ``distance`` map units upstream along the streams.

    ``values`` float64, ``streams`` boolean mask, ``direction`` int8 codes; returns float32
    (``out_fill`` where no inlet's path passes: inlets themselves and non-stream cells), with
    the reference's arithmetic (fp64 running sum along each path, last inlet in row-major order
    wins on shared reaches)."""
    dev = D.require_cuda()
    on_host = not isinstance(values, torch.Tensor)
    v = D.as_2d(values)
    rows, cols = int(v.shape[0]), int(v.shape[1])
    s, d = D.as_2d(streams), D.as_2d(direction)
    if tuple(s.shape) != (rows, cols) or tuple(d.shape) != (rows, cols):
        raise ValueError("values, streams and direction grids do not align")
    vbuf, vld = D.upload(v, torch.float64, 2, dev)
    s8 = s.to(torch.uint8) if isinstance(s, torch.Tensor) else np.ascontiguousarray(s, dtype=np.uint8)
    sbuf, sld = D.upload(s8, torch.uint8, 16, dev)
    dbuf, dld = D.upload(d, torch.int8, 16, dev)
    out, old = D.empty_padded(rows, cols, torch.float32, dev, 4)
    out.fill_(out_fill)
    max_window = int(math.ceil(float(distance) / min(abs(csx), abs(csy)))) + 2  # morphology.py:331
    max_inlets = int(torch.count_nonzero(sbuf[:, :cols]).item())
    ws = _ws(rows, cols, dev, max_window, max_inlets)
    got = C.c_longlong(0)
    N.call("ht_upstream_mean_f64", vbuf.data_ptr(), vld, sbuf.data_ptr(), sld, dbuf.data_ptr(), dld,
           rows, cols, float(abs(csx)), float(abs(csy)), float(distance), max_window, max_inlets,
           out.data_ptr(), old, ws.data_ptr(), ws.numel(), C.byref(got), D.stream_ptr())
    view = out[:, :cols]
    return D.download(view) if on_host else view


def basin_mask(direction: D.ArrayLike, row: int, col: int):
    """uint8 mask of the cells that drain through (row, col), the outlet included."""
    dev = D.require_cuda()
    on_host = not isinstance(direction, torch.Tensor)
    d = D.as_2d(direction)
    rows, cols = int(d.shape[0]), int(d.shape[1])
    if not (0 <= row < rows and 0 <= col < cols):
        raise ValueError("outlet outside the raster")
    dbuf, dld = D.upload(d, torch.int8, 16, dev)
    out, old = D.empty_padded(rows, cols, torch.uint8, dev, 16)
    ws = _ws(rows, cols, dev)
    N.call("ht_basin_u8", dbuf.data_ptr(), dld, rows, cols, int(row), int(col), out.data_ptr(), old,
           ws.data_ptr(), ws.numel(), D.stream_ptr())
    view = out[:, :cols]
    return D.download(view) if on_host else view


# ---------------------------------------------------------------- path-level mirrors
def upstream_stats(values: str, streams: str, flow_direction: str, gradient_dst: str, distance: float):
    """Same signature as /root/reference/src/hydrotools/morphology.py:292-298."""
    s, smeta = _io.read_raster(streams)
    v, vmeta = _io.read_raster(values)
    d, dmeta = _io.read_raster(flow_direction)
    mask = ~_mask_of(s, smeta["nodata"])
    vals = np.where(_mask_of(v, vmeta["nodata"]), np.nan, v.astype(np.float64))
    fd = _dir_to_int8(d, dmeta["nodata"])
    out = upstream_mean(vals, mask, fd, smeta["csx"], smeta["csy"], float(distance))
    nd = float(np.finfo("float32").min)  # infer_nodata("float32"), what to_raster fills masked cells with
    _io.write_raster(np.where(np.isnan(out), np.float32(nd), out), streams, gradient_dst, nodata=nd)


def basin(flow_direction: str, x: float, y: float, dst: str):
    """Same signature as /root/reference/src/hydrotools/watershed.py:292-300; raster
    destinations (``.tif``) only -- vectorising the mask stays with the reference's
    ``vectorize``."""
    if not dst.lower().endswith(".tif"):
        raise NotImplementedError("vector output is outside the CUDA path; write a .tif and vectorize it")
    d, meta = _io.read_raster(flow_direction)
    col = int(math.floor((x - meta["left"]) / meta["csx"]))
    row = int(math.floor((meta["top"] - y) / meta["csy"]))
    m = basin_mask(_dir_to_int8(d, meta["nodata"]), row, col)
    _io.write_raster(m.astype(np.uint8), flow_direction, dst, nodata=0)


def _mask_of(a: np.ndarray, nodata: Optional[float]) -> np.ndarray:
    m = np.zeros(a.shape, bool)
    if nodata is not None and not (isinstance(nodata, float) and np.isnan(nodata)):
        m |= a == nodata
    if a.dtype.kind == "f":
        m |= np.isnan(a)
    return m
