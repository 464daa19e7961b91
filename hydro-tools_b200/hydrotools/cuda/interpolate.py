"""B200 replacement for ``hydrotools.interpolate.raster_filter`` /
``custom_raster_filter`` (/root/reference/src/hydrotools/interpolate.py:16-247):
moving-window reducers over a rectangular or arbitrary boolean kernel.

The reference runs a numba loop per dask chunk under
``da.map_overlap(depth, boundary=nan)``; here one sm_100a kernel
(``ht_window_filter_f32``, csrc/ht_filter.cu) filters the raster with the tile +
halo staged by TMA, NaN outside the raster.  Reducers: sum, min, max, diff, mean,
std, variance, median, mode.  ``percentile`` and ``quantile`` (which cannot work in the
reference either: its reducers call ``np.nanpercentile(sample)`` without ``q``,
interpolate.py:230-240) and custom numba callables raise ``NotImplementedError`` -- there
is no CPU fallback.
"""
import ctypes as C
from typing import Optional, Union

import numpy as np
import torch

from . import _native as N
from . import _device as D
from . import _io

METHODS = {"sum": 0, "min": 1, "max": 2, "diff": 3, "mean": 4, "std": 5, "variance": 6, "median": 7,
           "mode": 8}
_UNSUPPORTED = ("percentile", "quantile")
MAX_KERNEL = 33  # largest kernel edge the direct kernel stages (halo of 16 cells)
RUNS_FROM = 64   # sum / mean with at least this many selected cells go through row prefix sums


def _as_kernel(kernel: Union[np.ndarray, tuple]) -> np.ndarray:
    # same checks and messages as interpolate.py:66-80
    if isinstance(kernel, tuple):
        if len(kernel) != 2:
            raise IndexError(f"Input kernel {kernel} does not have values for two dimensions")
        return np.ones(kernel, bool)
    kernel = np.asarray(kernel)
    if kernel.ndim != 2:
        raise IndexError(f"Input kernel of shape {kernel.shape} does not have two dimensions")
    return np.array(kernel, dtype=bool)


def _max_runs(k: np.ndarray) -> int:
    """Largest number of runs of selected cells in one kernel row."""
    p = np.pad(k.astype(np.int8), ((0, 0), (1, 0)))
    return int((np.diff(p, axis=1) == 1).sum(axis=1).max()) if k.size else 0


def window_filter(a: D.ArrayLike, kernel: Union[np.ndarray, tuple], method: str,
                  nodata: Optional[float] = None, halo_top: int = 0, halo_bot: int = 0):
    """Array-level ``raster_filter``: float32 (rows, cols) in, float32 out, NaN = no data
    (``nodata``, if given, is treated as NaN as well, like ``da.ma.filled(..., nan)`` at
    interpolate.py:66).  numpy in -> numpy out, CUDA tensor in -> CUDA tensor out.  With
    ``halo_top`` / ``halo_bot`` the first / last that many rows of ``a`` belong to
    neighbour bands (row-band sharding; pass at least ``kernel.shape[0] // 2`` rows)."""
    if method in _UNSUPPORTED or callable(method):
        raise NotImplementedError(f"reducer {method!r} is outside the CUDA path")
    if method not in METHODS:
        raise ValueError(f"unknown method {method!r}")
    k = _as_kernel(kernel)
    big = k.shape[0] > MAX_KERNEL or k.shape[1] > MAX_KERNEL
    # sum / mean over many cells (the discs of kernel_from_distance the reference filters
    # with): row prefix sums, any kernel size
    use_runs = method in ("sum", "mean") and (big or int(k.sum()) >= RUNS_FROM) and _max_runs(k) <= 64
    if big and not use_runs:
        raise ValueError(f"{method!r} over kernels larger than {MAX_KERNEL} x {MAX_KERNEL} is not "
                         f"supported (got {k.shape[0]} x {k.shape[1]}); sum and mean are")
    dev = D.require_cuda()
    on_host = not isinstance(a, torch.Tensor)
    a = D.as_2d(a)
    halo_top, halo_bot = int(halo_top), int(halo_bot)
    rows, cols = int(a.shape[0]) - halo_top - halo_bot, int(a.shape[1])
    if rows < 0:
        raise ValueError("raster smaller than its halos")
    buf, ld = D.upload(a, torch.float32, 4, dev)
    out, out_ld = D.empty_padded(rows, cols, torch.float32, dev, 4)
    kb = np.ascontiguousarray(k, dtype=np.uint8)
    has, nd = D.nodata_args(nodata)
    if use_runs:
        need = int(N.lib.ht_window_filter_runs_workspace_bytes(int(a.shape[0]), cols, int(k.shape[0])))
        ws = torch.empty(need + 256, dtype=torch.uint8, device=dev)
        N.call("ht_window_filter_runs_f32", buf.data_ptr() + halo_top * ld * 4, rows, cols, ld, has,
               nd, kb.ctypes.data_as(C.c_void_p), int(k.shape[0]), int(k.shape[1]), METHODS[method],
               out.data_ptr(), out_ld, halo_top, halo_bot, ws.data_ptr(), ws.numel(),
               D.stream_ptr())
        view = out[:, :cols]
        return D.download(view) if on_host else view
    N.call("ht_window_filter_f32", buf.data_ptr() + halo_top * ld * 4, rows, cols, ld, has, nd,
           kb.ctypes.data_as(C.c_void_p), int(k.shape[0]), int(k.shape[1]), METHODS[method],
           out.data_ptr(), out_ld, halo_top, halo_bot, D.stream_ptr())
    view = out[:, :cols]
    return D.download(view) if on_host else view


def raster_filter(raster_source: str, kernel: Union[np.ndarray, tuple], method: str,
                  filter_dst: str, *args, out_dtype: Optional[str] = None):
    """Filter a raster using a defined kernel and method.

    Same signature as /root/reference/src/hydrotools/interpolate.py:110-117; the output
    is float32 with the reference's float32 nodata (the type's minimum) unless ``out_dtype``
    is given."""
    if args:
        raise NotImplementedError("percentile / quantile arguments are outside the CUDA path")
    a, meta = _io.read_raster(raster_source)
    a = np.ma.getdata(a).astype(np.float32)
    res = window_filter(a, kernel, method, nodata=meta["nodata"])
    if out_dtype is not None:
        nd = _io.infer_nodata(out_dtype)
        filled = np.where(np.isnan(res), nd, res).astype(out_dtype)
        _io.write_raster(filled, raster_source, filter_dst, nodata=nd)
    else:
        # to_raster fills masked cells with infer_nodata(float32) (raster.py:807-811)
        nd = _io.infer_nodata("float32")
        _io.write_raster(np.where(np.isnan(res), np.float32(nd), res), raster_source, filter_dst, nodata=nd)


def custom_raster_filter(raster_source, kernel, func, filter_dst, *args, out_dtype=None):
    """Custom numba reducers cannot run inside a CUDA kernel: not implemented (the named
    reducers of ``raster_filter`` are)."""
    raise NotImplementedError("custom reducers are outside the CUDA path; use raster_filter")
