"""B200 replacements for ``hydrotools.watershed.flow_direction_accumulation`` and
``hydrotools.watershed.FlowAccumulation``
(/root/reference/src/hydrotools/watershed.py:30-75, 321-409).

The reference shells out to GRASS ``r.watershed -s [-a]`` and the addon
``r.flowaccumulation``; here the DEM goes through the sm_100a kernels behind
include/hydrotools_b200.h: ``ht_d8_f32`` (direction codes), ``ht_acc_*``
(tile-local propagation, boundary-graph resolve, final pass).

Scope (BASELINE.json north_star): single flow direction.  Hydrologically conditioned
DEMs take the one-pass local-rule kernel; every other DEM (equal elevations, flats,
depressions -- e.g. the reference's own test fixture) the general kernel
(``ht_d8_general_f32``), which reproduces r.watershed's least-cost search exactly.
``single=False`` (MFD) and ``beautify=True`` raise ``NotImplementedError``; there is no
CPU fallback to GRASS.
"""
import os
from tempfile import TemporaryDirectory
from typing import Optional, Tuple, Union

import numpy as np
import torch

from . import _native as N
from . import _device as D
from . import _io

DIR_NODATA = -128  # hydrotools.utils.infer_nodata("int8")


def _workspace(rows: int, cols: int, dev, mode: int = 0) -> torch.Tensor:
    return torch.empty(int(N.lib.ht_acc_workspace_bytes_mode(rows, cols, mode)), dtype=torch.uint8,
                       device=dev)


def validate_conditioned(dem: D.ArrayLike, nodata: Optional[float] = None) -> Tuple[int, int]:
    """(pits, ties) of a DEM after GRASS integer-millimetre scaling; both must be
    0 for the D8 kernel to reproduce r.watershed."""
    dev = D.require_cuda()
    buf, ld = D.upload(dem, torch.float32, 4, dev)
    rows, cols = D.as_2d(dem).shape
    return _validate(buf, ld, int(rows), int(cols), nodata, dev)


def _validate(buf, ld, rows, cols, nodata, dev):
    counts = torch.zeros(3, dtype=torch.int64, device=dev)
    has, nd = D.nodata_args(nodata)
    N.call("ht_d8_validate_conditioned_f32", buf.data_ptr(), rows, cols, ld, has, nd, 0, rows,
           0, 0, counts.data_ptr(), D.stream_ptr())
    c = counts.cpu()
    if int(c[2]):
        raise ValueError(f"{int(c[2])} cells have |elevation| >= 2**24 mm (16 777 m); not supported")
    return int(c[0]), int(c[1])


def d8(dem: torch.Tensor, ld: int, rows: int, cols: int, nodata: Optional[float],
       csx: float, csy: float) -> Tuple[torch.Tensor, int]:
    """Packed direction bytes (device) of a pitch-aligned device DEM."""
    dev = dem.device
    packed, pld = D.empty_padded(rows, cols, torch.uint8, dev, 16)
    has, nd = D.nodata_args(nodata)
    N.call("ht_d8_f32", dem.data_ptr(), rows, cols, ld, has, nd, float(abs(csx)),
           float(abs(csy)), 0, rows, 0, 0, packed.data_ptr(), pld, D.stream_ptr())
    return packed, pld


def d8_general(dem: torch.Tensor, ld: int, rows: int, cols: int, nodata: Optional[float],
               csx: float, csy: float) -> Tuple[torch.Tensor, int, dict]:
    """Packed direction bytes of ANY DEM (ties, flats, depressions): r.watershed's
    least-cost search reproduced exactly (csrc/ht_astar.cu).  Returns (packed, pitch, stats)."""
    import ctypes
    dev = dem.device
    packed, pld = D.empty_padded(rows, cols, torch.uint8, dev, 16)
    ws = torch.empty(int(N.lib.ht_d8_general_workspace_bytes(rows, cols)), dtype=torch.uint8, device=dev)
    has, nd = D.nodata_args(nodata)
    stats = (ctypes.c_longlong * 8)()
    N.call("ht_d8_general_f32", dem.data_ptr(), rows, cols, ld, has, nd, float(abs(csx)),
           float(abs(csy)), packed.data_ptr(), pld, ws.data_ptr(), ws.numel(), stats, D.stream_ptr())
    names = ("rounds", "episodes", "largest_episode", "depression_cells", "fill_passes",
             "generation_sweeps", "out_of_range", "errors")
    info = dict(zip(names, (int(v) for v in stats)))
    if info["out_of_range"]:
        raise ValueError(f"{info['out_of_range']} cells have |elevation| >= 2**24 mm (16 777 m); not supported")
    return packed, pld, info


LAST_STATS = {}  # statistics of the last general-path run (diagnostics)


def accumulate(packed: torch.Tensor, pld: int, rows: int, cols: int, mode: int,
               weight: Optional[torch.Tensor] = None, wld: int = 0,
               weight_nodata: Optional[float] = None, want_dir: bool = False,
               ws: Optional[torch.Tensor] = None):
    """Stages A+B+C on one device.  Returns (acc fp64 [rows, ald], ald, dir int8 or None, dld)."""
    dev = packed.device
    acc, ald = D.empty_padded(rows, cols, torch.float64, dev, 2)
    dout, dld = (D.empty_padded(rows, cols, torch.int8, dev, 16) if want_dir else (None, 0))
    if ws is None:
        ws = _workspace(rows, cols, dev, mode)
    has, nd = D.nodata_args(weight_nodata)
    N.call("ht_acc_f64", packed.data_ptr(), pld, rows, cols, mode,
           weight.data_ptr() if weight is not None else None, wld, has, nd,
           acc.data_ptr(), ald, dout.data_ptr() if dout is not None else None, dld,
           ws.data_ptr(), ws.numel(), D.stream_ptr())
    return acc, ald, dout, dld


def flow_direction_accumulation_arrays(dem: D.ArrayLike, nodata: Optional[float] = None,
                                       csx: float = 1.0, csy: float = 1.0,
                                       positive_only: bool = True, validate: bool = True,
                                       out_direction: Optional[np.ndarray] = None,
                                       out_accumulation: Optional[np.ndarray] = None,
                                       method: str = "auto"):
    """Array-level ``flow_direction_accumulation``.

    Returns ``(direction int8, accumulation float64)``: GRASS codes 1..8 (1 = NE,
    counter-clockwise), negative = flows off the region / into NULL, -128 where
    the DEM is NULL; accumulation = upslope cell count incl. the cell, NaN where
    NULL, negative "likely underestimates" unless ``positive_only``.
    numpy in -> numpy out; CUDA tensor in -> CUDA tensors out.  ``out_direction``
    / ``out_accumulation`` (host arrays, ideally page-locked) receive the results
    without an intermediate copy.

    ``method``: "auto" counts pits and ties first and takes the one-pass local-rule kernel
    when the DEM is hydrologically conditioned (pit-free, tie-free after GRASS's millimetre
    rounding), else the general kernel, which reproduces r.watershed's search on any DEM
    (equal elevations, flats, depressions) in parallel rounds; "local" / "general" force one
    ("local" raises ``ValueError`` on an unconditioned DEM unless ``validate=False``).
    """
    dev = D.require_cuda()
    on_host = not isinstance(dem, torch.Tensor)
    a = D.as_2d(dem)
    rows, cols = int(a.shape[0]), int(a.shape[1])
    if rows == 0 or cols == 0:
        raise ValueError("empty raster")
    if method not in ("auto", "local", "general"):
        raise ValueError(f"unknown method {method!r}")
    buf, ld = D.upload(a, torch.float32, 4, dev)
    general = method == "general"
    if method != "general" and (validate or method == "auto"):
        pits, ties = _validate(buf, ld, rows, cols, nodata, dev)
        if pits or ties:
            if method == "local":
                raise ValueError(
                    f"DEM is not hydrologically conditioned ({pits} pits, {ties} cells with "
                    "tied elevations after GRASS mm rounding): the local-rule kernel does "
                    "not apply; use method='auto' or 'general'")
            general = True
    if general:
        packed, pld, info = d8_general(buf, ld, rows, cols, nodata, csx, csy)
        LAST_STATS.clear()
        LAST_STATS.update(info)
    else:
        packed, pld = d8(buf, ld, rows, cols, nodata, csx, csy)
    ws = _workspace(rows, cols, dev)
    acc, ald, dout, dld = accumulate(packed, pld, rows, cols, N.MODE_COUNT, want_dir=True, ws=ws)
    if not positive_only:
        taint, tld, _, _ = accumulate(packed, pld, rows, cols, N.MODE_TAINT, ws=ws)
        N.call("ht_acc_apply_sign", acc.data_ptr(), ald, taint.data_ptr(), tld,
               packed.data_ptr(), pld, rows, cols, D.stream_ptr())
    dview, aview = dout[:, :cols], acc[:, :cols]
    if on_host:
        return D.download(dview, out_direction), D.download(aview, out_accumulation)
    return dview, aview


def flow_direction_accumulation(
    dem: str,
    direction_grid: str,
    accumulation_grid: str,
    single: bool = True,
    positive_only: bool = True,
    memory: Union[int, None] = None,
    beautify: Union[bool, None] = None,
):
    """Generate Flow Direction and Flow Accumulation grids on the GPU.

    Same signature as /root/reference/src/hydrotools/watershed.py:30-38.
    ``memory`` is accepted and ignored (GRASS's disk-swap mode has no meaning
    here); ``single=False`` and ``beautify=True`` are not implemented.
    """
    if not single:
        raise NotImplementedError("multiple flow direction (single=False) is outside the CUDA path")
    if beautify:
        raise NotImplementedError("beautify (r.watershed -b) is outside the CUDA path")
    a, meta = _io.read_raster(dem)
    if _io.device_io():
        # results stay in HBM: their 512 x 512 tiles are deflated there and only the
        # compressed bytes cross PCIe (hydrotools.cuda._cog)
        dev = D.require_cuda()
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
        fd, fa = flow_direction_accumulation_arrays(
            t, nodata=meta["nodata"], csx=meta["csx"], csy=meta["csy"], positive_only=positive_only)
    else:
        fd, fa = flow_direction_accumulation_arrays(
            a, nodata=meta["nodata"], csx=meta["csx"], csy=meta["csy"], positive_only=positive_only)
    _io.write_raster(fd, dem, direction_grid, nodata=DIR_NODATA)
    _io.write_raster(fa, dem, accumulation_grid, nodata=float("nan"))


def _dir_to_int8(a: np.ndarray, nodata) -> np.ndarray:
    """Direction raster from disk (any integer/float type) -> int8 codes, -128 NULL."""
    a = np.ma.getdata(a)
    null = np.zeros(a.shape, bool)
    if nodata is not None and not (isinstance(nodata, float) and np.isnan(nodata)):
        null |= a == nodata
    if a.dtype.kind == "f":
        null |= np.isnan(a)
        a = np.where(null, 0, a)
    out = np.clip(a, -8, 8).astype(np.int8)
    out[null] = DIR_NODATA
    return out


def flow_accumulation_arrays(direction: D.ArrayLike, weight: Optional[D.ArrayLike] = None,
                             weight_nodata: Optional[float] = None):
    """r.flowaccumulation on arrays: ``direction`` int8 GRASS codes (-128 NULL);
    returns float64 cell counts, or weighted sums when ``weight`` is given."""
    dev = D.require_cuda()
    on_host = not isinstance(direction, torch.Tensor)
    d = D.as_2d(direction)
    rows, cols = int(d.shape[0]), int(d.shape[1])
    dbuf, dld = D.upload(d, torch.int8, 16, dev)
    packed, pld = D.empty_padded(rows, cols, torch.uint8, dev, 16)
    N.call("ht_dir_pack_i8", dbuf.data_ptr(), dld, rows, cols, packed.data_ptr(), pld,
           D.stream_ptr())
    if weight is None:
        acc, ald, _, _ = accumulate(packed, pld, rows, cols, N.MODE_COUNT)
    else:
        w = D.as_2d(weight)
        if tuple(w.shape) != (rows, cols):
            raise ValueError("Input raster does not align with direction grid")
        wbuf, wld = D.upload(w, torch.float32, 4, dev)
        acc, ald, _, _ = accumulate(packed, pld, rows, cols, N.MODE_WEIGHTED, wbuf, wld,
                                    weight_nodata)
    view = acc[:, :cols]
    return D.download(view) if on_host else view


class FlowAccumulation:
    """GPU version of ``hydrotools.watershed.FlowAccumulation``
    (/root/reference/src/hydrotools/watershed.py:321-409): sums and means of raster
    parameters over the upslope area of every cell."""

    def __init__(self, direction: str):
        """
        Args:
            direction (str): Path to a flow direction raster (GRASS 1..8 codes)
        """
        self.direction = direction
        self.wkdir = TemporaryDirectory()
        self._modals = None
        self._dir, self._meta = None, None

    def _direction(self):
        if self._dir is None:
            a, self._meta = _io.read_raster(self.direction)
            self._dir = _dir_to_int8(a, self._meta["nodata"])
        return self._dir

    @property
    def modals(self):
        if self._modals is None:
            self._modals = os.path.join(self.wkdir.name, "modals.tif")
            cnt = flow_accumulation_arrays(self._direction()).astype(np.float32)
            _io.write_raster(cnt, self.direction, self._modals, nodata=float("nan"), as_cog=False)
        return self._modals

    def calculate(self, raster_source: str, dst: str, method: str = "sum"):
        """Calculate flow accumulation statistics ("sum" or "mean")."""
        if not _io.same_grid(raster_source, self.direction):
            raise ValueError("Input raster does not align with direction grid")
        w, meta = _io.read_raster(raster_source)
        total = flow_accumulation_arrays(self._direction(), w.astype(np.float32, copy=False),
                                         meta["nodata"])
        if method == "sum":
            _io.write_raster(total, self.direction, dst, nodata=float("nan"))
            return
        modals, _ = _io.read_raster(self.modals)
        with np.errstate(invalid="ignore", divide="ignore"):
            avg = (total / modals).astype(np.float32)
        nd = float(np.finfo("float32").min)  # infer_nodata("float32")
        avg[(modals == 0) | np.isnan(avg)] = nd
        _io.write_raster(avg, self.direction, dst, nodata=nd)

    def contributing_area(self, dst: str):
        """Save the contributing area raster in km**2."""
        modals, meta = _io.read_raster(self.modals)
        # float32 like the reference (the modals raster is float32, watershed.py:341-355, 392-405)
        ca = ((modals.astype(np.float64) * meta["csx"] * meta["csy"]) / 1e6).astype(np.float32)
        nd = float(np.finfo("float32").min)
        ca[(ca == 0) | np.isnan(ca)] = nd
        _io.write_raster(ca, self.direction, dst, nodata=nd)

    def __del__(self):
        self.wkdir.cleanup()
