"""B200 replacements for ``hydrotools.elevation.slope`` / ``aspect`` /
``terrain_ruggedness_index`` (/root/reference/src/hydrotools/elevation.py:24-56,
59-80, 106-128).  Same signatures; the three gdaldem subprocesses become one
fused sm_100a stencil (``ht_terrain_fused_f32``) over a single read of the DEM.
"""
from typing import Dict, Iterable, Optional

import torch

from . import _native as N
from . import _device as D
from . import _io

TERRAIN_NODATA = N.HT_TERRAIN_NODATA
_OUTPUTS = ("slope", "aspect", "tri")


def terrain(dem: D.ArrayLike, nodata: Optional[float] = None, csx: float = 1.0,
            csy: float = 1.0, scale: float = 1.0, units: str = "degrees",
            outputs: Iterable[str] = _OUTPUTS, halo_top: int = 0,
            halo_bot: int = 0) -> Dict[str, D.ArrayLike]:
    """Array-level fused terrain derivatives.

    ``dem`` is a numpy array (results come back as numpy) or a CUDA
    ``torch.Tensor`` (results stay in HBM).  With ``halo_top`` / ``halo_bot`` the
    first / last row of ``dem`` is a neighbour band's row and is not computed
    (row-band sharding).  Output nodata is gdaldem's -9999.
    """
    outputs = tuple(outputs)
    for o in outputs:
        if o not in _OUTPUTS:
            raise ValueError(f"unknown output {o!r}")
    if not outputs:
        raise ValueError("no outputs requested")
    if units.lower() not in ("degrees", "percent"):
        raise ValueError('units must be "degrees" or "percent"')
    dev = D.require_cuda()
    on_host = not isinstance(dem, torch.Tensor)
    dem = D.as_2d(dem)
    halo_top, halo_bot = int(bool(halo_top)), int(bool(halo_bot))
    rows = int(dem.shape[0]) - halo_top - halo_bot
    cols = int(dem.shape[1])
    if rows < 0:
        raise ValueError("raster smaller than its halos")
    buf, ld = D.upload(dem, torch.float32, 4, dev)
    outs = {}
    ptrs = {}
    out_ld = 0
    for o in _OUTPUTS:
        if o in outputs:
            t, out_ld = D.empty_padded(rows, cols, torch.float32, dev, 4)
            outs[o] = t
            ptrs[o] = t.data_ptr()
        else:
            ptrs[o] = None
    has, nd = D.nodata_args(nodata)
    first_owned = buf.data_ptr() + halo_top * ld * 4
    N.call("ht_terrain_fused_f32", first_owned, rows, cols, ld, has, nd,
           float(abs(csx)), float(abs(csy)), float(scale),
           int(units.lower() == "percent"), ptrs["slope"], ptrs["aspect"], ptrs["tri"],
           out_ld, halo_top, halo_bot, D.stream_ptr())
    res = {}
    for o, t in outs.items():
        view = t[:, :cols]
        res[o] = D.download(view) if on_host else view
    return res


def _run_to_file(dem: str, destination: str, what: str, **kw):
    a, meta = _io.read_raster(dem)
    out = terrain(a, nodata=meta["nodata"], csx=meta["csx"], csy=meta["csy"],
                  outputs=(what,), **kw)[what]
    _io.write_raster(out, dem, destination, nodata=TERRAIN_NODATA)


def slope(dem: str, destination: str, units: str = "degrees", scale: float = 1):
    """Calculate topographic slope (Horn), as ``gdaldem slope -compute_edges``.

    Args mirror /root/reference/src/hydrotools/elevation.py:24-29.
    """
    _run_to_file(dem, destination, "slope", units=units, scale=scale)


def aspect(dem: str, destination: str):
    """Calculate aspect (azimuth, 0 = north, clockwise; flat = nodata), as
    ``gdaldem aspect -compute_edges``
    (/root/reference/src/hydrotools/elevation.py:59-62)."""
    _run_to_file(dem, destination, "aspect")


def terrain_ruggedness_index(dem: str, destination: str):
    """Calculate the Terrain Ruggedness Index (Riley), as
    ``gdaldem TRI -compute_edges``
    (/root/reference/src/hydrotools/elevation.py:106-109)."""
    _run_to_file(dem, destination, "tri")
