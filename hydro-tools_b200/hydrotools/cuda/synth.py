"""Device-side synthetic workloads (BASELINE.json configs 2-5): the integer-exact
conditioned fractal DEM of include/ht_synth.h and precipitation-like weights."""
import torch

from . import _native as N
from . import _device as D


def synth_dem(seed: int, rows: int, cols: int, row0: int = 0, total_rows=None,
              with_nulls: bool = False, nodata: float = -32767.0,
              extra_top: int = 0, extra_bot: int = 0):
    """Generate rows [row0, row0+rows) of the synthetic DEM in HBM.

    Returns (buffer, ld): ``buffer`` is (extra_top + rows + extra_bot, ld) fp32;
    the extra rows (halo space for row-band sharding) are generated too when
    they fall inside the whole raster.
    """
    dev = D.require_cuda()
    total_rows = rows if total_rows is None else total_rows
    buf, ld = D.empty_padded(rows, cols, torch.float32, dev, 4, extra_top, extra_bot)
    r_lo = max(row0 - extra_top, 0)
    r_hi = min(row0 + rows + extra_bot, total_rows)
    first = r_lo - (row0 - extra_top)
    N.call("ht_synth_dem_f32", buf.data_ptr() + first * ld * 4, r_hi - r_lo, cols, ld,
           seed, r_lo, total_rows, int(with_nulls), float(nodata), D.stream_ptr())
    return buf, ld


def synth_weight(seed: int, rows: int, cols: int, row0: int = 0):
    dev = D.require_cuda()
    buf, ld = D.empty_padded(rows, cols, torch.float32, dev, 4)
    N.call("ht_synth_weight_f32", buf.data_ptr(), rows, cols, ld, seed, row0, D.stream_ptr())
    return buf, ld
