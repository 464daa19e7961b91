"""oracle/rasterfilter.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of the reference's moving-window filter
(/root/reference/src/hydrotools/interpolate.py:16-107 `custom_raster_filter`,
:110-247 `raster_filter`): for every non-NaN cell, the reducer of the cells the
boolean kernel selects in the window

    rows  i - i_start .. i + i_end,   i_start = ceil((kh - 1) / 2), i_end = kh - 1 - i_start
    cols  j - j_start .. j + j_end    (same with kw)

with NaN outside the raster (`da.map_overlap(..., boundary=np.nan)`,
interpolate.py:94-107), sample and reducer in float64 (`np.nansum`, `np.nanmin`,
`np.nanmax`, max - min, `np.nanmean`, `np.nanstd`, `np.nanvar`; interpolate.py:147-
222), result stored as float32, NaN where the centre is NaN or the sample is empty.

PARITY PINNED: `tools/make_filter_goldens.py` executes the reference's own numba
code (extracted from the file above) in the build container and commits its
outputs as `tests/golden/filter.npz`; `tests/test_rasterfilter.py` checks this
restatement against them (and against the live reference when it is present).
"""
import numpy as np

METHODS = ("sum", "min", "max", "diff", "mean", "std", "variance", "median", "mode")


def window_geometry(kh, kw):
    i_start = int(np.ceil((kh - 1) / 2.0))
    j_start = int(np.ceil((kw - 1) / 2.0))
    return i_start, kh - 1 - i_start, j_start, kw - 1 - j_start


def as_kernel(kernel):
    """tuple -> full rectangle; array -> boolean (interpolate.py:66-80)."""
    if isinstance(kernel, tuple):
        if len(kernel) != 2:
            raise IndexError(f"Input kernel {kernel} does not have values for two dimensions")
        return np.ones(kernel, bool)
    kernel = np.asarray(kernel)
    if kernel.ndim != 2:
        raise IndexError(f"Input kernel of shape {kernel.shape} does not have two dimensions")
    return kernel.astype(bool)


def raster_filter(a, kernel, method):
    """a: float32 (rows, cols) with NaN = no data.  Returns float32 (rows, cols)."""
    if method not in METHODS:
        raise ValueError(method)
    a = np.asarray(a, np.float32)
    k = as_kernel(kernel)
    kh, kw = k.shape
    i0, i1, j0, j1 = window_geometry(kh, kw)
    rows, cols = a.shape
    pad = np.full((rows + i0 + i1, cols + j0 + j1), np.nan, np.float64)
    pad[i0:i0 + rows, j0:j0 + cols] = a
    layers = [pad[ki:ki + rows, kj:kj + cols] for ki in range(kh) for kj in range(kw) if k[ki, kj]]
    out = np.full((rows, cols), np.nan, np.float32)
    if not layers:
        return out
    s = np.stack(layers)  # (n, rows, cols) float64, in the reference's sample order
    n = (~np.isnan(s)).sum(0)
    ok = ~np.isnan(a) & (n > 0)
    with np.errstate(all="ignore"):
        z = np.where(np.isnan(s), 0.0, s)
        if method in ("sum", "mean", "std", "variance"):
            tot = np.zeros((rows, cols))
            for layer in z:  # sequential float64 accumulation, like the numba loop
                tot = tot + layer
            if method == "sum":
                r = tot
            else:
                mean = tot / np.maximum(n, 1)
                if method == "mean":
                    r = mean
                else:
                    ss = np.zeros((rows, cols))
                    for layer in s:
                        d = np.where(np.isnan(layer), 0.0, layer - mean)
                        ss = ss + d * d
                    var = ss / np.maximum(n, 1)
                    r = var if method == "variance" else np.sqrt(var)
        elif method == "median":
            r = np.nanmedian(np.where(n[None] > 0, s, 0.0), axis=0)  # all-NaN windows are masked below
        elif method == "mode":
            # the reference's dictionary walk (interpolate.py:174-201): the first value, in
            # sample order, whose multiplicity is strictly the largest
            r = np.full((rows, cols), np.nan)
            best = np.zeros((rows, cols), np.int64)
            for layer in s:
                cnt = (s == layer[None]).sum(0)  # NaN never equals
                take = cnt > best
                r = np.where(take, layer, r)
                best = np.where(take, cnt, best)
        else:
            mn = np.where(np.isnan(s), np.inf, s).min(0)
            mx = np.where(np.isnan(s), -np.inf, s).max(0)
            r = mn if method == "min" else mx if method == "max" else mx - mn
    out[ok] = r[ok].astype(np.float32)
    return out
