"""oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front end of the CPU oracle (``oracle/*.c``): restatements of GRASS
``r.watershed`` (ram, ``-s [-a]``), GRASS addon ``r.flowaccumulation`` and GDAL
``gdaldem slope|aspect|TRI -compute_edges`` -- the third-party engines the
reference shells out to for its hot path
(``/root/reference/src/hydrotools/watershed.py:64-75, 341-382``,
``/root/reference/src/hydrotools/elevation.py:41-54, 70-78, 118-126``).

PARITY UNPINNED: the reference's tests hold no golden output for this path and
neither GRASS nor GDAL can run in this image.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  ``hydrotools.cuda`` never
does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
DIR_NULL = -128

_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only)."""
    if force or not os.path.isfile(_LIB_PATH):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []),
                       check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
        i8p = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")
        f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
        L.ora_rwatershed_sfd.argtypes = [
            f32p, C.c_int64, C.c_int64, C.c_int, C.c_float, C.c_double,
            C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ora_check_conditioned.argtypes = [
            f32p, C.c_int64, C.c_int64, C.c_int, C.c_float,
            C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.ora_condition_by_rank.argtypes = [
            f32p, C.c_int64, C.c_int64, C.c_int, C.c_float, C.c_double,
            C.c_double, f32p]
        L.ora_gdaldem.argtypes = [
            f32p, C.c_int64, C.c_int64, C.c_int, C.c_float, C.c_double,
            C.c_double, C.c_double, C.c_int, f32p, f32p, f32p]
        L.ora_flowacc_count.argtypes = [i8p, C.c_int64, C.c_int64, f64p]
        L.ora_flowacc_weighted.argtypes = [
            i8p, f32p, C.c_int, C.c_float, C.c_int64, C.c_int64, f64p]
        L.ora_synth_dem.argtypes = [
            C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int,
            C.c_float, f32p]
        L.ora_synth_weight.argtypes = [
            C.c_uint64, C.c_int64, C.c_int64, C.c_int64, f32p]
        L.ora_fill_mm.argtypes = [f32p, C.c_int64, C.c_int64, C.c_int, C.c_float,
                                  np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")]
        L.ora_alt_from_f32.argtypes = [C.c_float]
        L.ora_alt_from_f32.restype = C.c_int32
        _lib = L
    return _lib


def _nd(nodata):
    return (0, 0.0) if nodata is None else (1, float(nodata))


def _dem(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2:
        raise ValueError("expected a 2-D array")
    return a


def rwatershed(dem, nodata=None, ewres=1.0, nsres=1.0, positive_only=True,
               edge_rule=True):
    """r.watershed -s [-a]: returns (drainage int8 [-128 = NULL], accumulation
    float64 [NaN = NULL])."""
    dem = _dem(dem)
    rows, cols = dem.shape
    has, nd = _nd(nodata)
    d = np.empty((rows, cols), np.int8)
    a = np.empty((rows, cols), np.float64)
    rc = lib().ora_rwatershed_sfd(dem, rows, cols, has, nd, ewres, nsres,
                                  int(positive_only), int(edge_rule),
                                  d.ctypes.data, a.ctypes.data, None)
    if rc:
        raise MemoryError("oracle r.watershed failed")
    return d, a


def check_conditioned(dem, nodata=None):
    """(n_pits, n_ties) as defined in SURVEY.md section 8(c)."""
    dem = _dem(dem)
    has, nd = _nd(nodata)
    p, t = C.c_int64(), C.c_int64()
    lib().ora_check_conditioned(dem, dem.shape[0], dem.shape[1], has, nd,
                                C.byref(p), C.byref(t))
    return p.value, t.value


def condition_by_rank(dem, nodata=None, ewres=1.0, nsres=1.0):
    """Pit-free, tie-free version of ``dem`` (elevation := A* pop rank, mm)."""
    dem = _dem(dem)
    has, nd = _nd(nodata)
    out = np.empty_like(dem)
    rc = lib().ora_condition_by_rank(dem, dem.shape[0], dem.shape[1], has, nd,
                                     ewres, nsres, out)
    if rc:
        raise MemoryError("oracle conditioning failed")
    return out


ALT_NULL = 0x7FFFFFFF


def fill(dem, nodata=None):
    """Priority-flood fill in GRASS integer millimetres, seeded like r.watershed's A*
    (border and NULL-adjacent cells).  Returns (W int32 [ALT_NULL = NULL], filled DEM
    float32 in metres with NULLs kept)."""
    dem = _dem(dem)
    has, nd = _nd(nodata)
    w = np.empty(dem.shape, np.int32)
    if lib().ora_fill_mm(dem, dem.shape[0], dem.shape[1], has, nd, w):
        raise MemoryError("oracle fill failed")
    out = (w.astype(np.float64) / 1000.0).astype(np.float32)
    null = w == ALT_NULL
    out[null] = dem[null]
    return w, out


def gdaldem(dem, nodata=None, ewres=1.0, nsres=1.0, scale=1.0, percent=False, threads=1):
    """gdaldem slope/aspect/TRI -compute_edges: (slope, aspect, tri) float32,
    nodata -9999.  ``threads`` > 1: rows in parallel (pthreads), for the all-core CPU
    baseline; gdaldem itself is single-threaded."""
    dem = _dem(dem)
    lib().ora_set_threads(int(threads))
    has, nd = _nd(nodata)
    s, a, t = (np.empty_like(dem) for _ in range(3))
    lib().ora_gdaldem(dem, dem.shape[0], dem.shape[1], has, nd, ewres, nsres,
                      scale, int(percent), s, a, t)
    return s, a, t


def flowacc(direction, weight=None, weight_nodata=None):
    """r.flowaccumulation: cell counts, or weighted sums when ``weight``."""
    d = np.ascontiguousarray(direction, dtype=np.int8)
    out = np.empty(d.shape, np.float64)
    if weight is None:
        lib().ora_flowacc_count(d, d.shape[0], d.shape[1], out)
    else:
        w = _dem(weight)
        has, nd = _nd(weight_nodata)
        lib().ora_flowacc_weighted(d, w, has, nd, d.shape[0], d.shape[1], out)
    return out


def synth_dem(seed, rows, cols, row0=0, total_rows=None, with_nulls=False,
              nodata=-32767.0):
    out = np.empty((rows, cols), np.float32)
    lib().ora_synth_dem(seed, rows, cols, row0,
                        rows if total_rows is None else total_rows,
                        int(with_nulls), nodata, out)
    return out


def synth_weight(seed, rows, cols, row0=0):
    out = np.empty((rows, cols), np.float32)
    lib().ora_synth_weight(seed, rows, cols, row0, out)
    return out
