"""CPU: the branch of hydrotools.cuda._io a real hydro-tools install takes -- reading with
hydrotools.raster.from_raster and writing with hydrotools.raster.to_raster
(/root/reference/src/hydrotools/raster.py:773-791, 794-846) -- driven through stubs of
`hydrotools.raster`, `hydrotools.config` and `dask.array` that keep the reference's call
contract (masked (1, rows, cols) array with .compute(); to_raster(data, template,
destination, as_cog=..., nodata_value=...); Raster.matches).  rasterio / dask / GDAL are not
in this image, so the real functions cannot be imported; what is checked is that the
product calls them the way the reference defines them."""
import importlib
import sys
import types

import numpy as np
import pytest


class FakeLazy:
    """Stand-in for a dask array: numpy-backed, materialised by .compute()."""

    def __init__(self, a):
        self.a = a
        self.dtype = a.dtype
        self.ndim = a.ndim
        self.shape = a.shape

    def compute(self):
        return self.a

    def __eq__(self, other):  # `arr == nodata` inside the product's write branch
        return FakeLazy(np.ma.getdata(self.a) == other)


@pytest.fixture
def ref_io(tmp_path):
    store = {}   # path -> (array (1, rows, cols), nodata)
    calls = []

    class Raster:
        def __init__(self, src):
            self.src = src
            a, nd = store[src]
            self.nodata = nd
            self.csx, self.csy, self.left, self.top = 5.0, 5.0, 1000.0, 2000.0
            self.shape = a.shape

        def matches(self, other, tolerance=1e-6):
            calls.append(("matches", self.src, other, tolerance))
            return store[other][0].shape == self.shape

    def from_raster(src, chunks=None):
        r = src if isinstance(src, Raster) else Raster(src)
        a, nd = store[r.src]
        calls.append(("from_raster", r.src))
        m = np.ma.masked_where((a == nd) | np.isnan(a.astype(np.float64)), a)
        m.fill_value = np.finfo("float32").min if a.dtype.kind == "f" else np.iinfo(a.dtype).min
        return FakeLazy(m)

    def to_raster(data, template, destination, as_cog=True, **kwargs):
        calls.append(("to_raster", template, destination, as_cog, dict(kwargs)))
        a = data.compute() if hasattr(data, "compute") else data
        assert a.ndim == 3 and a.shape[0] == 1, "the reference stores (bands, rows, cols)"
        nd = kwargs.get("nodata_value")
        store[destination] = (np.ma.filled(a, nd), nd)

    raster_mod = types.ModuleType("hydrotools.raster")
    raster_mod.Raster, raster_mod.from_raster, raster_mod.to_raster = Raster, from_raster, to_raster
    config_mod = types.ModuleType("hydrotools.config")
    config_mod.CHUNKS = (1, 4096, 4096)
    dask_mod = types.ModuleType("dask")
    da_mod = types.ModuleType("dask.array")
    da_mod.from_array = lambda a, chunks=None: FakeLazy(np.asarray(a))
    ma_mod = types.SimpleNamespace(
        masked_where=lambda cond, a: FakeLazy(np.ma.masked_where(cond.a if isinstance(cond, FakeLazy) else cond,
                                                                a.a if isinstance(a, FakeLazy) else a)))
    da_mod.ma = ma_mod
    dask_mod.array = da_mod
    saved = {k: sys.modules.get(k) for k in ("hydrotools.raster", "hydrotools.config", "dask", "dask.array")}
    sys.modules.update({"hydrotools.raster": raster_mod, "hydrotools.config": config_mod,
                        "dask": dask_mod, "dask.array": da_mod})
    from hydrotools.cuda import _io
    importlib.reload(_io)
    assert _io.HAVE_REFERENCE_IO
    yield _io, store, calls
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v
    importlib.reload(_io)
    assert not _io.HAVE_REFERENCE_IO


def test_read_goes_through_from_raster(hc, ref_io):
    _io, store, calls = ref_io
    a = np.arange(12, dtype=np.float32).reshape(1, 3, 4)
    a[0, 1, 2] = -32767.0
    a[0, 2, 3] = np.nan
    store["dem.tif"] = (a, -32767.0)
    data, meta = _io.read_raster("dem.tif")
    assert ("from_raster", "dem.tif") in calls
    assert data.shape == (3, 4) and not isinstance(data, np.ma.MaskedArray)
    # masked cells (nodata AND NaN, raster.py:788) come back holding the nodata value
    assert data[1, 2] == -32767.0 and data[2, 3] == -32767.0 and data[0, 1] == 1.0
    assert meta["nodata"] == -32767.0 and meta["csx"] == 5.0 and meta["csy"] == 5.0
    assert meta["left"] == 1000.0 and meta["top"] == 2000.0 and meta["shape"] == (3, 4)


def test_write_goes_through_to_raster(hc, ref_io):
    _io, store, calls = ref_io
    store["tmpl.tif"] = (np.zeros((1, 3, 4), np.float32), -32767.0)
    out = np.array([[1, 2, -128, 4], [5, 6, 7, 8], [-128, 1, 2, 3]], np.int8)
    _io.write_raster(out, "tmpl.tif", "fd.tif", nodata=-128)
    kind, template, dst, as_cog, kw = [c for c in calls if c[0] == "to_raster"][-1]
    assert (template, dst, as_cog) == ("tmpl.tif", "fd.tif", True)
    assert kw == {"nodata_value": -128}
    stored, nd = store["fd.tif"]
    assert stored.shape == (1, 3, 4) and nd == -128
    np.testing.assert_array_equal(stored[0], out)
    # without a nodata the reference infers one: no keyword is passed
    _io.write_raster(out.astype(np.float32), "tmpl.tif", "x.tif", as_cog=False)
    kind, template, dst, as_cog, kw = [c for c in calls if c[0] == "to_raster"][-1]
    assert as_cog is False and kw == {}


def test_same_grid_is_raster_matches(hc, ref_io):
    _io, store, calls = ref_io
    store["a.tif"] = (np.zeros((1, 3, 4), np.float32), None)
    store["b.tif"] = (np.zeros((1, 3, 4), np.float32), None)
    store["c.tif"] = (np.zeros((1, 2, 4), np.float32), None)
    assert _io.same_grid("a.tif", "b.tif") and not _io.same_grid("a.tif", "c.tif")
    assert ("matches", "a.tif", "b.tif", 1e-6) in calls
