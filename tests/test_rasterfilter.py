"""Moving-window filter (reference:
/root/reference/src/hydrotools/interpolate.py:16-247).

PARITY PINNED: tests/golden/filter.npz holds outputs of the reference's own numba
code (tools/make_filter_goldens.py cuts it out of the reference file and runs it
unchanged).  CPU tests pin the numpy restatement (oracle/rasterfilter.py) to them;
GPU tests compare the CUDA kernel with both."""
import os

import numpy as np
import pytest

from oracle import rasterfilter as orf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METHODS = ["sum", "min", "max", "diff", "mean", "std", "variance", "median", "mode"]
RTOL = 1e-6  # fp64 accumulation on both sides, float32 results


@pytest.fixture(scope="module")
def fgold():
    z = np.load(os.path.join(ROOT, "tests", "golden", "filter.npz"))
    return {k: z[k] for k in z.files}


def _cases(fgold):
    return sorted({k.split("__")[0] for k in fgold})


def _close(got, exp, what):
    assert got.dtype == np.float32
    np.testing.assert_array_equal(np.isnan(got), np.isnan(exp), err_msg=what)
    m = ~np.isnan(exp)
    if m.any():
        np.testing.assert_allclose(got[m], exp[m], rtol=RTOL,
                                   atol=1e-6 * float(np.abs(exp[m]).max() + 1), err_msg=what)


def test_restatement_matches_reference_goldens(fgold):
    for c in _cases(fgold):
        a, k = fgold[f"{c}__in"], fgold[f"{c}__kernel"]
        for m in METHODS:
            if f"{c}__{m}" in fgold:  # mode: the class rasters only (interpreted reference)
                _close(orf.raster_filter(a, k, m), fgold[f"{c}__{m}"], f"{c} {m}")
    assert "classes_disc3__mode" in fgold and "classes_4x5__median" in fgold


def test_window_geometry():
    # interpolate.py:82-85
    assert orf.window_geometry(3, 3) == (1, 1, 1, 1)
    assert orf.window_geometry(4, 2) == (2, 1, 1, 0)
    assert orf.window_geometry(1, 5) == (0, 0, 2, 2)
    with pytest.raises(IndexError):
        orf.as_kernel((3,))
    with pytest.raises(IndexError):
        orf.as_kernel(np.ones(3, bool))


@pytest.mark.skipif(not os.path.isfile("/root/reference/src/hydrotools/interpolate.py"),
                    reason="reference sources only exist in the build container")
def test_restatement_matches_live_reference():
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "make_filter_goldens", os.path.join(ROOT, "tools", "make_filter_goldens.py"))
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    apply_filter, reducers = g.reference_functions()
    rng = np.random.default_rng(5)
    a = (rng.random((23, 31)) * 50 - 10).astype(np.float32)
    a[rng.random(a.shape) < 0.2] = np.nan
    for kernel in [(3, 3), (2, 5), g.disc(2)]:
        for m in METHODS:
            exp = g.reference_filter(apply_filter, reducers[m], a, kernel)
            _close(orf.raster_filter(a, kernel, m), exp, f"{kernel} {m}")


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gpu_matches_reference_goldens(hc, fgold):
    for c in _cases(fgold):
        a, k = fgold[f"{c}__in"], fgold[f"{c}__kernel"]
        for m in METHODS:
            if f"{c}__{m}" in fgold:
                _close(hc.window_filter(a, k, m), fgold[f"{c}__{m}"], f"{c} {m}")


@pytest.mark.gpu
@pytest.mark.parametrize("shape,kernel", [
    ((1, 1), (3, 3)), ((5, 300), (3, 3)), ((300, 5), (1, 7)), ((33, 129), (6, 4)),
    ((257, 513), (3, 3)), ((700, 900), (6, 4)), ((257, 513), "ring"), ((130, 140), "disc9"),
    ((70, 150), (33, 33)), ((5, 300), (33, 33)), ((300, 5), "disc9"), ((1, 1), (33, 33)),
])
def test_gpu_matches_restatement(hc, shape, kernel):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    a = (rng.random(shape) * 1000).astype(np.float32)
    a[rng.random(shape) < 0.15] = np.nan
    if kernel == "disc9":
        y, x = np.mgrid[-9:10, -9:10]
        kernel = (x * x + y * y) <= 81
    elif kernel == "ring":
        kernel = np.ones((5, 5), bool)
        kernel[1:4, 1:4] = False
    classes = np.round(a / 200.0)  # repeated values, so that mode and median ties mean something
    for m in METHODS:
        src = classes if m == "mode" else a
        _close(hc.window_filter(src, kernel, m), orf.raster_filter(src, kernel, m), f"{shape} {m}")
    _close(hc.window_filter(classes, kernel, "median"), orf.raster_filter(classes, kernel, "median"),
           f"{shape} median of classes")


@pytest.mark.gpu
@pytest.mark.parametrize("shape,radius", [((300, 400), 20), ((150, 1100), 40), ((64, 64), 50), ((1, 1), 17)])
def test_gpu_large_discs_by_row_prefix_sums(hc, shape, radius):
    """sum / mean over distance discs of any size (kernel_from_distance,
    /root/reference/src/hydrotools/utils.py:408-426): the row-prefix-sum kernel."""
    rng = np.random.default_rng(radius)
    a = (rng.random(shape) * 1000).astype(np.float32)
    a[rng.random(shape) < 0.2] = np.nan
    y, x = np.mgrid[-radius:radius + 1, -radius:radius + 1]
    disc = (x * x + y * y) <= radius * radius
    disc[radius, radius] = False  # kernel_from_distance clears nothing, but a hole makes two runs
    for m in ("sum", "mean"):
        _close(hc.window_filter(a, disc, m), orf.raster_filter(a, disc, m), f"{shape} r={radius} {m}")
    # even-sized rectangle, and the same answer as the direct kernel where both apply
    rect = np.ones((12, 20), bool)
    _close(hc.window_filter(a, rect, "mean"), orf.raster_filter(a, rect, "mean"), "rect")


@pytest.mark.gpu
def test_gpu_nodata_value_and_device_tensors(hc):
    import torch
    rng = np.random.default_rng(1)
    a = (rng.random((200, 333)) * 10).astype(np.float32)
    hole = rng.random(a.shape) < 0.1
    b = a.copy(); b[hole] = -9999.0
    a[hole] = np.nan
    exp = orf.raster_filter(a, (5, 5), "mean")
    _close(hc.window_filter(b, (5, 5), "mean", nodata=-9999.0), exp, "nodata value")
    out = hc.window_filter(torch.from_numpy(a).cuda(), (5, 5), "mean")
    assert out.is_cuda
    _close(out.cpu().numpy(), exp, "device tensor")


@pytest.mark.gpu
def test_gpu_limits_and_errors(hc):
    a = np.zeros((10, 10), np.float32)
    with pytest.raises(NotImplementedError):
        hc.window_filter(a, (3, 3), "percentile")
    with pytest.raises(ValueError):
        hc.window_filter(a, (35, 3), "max")  # larger than the 33 x 33 the direct kernel stages
    with pytest.raises(IndexError):
        hc.window_filter(a, (3,), "sum")


@pytest.mark.gpu
def test_gpu_path_level(hc, fgold, tmp_path):
    from hydrotools.cuda import _io
    from test_terrain_gpu import _write_template
    a = fgold["dem_5x5__in"]
    tmpl, src, dst = str(tmp_path / "t.tif"), str(tmp_path / "a.tif"), str(tmp_path / "f.tif")
    nd = -32767.0
    filled = np.where(np.isnan(a), nd, a).astype(np.float32)
    _write_template(tmpl, filled, nd)
    _io.write_raster(filled, tmpl, src, nodata=nd)
    hc.raster_filter(src, (5, 5), "std", dst)
    got, meta = _io.read_raster(dst)
    got = np.where(got == meta["nodata"], np.nan, got) if meta["nodata"] == meta["nodata"] else got
    _close(got.astype(np.float32), fgold["dem_5x5__std"], "path level")


# ------------------------------------------------------------------ host logic (CPU)
def test_host_kernel_helpers(hc):
    from hydrotools.cuda import interpolate as I
    assert I._as_kernel((2, 3)).shape == (2, 3) and I._as_kernel((2, 3)).all()
    with pytest.raises(IndexError):
        I._as_kernel((1, 2, 3))
    k = np.array([[1, 0, 1, 1, 0, 1], [0, 0, 0, 0, 0, 0], [1, 1, 1, 1, 1, 1]], bool)
    assert I._max_runs(k) == 3
    assert I._max_runs(np.ones((4, 4), bool)) == 1
    assert set(I.METHODS) == set(orf.METHODS)
    from hydrotools.cuda import _io
    # /root/reference/src/hydrotools/utils.py:253-265
    assert _io.infer_nodata("float32") == float(np.finfo("float32").min)
    assert _io.infer_nodata("uint16") == 65535 and _io.infer_nodata("int8") == -128
    assert _io.infer_nodata("bool") is False


@pytest.mark.gpu
def test_gpu_infinite_values_in_disc_sums(hc):
    """+-inf are data for np.nansum / np.nanmean (interpolate.py:155-160, 205-208): a window
    that holds +inf gives +inf, both signs give NaN, every other window stays finite -- also
    in the run-difference kernel, whose prefix sums must not turn into inf - inf."""
    rng = np.random.default_rng(3)
    a = rng.normal(size=(150, 300)).astype(np.float32)
    a[40, 100] = np.inf
    a[40, 180] = -np.inf
    a[90, 50] = np.inf
    a[91, 52] = -np.inf
    a[rng.random(a.shape) < 0.01] = np.nan
    yy, xx = np.mgrid[-8:9, -8:9]
    k = (xx * xx + yy * yy) <= 64
    for meth in ("sum", "mean"):
        g = hc.window_filter(a, k, meth)
        with np.errstate(invalid="ignore"):
            e = orf.raster_filter(a, k, meth)
        assert np.array_equal(np.isnan(g), np.isnan(e))
        assert np.array_equal(np.isposinf(g), np.isposinf(e)) and np.array_equal(np.isneginf(g), np.isneginf(e))
        m = np.isfinite(e)
        assert np.max(np.abs(g[m] - e[m]) / np.maximum(np.abs(e[m]), 1e-3)) < 1e-6
