"""GPU parity: D8 direction codes and cell-count accumulation are BIT-EXACT
against the r.watershed oracle on hydrologically conditioned DEMs; weighted
accumulation within 1e-5 relative (BASELINE.json north_star)."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

ND = -32767.0


def check_fd_fa(hc, dem, nodata, csx, csy, positive_only=True):
    fd, fa = hc.flow_direction_accumulation_arrays(dem, nodata=nodata, csx=csx, csy=csy,
                                                   positive_only=positive_only)
    efd, efa = oracle.rwatershed(dem, nodata, csx, csy, positive_only=positive_only)
    assert fd.dtype == np.int8 and fa.dtype == np.float64
    np.testing.assert_array_equal(fd, efd)
    np.testing.assert_array_equal(fa, efa)  # NaN == NaN positions included
    return fd, fa


def test_conditioned_bundled_dem(hc, golden):
    """configs[0] input, conditioned as north_star stipulates (rank of the oracle's
    own A* order); committed golden vectors."""
    nd, res = float(golden["dem_nodata"]), float(golden["res"])
    cdem = oracle.condition_by_rank(golden["dem"], nd, res, res)
    fd, fa = check_fd_fa(hc, cdem, nd, res, res)
    np.testing.assert_array_equal(fd, golden["cfd"])
    np.testing.assert_array_equal(np.nan_to_num(fa, nan=-1).astype(np.float32),
                                  np.nan_to_num(golden["cfa"], nan=-1))
    # sign ("likely underestimate") parity
    _, signed = check_fd_fa(hc, cdem, nd, res, res, positive_only=False)
    neg = np.unpackbits(golden["cneg"])[:signed.size].reshape(signed.shape).astype(bool)
    np.testing.assert_array_equal(signed < 0, neg)


def test_bundled_dem_raw(hc, golden):
    """BASELINE.json configs[0]: the reference's own fixture (tests/data-files/dem.tif,
    /root/reference/tests/test_watershed.py:10-16), RAW -- 2 326 pits, 30 950 tie cells, a
    quarter of the cells inside depressions -- bit-exact against the oracle and the
    committed golden vectors."""
    nd, res = float(golden["dem_nodata"]), float(golden["res"])
    pits, ties = hc.validate_conditioned(golden["dem"], nd)
    assert (pits, ties) == oracle.check_conditioned(golden["dem"], nd)
    assert pits == 2326
    with pytest.raises(ValueError, match="not hydrologically conditioned"):
        hc.flow_direction_accumulation_arrays(golden["dem"], nodata=nd, csx=res, csy=res, method="local")
    fd, fa = check_fd_fa(hc, golden["dem"], nd, res, res)
    np.testing.assert_array_equal(fd, golden["fd"])
    np.testing.assert_array_equal(np.nan_to_num(fa, nan=-1).astype(np.float32),
                                  np.nan_to_num(golden["fa"], nan=-1))
    check_fd_fa(hc, golden["dem"], nd, res, res, positive_only=False)
    from hydrotools.cuda import watershed as W
    assert W.LAST_STATS["episodes"] > 100 and W.LAST_STATS["errors"] == 0


def _fbm_unit(n, m, seed, hurst=0.8):
    rng = np.random.default_rng(seed)
    fy, fx = np.fft.fftfreq(n)[:, None], np.fft.fftfreq(m)[None, :]
    f = np.sqrt(fx * fx + fy * fy)
    f[0, 0] = 1
    z = np.fft.ifft2((rng.normal(size=(n, m)) + 1j * rng.normal(size=(n, m))) / f ** (hurst + 1)).real
    return (z - z.min()) / (z.max() - z.min())


@pytest.mark.parametrize("case", range(8))
def test_general_path_random(hc, case):
    """The general kernel on DEMs full of ties, flats, pits and NULL blobs (integer-valued
    noise; spectral fBm surfaces rounded to cm / dm), square and anisotropic cells."""
    rng = np.random.default_rng(100 + case)
    if case < 4:
        r, c = rng.integers(3, 150, 2)
        dem = rng.integers(0, [3, 6, 30, 1000][case], (r, c)).astype(np.float32)
    else:
        n, m, scale, dec = [(300, 400, 500.0, 2), (500, 500, 50.0, 1), (350, 450, 2000.0, 3),
                            (250, 250, 5.0, 0)][case - 4]
        dem = (_fbm_unit(n, m, case) * scale).round(dec).astype(np.float32)
    if case % 2:
        for _ in range(10):
            r0, c0 = rng.integers(0, dem.shape[0]), rng.integers(0, dem.shape[1])
            dem[r0:r0 + rng.integers(1, 12), c0:c0 + rng.integers(1, 12)] = ND
    csx, csy = [(10.0, 10.0), (3.0, 4.0), (4.0, 3.0)][case % 3]
    check_fd_fa(hc, dem, ND, csx, csy)
    check_fd_fa(hc, dem, ND, csx, csy, positive_only=False)


def test_general_equals_local_on_conditioned(hc):
    """On a conditioned DEM both kernels are r.watershed: identical bytes."""
    dem = oracle.synth_dem(77, 700, 900, with_nulls=True)
    a = hc.flow_direction_accumulation_arrays(dem, nodata=ND, csx=10.0, csy=10.0, method="local")
    b = hc.flow_direction_accumulation_arrays(dem, nodata=ND, csx=10.0, csy=10.0, method="general")
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])


def test_general_flat_and_lake(hc):
    """A perfectly flat DEM (breadth-first from the border) and a single deep lake."""
    flat = np.full((90, 130), 12.5, np.float32)
    check_fd_fa(hc, flat, None, 10.0, 10.0)
    lake = np.full((120, 100), 50.0, np.float32)
    yy, xx = np.mgrid[0:120, 0:100]
    lake -= np.maximum(0, 30 - np.hypot(yy - 60, xx - 50)).astype(np.float32)
    lake += (yy * 0.01).astype(np.float32)
    check_fd_fa(hc, lake, None, 10.0, 10.0)


@pytest.mark.parametrize("shape", [(1, 1), (1, 9), (9, 1), (2, 2), (3, 3), (5, 7), (63, 65),
                                   (64, 64), (65, 129), (130, 70), (200, 300), (257, 513)])
def test_ragged_shapes(hc, shape):
    dem = oracle.synth_dem(shape[0] * 7 + shape[1], shape[0], shape[1])
    assert hc.validate_conditioned(dem) == (0, 0)
    check_fd_fa(hc, dem, None, 10.0, 10.0)
    check_fd_fa(hc, dem, None, 10.0, 10.0, positive_only=False)


@pytest.mark.parametrize("seed,shape", [(1, (90, 120)), (2, (300, 500)), (3, (700, 333))])
def test_synthetic_with_nulls(hc, seed, shape):
    dem = oracle.synth_dem(seed, shape[0], shape[1], with_nulls=True)
    check_fd_fa(hc, dem, ND, 10.0, 10.0)
    check_fd_fa(hc, dem, ND, 10.0, 10.0, positive_only=False)


def test_anisotropic_cells(hc):
    """ns_res != ew_res changes the diagonal-skip rule (rule R4)."""
    dem = oracle.synth_dem(9, 220, 310)
    check_fd_fa(hc, dem, None, 30.0, 10.0)
    check_fd_fa(hc, dem, None, 5.0, 12.5)


def test_all_null_and_nan_nodata(hc):
    dem = np.full((40, 50), ND, np.float32)
    fd, fa = hc.flow_direction_accumulation_arrays(dem, nodata=ND)
    assert (fd == -128).all() and np.isnan(fa).all()
    d = oracle.synth_dem(4, 100, 100, with_nulls=True)
    d[d == ND] = np.nan
    check_fd_fa(hc, d, None, 10.0, 10.0)


def test_medium_synthetic(hc):
    """6 M cells: many tiles, long rivers through the boundary graph."""
    dem = oracle.synth_dem(20261019, 2000, 3000)
    fd, fa = check_fd_fa(hc, dem, None, 10.0, 10.0)
    assert fa.max() > 1e5


def test_size_independent_properties(hc):
    """At sizes the oracle is too slow for: mass balance of the accumulation."""
    import torch
    from hydrotools.cuda import synth
    n = 4096
    buf, ld = synth.synth_dem(77, n, n)
    fd, fa = hc.flow_direction_accumulation_arrays(buf[:, :n], csx=10.0, csy=10.0)
    fd, fa = fd.cpu().numpy(), fa.cpu().numpy()
    assert (fd != -128).all() and (fd != 0).all()
    dr = np.array([0, -1, -1, -1, 0, 1, 1, 1, 0]); dc = np.array([0, 1, 0, -1, -1, -1, 0, 1, 1])
    r, c = np.indices(fd.shape)
    edge = (r == 0) | (c == 0) | (r == n - 1) | (c == n - 1)
    donates = (fd > 0) & ~edge
    rr, cc = r + dr[np.abs(fd)], c + dc[np.abs(fd)]
    # every cell's accumulation = 1 + sum of its donors' accumulation
    inflow = np.zeros_like(fa)
    np.add.at(inflow, (rr[donates], cc[donates]), fa[donates])
    np.testing.assert_array_equal(fa, inflow + 1)
    # all flow ends in edge cells: their accumulations add up to the cell count
    assert fa[~donates].sum() == n * n


# ---------------------------------------------------------------- FlowAccumulation
def test_flowacc_counts_and_weights(hc, golden):
    nd, pnd = float(golden["dem_nodata"]), float(golden["precip_nodata"])
    for fd in (golden["fd"], golden["cfd"]):
        cnt = hc.flow_accumulation_arrays(fd)
        np.testing.assert_array_equal(cnt, oracle.flowacc(fd))
        w = hc.flow_accumulation_arrays(fd, golden["precip"], pnd)
        e = oracle.flowacc(fd, golden["precip"], pnd)
        np.testing.assert_array_equal(np.isnan(w), np.isnan(e))
        m = ~np.isnan(e)
        np.testing.assert_allclose(w[m], e[m], rtol=1e-5, atol=0)


def test_flowacc_synthetic_weighted(hc):
    dem = oracle.synth_dem(5, 900, 1100, with_nulls=True)
    fd, _ = oracle.rwatershed(dem, ND, 10.0, 10.0)
    w = oracle.synth_weight(5, 900, 1100)
    got = hc.flow_accumulation_arrays(fd, w)
    exp = oracle.flowacc(fd, w)
    m = ~np.isnan(exp)
    np.testing.assert_allclose(got[m], exp[m], rtol=1e-5, atol=0)
    # exact fixed point: integer adds commute, so the result is bit-reproducible
    for _ in range(3):
        again = hc.flow_accumulation_arrays(fd, w)
        np.testing.assert_array_equal(again, got)
    # ... and far tighter than the contract (quantisation at 2^-49 of a tile's largest weight)
    np.testing.assert_allclose(got[m], exp[m], rtol=1e-12, atol=0)


@pytest.mark.parametrize("case", ["signed", "tiny_and_huge", "zeros", "nodata_nan_inf"])
def test_flowacc_weight_ranges(hc, case):
    """Fixed-point edge cases of the weighted mode: negative weights, 12 orders of
    magnitude inside one raster, all-zero weights, NaN / inf / nodata weights."""
    rng = np.random.default_rng(11)
    dem = oracle.synth_dem(9, 400, 500, with_nulls=True)
    fd, _ = oracle.rwatershed(dem, ND, 10.0, 10.0)
    w = rng.uniform(0.5, 2.0, fd.shape).astype(np.float32)
    wnd = None
    if case == "signed":
        w *= rng.choice([-1.0, 1.0], fd.shape).astype(np.float32)
    elif case == "tiny_and_huge":
        w[:, :250] *= np.float32(1e-6)
        w[:, 250:] *= np.float32(1e6)
    elif case == "zeros":
        w[:] = 0
    else:
        wnd = -9999.0
        w[rng.random(fd.shape) < 0.1] = wnd
        w[rng.random(fd.shape) < 0.05] = np.nan
        w[7, 9] = np.inf  # treated as no data (the oracle sums it: compare without it)
    got = hc.flow_accumulation_arrays(fd, w, wnd)
    we = w.copy()
    we[~np.isfinite(we)] = np.nan
    exp = oracle.flowacc(fd, we, wnd)
    np.testing.assert_array_equal(np.isnan(got), np.isnan(exp))
    m = ~np.isnan(exp)
    if case == "signed":  # cancellation: compare against the magnitude that flowed in
        mag = oracle.flowacc(fd, np.abs(we), wnd)
        assert np.max(np.abs(got[m] - exp[m]) / np.maximum(mag[m], 1e-30)) < 1e-12
    elif case == "tiny_and_huge":
        # the documented bound of the fixed-point sums: every weight is quantised to 2^-49 of
        # the largest weight of its 64 x 64 tile, so a sum is off by at most that per cell
        # upstream; relative 1e-5 holds wherever the weights of a tile span < 9 decades
        cnt = oracle.flowacc(fd)
        assert np.all(np.abs(got[m] - exp[m]) <= 2.0 ** -48 * np.nanmax(np.abs(we)) * cnt[m])
        far = np.zeros(fd.shape, bool)
        far[:, :128] = True  # tiles that hold tiny weights only ...
        only_tiny = m & far & (exp < 1e-3)  # ... and cells that received nothing from the huge half
        np.testing.assert_allclose(got[only_tiny], exp[only_tiny], rtol=1e-9, atol=0)  # tile unit, not the global one
        big = m & (exp > 1.0)
        np.testing.assert_allclose(got[big], exp[big], rtol=1e-9, atol=0)
    else:
        np.testing.assert_allclose(got[m], exp[m], rtol=1e-9, atol=0)


@pytest.mark.timeout(900)
def test_config1_full_parity(hc):
    """BASELINE.json configs[1] at its real size: the bench's own DEM (seed 20261019,
    10 000 x 10 000) through the public array API, bit-exact against the oracle
    (/root/reference/src/hydrotools/watershed.py:64-75).  The oracle needs ~4 GB and
    about a minute on one core."""
    n = 10000
    dem = oracle.synth_dem(20261019, n, n)
    fd, fa = hc.flow_direction_accumulation_arrays(dem, nodata=None, csx=10.0, csy=10.0)
    efd, efa = oracle.rwatershed(dem, None, 10.0, 10.0)
    del dem
    np.testing.assert_array_equal(fd, efd)
    np.testing.assert_array_equal(fa, efa)


def test_misaligned_weight_raises(hc, golden):
    with pytest.raises(ValueError, match="does not align"):
        hc.flow_accumulation_arrays(golden["fd"], golden["precip"][:-1])


# ---------------------------------------------------------------- signatures / files
def test_unsupported_modes(hc):
    with pytest.raises(NotImplementedError):
        hc.flow_direction_accumulation("a.tif", "b.tif", "c.tif", single=False)
    with pytest.raises(NotImplementedError):
        hc.flow_direction_accumulation("a.tif", "b.tif", "c.tif", beautify=True)


def test_path_level_api(hc, golden, tmp_path):
    from hydrotools.cuda import _io
    from test_terrain_gpu import _write_template
    nd, res = float(golden["dem_nodata"]), float(golden["res"])
    cdem = oracle.condition_by_rank(golden["dem"], nd, res, res)
    tmpl, src = str(tmp_path / "t.tif"), str(tmp_path / "dem.tif")
    _write_template(tmpl, cdem, nd)
    _io.write_raster(cdem, tmpl, src, nodata=nd)
    fdp, fap = str(tmp_path / "fd.tif"), str(tmp_path / "fa.tif")
    hc.flow_direction_accumulation(src, fdp, fap, memory=4096)
    fd, m = _io.read_raster(fdp)
    fa, _ = _io.read_raster(fap)
    np.testing.assert_array_equal(fd, golden["cfd"])
    assert m["nodata"] == -128 and fa.dtype == np.float64
    # FlowAccumulation on the direction raster just written
    pp = str(tmp_path / "precip.tif")
    _io.write_raster(golden["precip"], tmpl, pp, nodata=float(golden["precip_nodata"]))
    facc = hc.FlowAccumulation(fdp)
    facc.calculate(pp, str(tmp_path / "sum.tif"), "sum")
    facc.calculate(pp, str(tmp_path / "mean.tif"), "mean")
    facc.contributing_area(str(tmp_path / "ca.tif"))
    s, _ = _io.read_raster(str(tmp_path / "sum.tif"))
    e = oracle.flowacc(golden["cfd"], golden["precip"], float(golden["precip_nodata"]))
    ok = ~np.isnan(e)
    np.testing.assert_allclose(s[ok], e[ok], rtol=1e-5)
    mean, mm = _io.read_raster(str(tmp_path / "mean.tif"))
    cnt = oracle.flowacc(golden["cfd"])
    np.testing.assert_allclose(mean[ok], (e[ok] / cnt[ok]).astype(np.float32), rtol=1e-5)
    ca, _ = _io.read_raster(str(tmp_path / "ca.tif"))
    assert ca.dtype == np.float32  # like the reference (float32 modals, watershed.py:392-405)
    np.testing.assert_allclose(ca[ok], cnt[ok] * 25.0 / 1e6, rtol=1e-6)


def test_cyclic_direction_raster_terminates(hc):
    """A corrupt direction raster (two cells pointing at each other, fed by a chain)
    must neither hang nor disturb the cells that are not downstream of the cycle."""
    dem = oracle.synth_dem(11, 200, 260)
    fd, _ = oracle.rwatershed(dem, None, 10.0, 10.0)
    fd = fd.copy()
    fd[100, 100], fd[100, 101] = 8, 4   # E <-> W
    fd[100, 99] = 8                     # flows into the cycle
    got = hc.flow_accumulation_arrays(fd)
    exp = oracle.flowacc(fd)            # Kahn queue: the cycle is never released
    # cells whose value does not depend on the cycle: everything the oracle resolved
    # outside the cycle's downstream (the cycle has none) and outside the cycle itself
    stuck = np.zeros(fd.shape, bool)
    stuck[100, 100] = stuck[100, 101] = True
    np.testing.assert_array_equal(got[~stuck], exp[~stuck])
    # weighted mode (pointer doubling in both stages) terminates as well
    w = oracle.synth_weight(11, 200, 260)
    gw = hc.flow_accumulation_arrays(fd, w)
    ew = oracle.flowacc(fd, w)
    np.testing.assert_allclose(gw[~stuck], ew[~stuck], rtol=1e-5)


def test_wide_band_shape_mass_balance(hc):
    """The per-GPU shape of BASELINE.json configs[4] in small: a 100 000-column band
    (1563 tile columns, 32-bit slot arithmetic near its largest column index); checked
    on the device through acc(c) = 1 + sum of acc over c's donors."""
    import torch
    from hydrotools.cuda import synth
    rows, cols = 700, 100000
    buf, ld = synth.synth_dem(5, rows, cols)
    fd, fa = hc.flow_direction_accumulation_arrays(buf[:, :cols], csx=10.0, csy=10.0)
    dr = torch.tensor([0, -1, -1, -1, 0, 1, 1, 1, 0], device=fd.device)
    dc = torch.tensor([0, 1, 0, -1, -1, -1, 0, 1, 1], device=fd.device)
    f = fd.to(torch.int64)
    rr = torch.arange(rows, device=fd.device).view(-1, 1).expand(-1, cols)
    cc = torch.arange(cols, device=fd.device).view(1, -1).expand(rows, -1)
    edge = (rr == 0) | (cc == 0) | (rr == rows - 1) | (cc == cols - 1)
    don = (f > 0) & ~edge
    code = f.clamp(min=0)
    idx = ((rr + dr[code]) * cols + cc + dc[code])[don]
    inflow = torch.zeros(rows * cols, dtype=torch.float64, device=fd.device)
    inflow.index_add_(0, idx, fa[don])
    assert torch.equal(fa, inflow.view(rows, cols) + 1)
    assert float(fa[~don].sum()) == rows * cols


def _plant_slope_ties(dem, csx, csy):
    """Raise single cells so that the slope towards the SE neighbour over the
    diagonal EQUALS the slope towards the E neighbour (3-4-5 cells: diagonal 5):
    (c - SE) / 5 == (c - E) / csx.  Keeps only plantings that leave the DEM
    conditioned.  Returns (dem, planted cell count)."""
    a = np.rint(dem.astype(np.float64) * 1000).astype(np.int64)
    out, n = dem.copy(), 0
    k_e = int(round(csx))  # run towards E in the 3-4-5 geometry
    for r in range(40, dem.shape[0] - 40, 13):
        for c in range(140, dem.shape[1] - 140, 23):
            ne, e = a[r + 1, c + 1], a[r, c + 1]  # `ne`: the diagonal neighbour (SE here)
            # want c0 with (c0 - se) * k_e == (c0 - e) * 5  ->  c0 = (5 e - k_e se) / (5 - k_e)
            num, den = 5 * e - k_e * ne, 5 - k_e
            if e <= ne or num % den:
                continue
            c0 = num // den
            win = a[r - 1:r + 2, c - 1:c + 2].copy()
            win[1, 1] = c0
            if c0 <= e or len(set(win.ravel().tolist())) != 9:
                continue
            trial = out.copy()
            trial[r, c] = np.float32(c0 / 1000.0)
            if int(round(float(trial[r, c]) * 1000)) != c0:
                continue
            if oracle.check_conditioned(trial[r - 3:r + 4, c - 3:c + 4].copy(), None)[1] != 0:
                continue
            out, n = trial, n + 1
            a[r, c] = c0
    return out, n


@pytest.mark.parametrize("csx,csy", [(3.0, 4.0), (4.0, 3.0), (30.0, 10.0), (10.0, 10.0)])
def test_plain_tiles_anisotropic(hc, csx, csy):
    """Rasters large enough to have D8 `plain` tiles (the fp32 fast path), square and
    anisotropic cells."""
    dem = oracle.synth_dem(31, 420, 900)
    check_fd_fa(hc, dem, None, csx, csy)


@pytest.mark.parametrize("csx,csy", [(3.0, 4.0), (4.0, 3.0)])
def test_plain_tiles_exact_slope_ties(hc, csx, csy):
    """With 3-4-5 cells the diagonal / cardinal slope comparison of rule R4 can tie
    exactly; GRASS then keeps the diagonal (`<`, not `<=`).  The fp32 fast path must
    hand such cells to the exact integer / double evaluation to stay bit-identical."""
    dem, planted = _plant_slope_ties(oracle.synth_dem(31, 420, 900), csx, csy)
    assert planted >= 10
    if oracle.check_conditioned(dem, None) != (0, 0):
        pytest.skip("planting broke the conditioning")
    check_fd_fa(hc, dem, None, csx, csy)


def _fbm(seed, rows, cols, hurst=0.8):
    """Spectral-synthesis fractional Brownian surface (natural-looking terrain with pits
    and flats everywhere once rounded to millimetres), in metres."""
    rng = np.random.default_rng(seed)
    fy = np.fft.fftfreq(rows)[:, None]
    fx = np.fft.rfftfreq(cols)[None, :]
    f = np.sqrt(fy * fy + fx * fx)
    f[0, 0] = 1.0
    amp = f ** (-(hurst + 1.0))
    amp[0, 0] = 0.0
    spec = amp * np.exp(2j * np.pi * rng.random(amp.shape))
    z = np.fft.irfft2(spec, s=(rows, cols))
    z = (z - z.min()) / (z.max() - z.min())
    return (z * 800.0).astype(np.float32)


def test_natural_terrain_conditioned_by_rank(hc):
    """SURVEY.md 8(d)'s harder case: a natural fBm surface (no built-in tilt, valleys
    in every direction), hydrologically conditioned the way the bundled DEM is in
    test_conditioned_bundled_dem -- elevation := r.watershed's own A* pop rank -- so
    that rivers wander for thousands of cells through tiles and bands."""
    raw = _fbm(7, 1100, 1400)
    assert oracle.check_conditioned(raw, None)[0] > 0  # the raw surface has pits
    dem = oracle.condition_by_rank(raw, None, 10.0, 10.0)
    assert oracle.check_conditioned(dem, None) == (0, 0)
    fd, fa = check_fd_fa(hc, dem, None, 10.0, 10.0)
    check_fd_fa(hc, dem, None, 10.0, 10.0, positive_only=False)
    assert fa.max() > 1e5
    # and in row bands on one device
    from test_bands_gpu import run_bands
    bfd, bfa = run_bands(dem, [0, 300, 640, 1100], None, 10.0, 10.0, delta=True)
    np.testing.assert_array_equal(bfd, fd)
    np.testing.assert_array_equal(bfa, fa)


def test_half_millimetre_rounding(hc):
    """Rule R1: every elevation sits on (or one float next to) a half millimetre, where
    GRASS's (int)(z * 1000 +- 0.5) in double decides between two integers.  The doubled grid
    keeps the DEM conditioned whichever way a cell rounds, so every direction code and count
    below depends on the exact integers."""
    base = oracle.synth_dem(21, 300, 420)
    zi = np.rint(base.astype(np.float64) * 1000.0).astype(np.int64)
    rng = np.random.default_rng(5)
    for sign in (1.0, -1.0):
        z = ((2 * zi + 0.5) / 1000.0).astype(np.float32)
        nudge = rng.integers(-1, 2, z.shape)
        z = np.where(nudge < 0, np.nextafter(z, np.float32(-1e9)),
                     np.where(nudge > 0, np.nextafter(z, np.float32(1e9)), z)).astype(np.float32)
        if sign < 0:  # negative elevations round away from zero as well
            z = (z - np.float32(2.0 * (zi.max() + 10) / 1000.0)).astype(np.float32)
            assert (z < 0).all()
        assert oracle.check_conditioned(z) == (0, 0)
        check_fd_fa(hc, z, None, 10.0, 10.0)
