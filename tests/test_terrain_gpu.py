"""GPU parity: fused slope/aspect/TRI kernel vs the gdaldem oracle.

Contract (BASELINE.json north_star): <= 1e-5 relative fp32; aspect is compared
on the circle (0 == 360) and the nodata mask must be identical except where the
oracle's own value sits within float noise of the flat test.
"""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

RTOL = 1e-5
ND = -9999.0


def compare(got, exp, dem_scale=1.0):
    for k in ("slope", "tri"):
        g, e = got[k], exp[k]
        assert g.shape == e.shape
        np.testing.assert_array_equal(g == ND, e == ND, err_msg=k)
        m = e != ND
        # absolute floor: fp32 cancellation noise of the window sums
        np.testing.assert_allclose(g[m], e[m], rtol=RTOL, atol=2e-6 * dem_scale, err_msg=k)
    g, e = got["aspect"], exp["aspect"]
    both = (g != ND) & (e != ND)
    d = np.abs(g[both] - e[both])
    d = np.minimum(d, 360.0 - d)
    assert d.max(initial=0) <= 360.0 * RTOL, d.max()
    # flat cells: dx == dy == 0 exactly in float on both sides
    np.testing.assert_array_equal(g == ND, e == ND)


def run_case(hc, dem, nodata, csx, csy, scale=1.0, units="degrees"):
    got = hc.terrain(dem, nodata=nodata, csx=csx, csy=csy, scale=scale, units=units)
    s, a, t = oracle.gdaldem(dem, nodata, csx, csy, scale, units == "percent")
    compare(got, dict(slope=s, aspect=a, tri=t), float(np.abs(dem[dem != (nodata or 0)]).max()))


def test_bundled_dem(hc, golden):
    run_case(hc, golden["dem"], float(golden["dem_nodata"]), 5.0, 5.0)
    run_case(hc, golden["dem"], float(golden["dem_nodata"]), 5.0, 5.0, 2.0, "percent")


@pytest.mark.parametrize("shape", [(2, 2), (3, 3), (2, 9), (9, 2), (33, 129), (32, 128),
                                   (31, 127), (64, 256), (65, 257), (100, 136), (257, 1031)])
def test_ragged_shapes(hc, shape):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    dem = (rng.random(shape, dtype=np.float32) * 50 + 100).astype(np.float32)
    run_case(hc, dem, None, 10.0, 10.0)
    dem[rng.random(shape) < 0.15] = -32767.0
    run_case(hc, dem, -32767.0, 10.0, 7.5)


@pytest.mark.parametrize("shape", [(1, 1), (1, 7), (7, 1)])
def test_degenerate_all_nodata(hc, shape):
    got = hc.terrain(np.ones(shape, np.float32))
    for v in got.values():
        assert v.shape == shape and (v == ND).all()


def test_nan_nodata_and_nan_values(hc):
    rng = np.random.default_rng(1)
    dem = rng.random((70, 150), dtype=np.float32) * 10
    dem[rng.random(dem.shape) < 0.1] = np.nan
    got = hc.terrain(dem, nodata=float("nan"), csx=2.0, csy=2.0)
    s, a, t = oracle.gdaldem(dem, float("nan"), 2.0, 2.0)
    compare(got, dict(slope=s, aspect=a, tri=t), 10.0)


def test_single_outputs_match_fused(hc, golden):
    dem, nd = golden["dem"], float(golden["dem_nodata"])
    fused = hc.terrain(dem, nodata=nd, csx=5.0, csy=5.0)
    for k in ("slope", "aspect", "tri"):
        one = hc.terrain(dem, nodata=nd, csx=5.0, csy=5.0, outputs=(k,))
        np.testing.assert_array_equal(one[k], fused[k])


def test_synthetic_with_nulls(hc):
    dem = oracle.synth_dem(11, 300, 500, with_nulls=True)
    run_case(hc, dem, -32767.0, 10.0, 10.0)


def test_many_tiles_per_cta(hc):
    """More tiles than 3 stages x 296 persistent CTAs: exercises stage reuse and
    mbarrier phase flips of the TMA pipeline."""
    dem = oracle.synth_dem(3, 2100, 4200, with_nulls=True)
    run_case(hc, dem, -32767.0, 10.0, 10.0)


def test_row_bands_equal_whole(hc):
    """Row-band sharding: bands with one halo row
    reproduce the unsharded result bit for bit."""
    dem = oracle.synth_dem(5, 200, 260, with_nulls=True)
    whole = hc.terrain(dem, nodata=-32767.0, csx=10.0, csy=10.0)
    cuts = [0, 37, 38, 120, 200]
    for k in ("slope", "aspect", "tri"):
        parts = []
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            top, bot = int(lo > 0), int(hi < 200)
            band = dem[lo - top:hi + bot]
            parts.append(hc.terrain(band, nodata=-32767.0, csx=10.0, csy=10.0,
                                    halo_top=top, halo_bot=bot)[k])
        np.testing.assert_array_equal(np.vstack(parts), whole[k])


def test_device_tensor_in_device_tensor_out(hc):
    import torch
    dem = oracle.synth_dem(2, 100, 140)
    out = hc.terrain(torch.from_numpy(dem).cuda(), csx=10.0, csy=10.0)
    assert all(v.is_cuda for v in out.values())
    host = hc.terrain(dem, csx=10.0, csy=10.0)
    for k in out:
        np.testing.assert_array_equal(out[k].cpu().numpy(), host[k])


def test_path_level_api_roundtrip(hc, golden, tmp_path):
    from hydrotools.cuda import _io
    src = str(tmp_path / "dem.tif")
    tmpl = str(tmp_path / "tmpl.tif")
    # a template GeoTIFF written by the fallback writer (5 m pixels like the fixture)
    _write_template(tmpl, golden["dem"], float(golden["dem_nodata"]))
    _io.write_raster(golden["dem"], tmpl, src, nodata=float(golden["dem_nodata"]))
    a, meta = _io.read_raster(src)
    np.testing.assert_array_equal(a, golden["dem"])
    assert meta["nodata"] == float(golden["dem_nodata"]) and meta["csx"] == 5.0
    dst = str(tmp_path / "slope.tif")
    hc.slope(src, dst)
    s, m2 = _io.read_raster(dst)
    exp, _, _ = oracle.gdaldem(golden["dem"], float(golden["dem_nodata"]), 5.0, 5.0)
    compare(dict(slope=s, tri=s, aspect=s), dict(slope=exp, tri=exp, aspect=exp), 100.0)
    assert m2["nodata"] == ND and m2["csx"] == 5.0
    hc.aspect(src, str(tmp_path / "aspect.tif"))
    hc.terrain_ruggedness_index(src, str(tmp_path / "tri.tif"))


def _write_template(path, dem, nodata):
    """Minimal georeferenced template, built by hand (no GDAL here)."""
    import struct
    h, w = dem.shape
    data = np.zeros((h, w), np.float32).tobytes()
    entries = [
        (256, 4, 1, struct.pack("<I", w)), (257, 4, 1, struct.pack("<I", h)),
        (258, 3, 1, struct.pack("<H", 32)), (259, 3, 1, struct.pack("<H", 1)),
        (262, 3, 1, struct.pack("<H", 1)), (273, 4, 1, struct.pack("<I", 8)),
        (277, 3, 1, struct.pack("<H", 1)), (278, 4, 1, struct.pack("<I", h)),
        (279, 4, 1, struct.pack("<I", len(data))), (339, 3, 1, struct.pack("<H", 3)),
        (33550, 12, 3, struct.pack("<3d", 5.0, 5.0, 0.0)),
        (33922, 12, 6, struct.pack("<6d", 0, 0, 0, -13622276.0, 6292862.0, 0)),
        (34735, 3, 4, struct.pack("<4H", 1, 1, 0, 0)),
    ]
    ifd_off = 8 + len(data)
    extra_off = ifd_off + 2 + 12 * len(entries) + 4
    ifd, extra = struct.pack("<H", len(entries)), b""
    for tag, typ, cnt, b in entries:
        if len(b) <= 4:
            ifd += struct.pack("<HHI", tag, typ, cnt) + b.ljust(4, b"\x00")
        else:
            ifd += struct.pack("<HHII", tag, typ, cnt, extra_off + len(extra))
            extra += b
    ifd += struct.pack("<I", 0)
    with open(path, "wb") as f:
        f.write(b"II" + struct.pack("<HI", 42, ifd_off) + data + ifd + extra)
