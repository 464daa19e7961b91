"""GPU: device-side GeoTIFF tile encoder and COG writer (reference:
/root/reference/src/hydrotools/raster.py:794-846, utils.py:78-98, config.py:23-45).  Every tile the
kernels emit must be a valid zlib stream (CPython's zlib is the judge: it checks the Huffman
codes, the stored sync blocks and the Adler-32) that inflates to the tile's bytes; the
container must read back through the TIFF reader of hydrotools.cuda._io."""
import struct
import zlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TILE = 512


def tiles_of(a, pad):
    rows, cols = a.shape
    for ty in range(-(-rows // TILE)):
        for tx in range(-(-cols // TILE)):
            t = np.full((TILE, TILE), pad, a.dtype)
            sub = a[ty * TILE:(ty + 1) * TILE, tx * TILE:(tx + 1) * TILE]
            t[:sub.shape[0], :sub.shape[1]] = sub
            yield t


@pytest.mark.parametrize("dtype,shape,pad", [
    (np.float64, (700, 900), np.nan), (np.float32, (513, 511), -9999.0), (np.int8, (1200, 600), -128),
    (np.uint8, (512, 512), 0), (np.int16, (100, 2000), -1), (np.int32, (1, 1), 7), (np.float64, (1030, 1), 0.0),
])
def test_tiles_inflate_to_the_raster(hc, dtype, shape, pad):
    from hydrotools.cuda import _cog
    rng = np.random.default_rng(shape[0] + shape[1])
    if np.dtype(dtype).kind == "f":
        a = np.round(rng.random(shape) * 50).astype(dtype)       # accumulation-like: small integers
        a[rng.random(shape) < 0.05] = rng.random() * 1e6
        a[rng.random(shape) < 0.02] = np.nan
    else:
        a = rng.integers(-8 if np.dtype(dtype).kind == "i" else 0, 9, shape).astype(dtype)
    data, off, size = _cog.encode_tiles(torch.from_numpy(a).cuda(), pad)
    assert len(off) == -(-shape[0] // TILE) * -(-shape[1] // TILE)
    assert off[0] == 0 and np.all(np.diff(off) == size[:-1]) and off[-1] + size[-1] == len(data)
    for k, t in enumerate(tiles_of(a, pad)):
        raw = zlib.decompress(data[off[k]:off[k] + size[k]].tobytes())
        assert raw == t.tobytes(), f"tile {k}"
    assert len(data) < len(off) * TILE * TILE * a.itemsize * 1.16  # fixed Huffman: 9 bits per byte at worst


def test_incompressible_and_constant_tiles(hc):
    from hydrotools.cuda import _cog
    rng = np.random.default_rng(3)
    noise = rng.integers(0, 256, (512, 1024), dtype=np.uint8)
    data, off, size = _cog.encode_tiles(torch.from_numpy(noise).cuda(), 0)
    for k, t in enumerate(tiles_of(noise, 0)):
        assert zlib.decompress(data[off[k]:off[k] + size[k]].tobytes()) == t.tobytes()
    const = np.full((600, 600), 3.25, np.float32)
    data, off, size = _cog.encode_tiles(torch.from_numpy(const).cuda(), 3.25)
    assert size.max() < 512 * 512 * 4 / 50  # long matches: at least 50 : 1
    for k, t in enumerate(tiles_of(const, 3.25)):
        assert zlib.decompress(data[off[k]:off[k] + size[k]].tobytes()) == t.tobytes()


def _ifds(path):
    """[(tags dict)] of every IFD of a little-endian BigTIFF."""
    out = []
    with open(path, "rb") as f:
        head = f.read(16)
        assert head[:4] == b"II\x2b\x00"
        off = struct.unpack("<Q", head[8:])[0]
        while off:
            f.seek(off)
            n = struct.unpack("<Q", f.read(8))[0]
            raw = f.read(20 * n + 8)
            tags = {}
            for i in range(n):
                tag, typ, cnt, val = struct.unpack("<HHQ8s", raw[20 * i:20 * i + 20])
                tags[tag] = (typ, cnt, val)
            out.append((off, tags))
            off = struct.unpack("<Q", raw[20 * n:])[0]
    return out


def test_cog_container_roundtrip(hc, golden, tmp_path):
    from hydrotools.cuda import _cog, _io
    from test_terrain_gpu import _write_template
    dem = golden["dem"].astype(np.float32)
    nd = float(golden["dem_nodata"])
    tmpl, dst = str(tmp_path / "t.tif"), str(tmp_path / "cog.tif")
    big = np.tile(dem, (4, 2))[:1500, :1300]  # several tiles, three overview levels
    _write_template(tmpl, big, nd)
    info = _cog.write_cog(torch.from_numpy(big).cuda(), _io._read_ifd(tmpl), dst, nodata=nd)
    got, meta = _io.read_raster(dst)
    np.testing.assert_array_equal(got, big)
    assert meta["nodata"] == nd and meta["csx"] == 5.0 and meta["shape"] == big.shape
    assert info["compressed"] < info["raw"] and info["levels"] == 3
    ifds = _ifds(dst)
    assert len(ifds) == 3
    # COG layout: every IFD before any tile data; overview data ahead of the full resolution
    first_data = []
    for off, tags in ifds:
        typ, cnt, val = tags[324]
        with open(dst, "rb") as f:
            if cnt == 1:
                offs = [struct.unpack("<Q", val)[0]]
            else:
                f.seek(struct.unpack("<Q", val)[0])
                offs = struct.unpack("<" + "Q" * cnt, f.read(8 * cnt))
        first_data.append(min(offs))
    assert max(off for off, _ in ifds) < min(first_data)
    assert first_data[2] < first_data[1] < first_data[0]
    assert 254 not in ifds[0][1] and struct.unpack("<Q", ifds[1][1][254][2])[0] == 1
    # the first overview is the nearest-neighbour half
    w = struct.unpack("<Q", ifds[1][1][256][2])[0]
    h = struct.unpack("<Q", ifds[1][1][257][2])[0]
    assert (h, w) == big[::2, ::2].shape


def test_path_level_writes_through_the_device_encoder(hc, golden, tmp_path, monkeypatch):
    from hydrotools.cuda import _io
    from test_terrain_gpu import _write_template
    monkeypatch.setenv("HYDROTOOLS_CUDA_IO", "device")
    dem = golden["dem"].astype(np.float32)
    nd, res = float(golden["dem_nodata"]), float(golden["res"])
    tmpl, src = str(tmp_path / "t.tif"), str(tmp_path / "dem.tif")
    fdp, fap = str(tmp_path / "fd.tif"), str(tmp_path / "fa.tif")
    _write_template(tmpl, dem, nd)
    _io.write_raster(dem, tmpl, src, nodata=nd)
    hc.flow_direction_accumulation(src, fdp, fap)
    fd, m = _io.read_raster(fdp)
    fa, _ = _io.read_raster(fap)
    np.testing.assert_array_equal(fd, golden["fd"])
    np.testing.assert_array_equal(np.nan_to_num(fa, nan=-1).astype(np.float32),
                                  np.nan_to_num(golden["fa"], nan=-1))
    assert m["nodata"] == -128 and fd.dtype == np.int8 and fa.dtype == np.float64
