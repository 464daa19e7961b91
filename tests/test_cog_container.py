"""CPU: the COG container hydrotools.cuda._cog writes around the device-compressed tiles
(/root/reference/src/hydrotools/raster.py:794-846, utils.py:78-98) -- with a
zlib stand-in for the tile encoder, so that the layout logic is checked without a GPU: the
file must read back through the product's TIFF reader and through libtiff (PIL), IFDs ahead of
all pixel data, overview data ahead of the full resolution.  The encoder kernels themselves
are checked on the GPU (tests/test_cog_gpu.py)."""
import struct
import zlib

import numpy as np
import pytest
import torch

TILE = 512


def zlib_stand_in(a, pad=0):
    a = a.numpy()
    rows, cols = a.shape
    blobs, off, size, pos = [], [], [], 0
    for ty in range(-(-rows // TILE)):
        for tx in range(-(-cols // TILE)):
            t = np.full((TILE, TILE), pad, a.dtype)
            sub = a[ty * TILE:(ty + 1) * TILE, tx * TILE:(tx + 1) * TILE]
            t[:sub.shape[0], :sub.shape[1]] = sub
            b = zlib.compress(t.tobytes(), 1)
            off.append(pos); size.append(len(b)); pos += len(b); blobs.append(b)
    return np.frombuffer(b"".join(blobs), np.uint8), np.array(off), np.array(size)


@pytest.mark.parametrize("dtype,nodata", [(np.int8, -128), (np.float64, float("nan")), (np.float32, -32767.0)])
def test_container_reads_back(hc, golden, tmp_path, monkeypatch, dtype, nodata):
    from hydrotools.cuda import _cog, _io
    from test_terrain_gpu import _write_template
    monkeypatch.setattr(_cog, "encode_tiles", zlib_stand_in)
    src = {np.int8: golden["fd"], np.float64: golden["fa"].astype(np.float64),
           np.float32: golden["dem"].astype(np.float32)}[dtype]
    a = np.tile(src, (3, 2))[:1100, :1300]  # 3 x 3 tiles, three levels
    tmpl, dst = str(tmp_path / "t.tif"), str(tmp_path / "out.tif")
    _write_template(tmpl, a.astype(np.float32), -32767.0)
    info = _cog.write_cog(torch.from_numpy(a), _io._read_ifd(tmpl), dst, nodata=nodata)
    assert info["levels"] == 3
    got, meta = _io.read_raster(dst)
    assert got.dtype == a.dtype and meta["csx"] == 5.0
    np.testing.assert_array_equal(np.nan_to_num(got, nan=-7), np.nan_to_num(a, nan=-7))
    if dtype == np.float32:  # libtiff's view of the same file
        from PIL import Image
        np.testing.assert_array_equal(np.array(Image.open(dst)), a)
    # IFD chain and data order
    with open(dst, "rb") as f:
        head = f.read(16)
        assert head[:4] == b"II\x2b\x00"
        off = struct.unpack("<Q", head[8:])[0]
        ifd_offs, firsts, scales = [], [], []
        while off:
            assert off % 8 == 0
            ifd_offs.append(off)
            f.seek(off)
            n = struct.unpack("<Q", f.read(8))[0]
            raw = f.read(20 * n + 8)
            for i in range(n):
                tag, typ, cnt, val = struct.unpack("<HHQ8s", raw[20 * i:20 * i + 20])
                if tag in (324, 33550):
                    here = f.tell()
                    if cnt * (8 if tag == 324 else 8) <= 8:
                        vals = struct.unpack("<Q" if tag == 324 else "<d", val)
                    else:
                        f.seek(struct.unpack("<Q", val)[0])
                        vals = struct.unpack("<" + ("Q" if tag == 324 else "d") * cnt, f.read(8 * cnt))
                    f.seek(here)
                    (firsts if tag == 324 else scales).append(vals)
            off = struct.unpack("<Q", raw[20 * n:])[0]
    assert len(ifd_offs) == 3 and max(ifd_offs) < min(min(v) for v in firsts)
    assert min(firsts[2]) < min(firsts[1]) < min(firsts[0])      # smallest overview first
    assert [s[0] for s in scales] == [5.0, 10.0, 20.0]            # pixel size doubles per level
