"""Raster I/O boundary of the drop-in.

The reference reads rasters with ``hydrotools.raster.from_raster`` (masked dask
array, chunks (1, 4096, 4096), mask = ``== nodata | isnan``;
/root/reference/src/hydrotools/raster.py:773-791) and writes them with
``hydrotools.raster.to_raster`` (tiled 512x512 LZW GeoTIFF -> COG;
raster.py:794-846).  When that package (rasterio + dask + GDAL) is importable
this module calls exactly those two functions, so file formats stay what the
reference produces.  Where it is not (this image has neither rasterio nor GDAL)
a minimal GeoTIFF reader/writer keeps the path-level API usable: same grid,
georeferencing tags copied from the template, 512x512 DEFLATE tiles.
"""
import struct
import zlib
from typing import Union,  Dict, Optional, Tuple

import numpy as np

try:  # the reference's own I/O layer, unchanged
    from hydrotools.raster import Raster as _RefRaster  # type: ignore
    from hydrotools.raster import from_raster as _ref_from_raster  # type: ignore
    from hydrotools.raster import to_raster as _ref_to_raster  # type: ignore
    HAVE_REFERENCE_IO = True
except Exception:  # pragma: no cover - depends on the environment
    HAVE_REFERENCE_IO = False

_GEO_TAGS = (33550, 33922, 34264, 34735, 34736, 34737)
_TYPE_FMT = {1: "B", 2: "c", 3: "H", 4: "I", 5: "II", 6: "b", 8: "h", 9: "i",
             11: "f", 12: "d", 16: "Q", 17: "q"}
_SAMPLE_DTYPE = {(1, 8): "u1", (1, 16): "u2", (1, 32): "u4", (1, 64): "u8",
                 (2, 8): "i1", (2, 16): "i2", (2, 32): "i4", (2, 64): "i8",
                 (3, 32): "f4", (3, 64): "f8"}


# --------------------------------------------------------------------------
# minimal TIFF
# --------------------------------------------------------------------------
def _read_ifd(path: str) -> Dict[int, tuple]:
    with open(path, "rb") as f:
        head = f.read(16)
        if head[:2] != b"II":
            raise ValueError("only little-endian TIFF is supported by the fallback reader")
        magic = struct.unpack("<H", head[2:4])[0]
        big = magic == 43
        if big:
            off = struct.unpack("<Q", head[8:16])[0]
        elif magic == 42:
            off = struct.unpack("<I", head[4:8])[0]
        else:
            raise ValueError("not a TIFF file")
        f.seek(off)
        n = struct.unpack("<Q" if big else "<H", f.read(8 if big else 2))[0]
        esz, vsz = (20, 8) if big else (12, 4)
        raw = f.read(n * esz)
        tags = {}
        for i in range(n):
            e = raw[i * esz:(i + 1) * esz]
            tag, typ = struct.unpack("<HH", e[:4])
            cnt = struct.unpack("<Q" if big else "<I", e[4:4 + vsz])[0]
            val = e[4 + vsz:]
            fmt = _TYPE_FMT.get(typ)
            if fmt is None:
                continue
            size = struct.calcsize("<" + fmt) * cnt
            if size > vsz:
                pos = struct.unpack("<Q" if big else "<I", val[:vsz])[0]
                here = f.tell()
                f.seek(pos)
                val = f.read(size)
                f.seek(here)
            else:
                val = val[:size]
            if typ == 2:
                tags[tag] = (val.rstrip(b"\x00").decode("latin1"),)
            else:
                tags[tag] = struct.unpack("<" + fmt * cnt, val)
        return tags


def _decode_pixels(path: str, tags: Dict[int, tuple]) -> Optional[np.ndarray]:
    comp = tags.get(259, (1,))[0]
    if comp not in (1, 8, 32946) or tags.get(317, (1,))[0] != 1 or tags.get(277, (1,))[0] != 1:
        return None
    w, h = tags[256][0], tags[257][0]
    dt = _SAMPLE_DTYPE.get((tags.get(339, (1,))[0], tags[258][0]))
    if dt is None:
        return None
    out = np.empty((h, w), dtype="<" + dt)
    tiled = 322 in tags
    offs = tags[324] if tiled else tags[273]
    cnts = tags[325] if tiled else tags[279]
    with open(path, "rb") as f:
        def block(i):
            f.seek(offs[i])
            b = f.read(cnts[i])
            return zlib.decompress(b) if comp != 1 else b
        if tiled:
            tw, th = tags[322][0], tags[323][0]
            nx = (w + tw - 1) // tw
            for i in range(len(offs)):
                ty, tx = divmod(i, nx)
                t = np.frombuffer(block(i), dtype="<" + dt).reshape(th, tw)
                y0, x0 = ty * th, tx * tw
                out[y0:y0 + th, x0:x0 + tw] = t[:min(th, h - y0), :min(tw, w - x0)]
        else:
            rps = tags.get(278, (h,))[0]
            for i in range(len(offs)):
                y0 = i * rps
                n = min(rps, h - y0)
                out[y0:y0 + n] = np.frombuffer(block(i), dtype="<" + dt)[:n * w].reshape(n, w)
    return out


def _fallback_read(path: str) -> Tuple[np.ndarray, dict]:
    tags = _read_ifd(path)
    a = _decode_pixels(path, tags)
    if a is None:  # LZW / predictor etc.: let PIL decode
        from PIL import Image
        Image.MAX_IMAGE_PIXELS = None
        a = np.array(Image.open(path))
    nodata = None
    if 42113 in tags:
        try:
            nodata = float(tags[42113][0].strip())
        except ValueError:
            nodata = None
    scale = tags.get(33550, (1.0, 1.0, 0.0))
    tie = tags.get(33922, (0.0, 0.0, 0.0, 0.0, 0.0, 0.0))
    meta = dict(nodata=nodata, csx=float(scale[0]), csy=float(scale[1]),
                left=float(tie[3] - tie[0] * scale[0]), top=float(tie[4] + tie[1] * scale[1]),
                shape=a.shape, dtype=a.dtype.name,
                geo_tags={t: tags[t] for t in _GEO_TAGS if t in tags})
    return a, meta


def _fallback_write(a: np.ndarray, template: str, dst: str, nodata) -> None:
    tags_t = _read_ifd(template)
    a = np.ascontiguousarray(a)
    if a.dtype.byteorder == ">":
        a = a.astype(a.dtype.newbyteorder("<"))
    h, w = a.shape
    kind = {"u": 1, "i": 2, "f": 3}[a.dtype.kind]
    bits = a.dtype.itemsize * 8
    tw = th = 512
    nx, ny = (w + tw - 1) // tw, (h + th - 1) // th
    blocks = []
    for ty in range(ny):
        for tx in range(nx):
            t = np.zeros((th, tw), a.dtype)
            sub = a[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw]
            t[:sub.shape[0], :sub.shape[1]] = sub
            blocks.append(zlib.compress(t.tobytes(), 1))
    entries = []  # (tag, type, count, bytes)

    def add(tag, typ, values):
        if typ == 2:
            b = values.encode("latin1") + b"\x00"
            entries.append((tag, 2, len(b), b))
        else:
            fmt = _TYPE_FMT[typ]
            entries.append((tag, typ, len(values), struct.pack("<" + fmt * len(values), *values)))

    add(256, 4, (w,)); add(257, 4, (h,)); add(258, 3, (bits,)); add(259, 3, (8,))
    add(262, 3, (1,)); add(277, 3, (1,)); add(284, 3, (1,))
    add(322, 4, (tw,)); add(323, 4, (th,)); add(339, 3, (kind,))
    for t in _GEO_TAGS:
        if t in tags_t:
            v = tags_t[t]
            if isinstance(v[0], str):
                add(t, 2, v[0])
            else:
                add(t, 12 if t in (33550, 33922, 34264, 34736) else 3, v)
    if nodata is not None:
        add(42113, 2, repr(float(nodata)) if a.dtype.kind == "f" else str(int(nodata)))
    # BigTIFF layout: header(16) | tile data | IFD | out-of-line values
    pos = 16
    offs = []
    for b in blocks:
        offs.append(pos)
        pos += len(b)
    add(324, 16, tuple(offs)); add(325, 16, tuple(len(b) for b in blocks))
    entries.sort(key=lambda e: e[0])
    ifd_off = (pos + 7) // 8 * 8
    extra_off = ifd_off + 8 + 20 * len(entries) + 8
    ifd = struct.pack("<Q", len(entries))
    extra = b""
    for tag, typ, cnt, b in entries:
        if len(b) <= 8:
            ifd += struct.pack("<HHQ", tag, typ, cnt) + b.ljust(8, b"\x00")
        else:
            ifd += struct.pack("<HHQQ", tag, typ, cnt, extra_off + len(extra))
            extra += b + b"\x00" * (-len(b) % 8)
    ifd += struct.pack("<Q", 0)
    with open(dst, "wb") as f:
        f.write(b"II" + struct.pack("<HHHQ", 43, 8, 0, ifd_off))
        for b in blocks:
            f.write(b)
        f.write(b"\x00" * (ifd_off - pos))
        f.write(ifd)
        f.write(extra)


# --------------------------------------------------------------------------
# public boundary
# --------------------------------------------------------------------------
def os_environ_mode() -> str:
    import os
    return os.environ.get("HYDROTOOLS_CUDA_IO", "").lower()


def infer_nodata(dtype) -> Union[float, int, bool]:
    """Nodata value the reference assigns to a data type
    (/root/reference/src/hydrotools/utils.py:242-265): the type's minimum for floats and
    signed integers, its maximum for unsigned integers, False for bool."""
    dt = np.dtype(dtype)
    if dt.kind == "f":
        return float(np.finfo(dt).min)
    if dt.kind == "i":
        return int(np.iinfo(dt).min)
    if dt.kind == "u":
        return int(np.iinfo(dt).max)
    if dt.kind == "b":
        return False
    raise KeyError(dt.name)


def read_raster(src: str) -> Tuple[np.ndarray, dict]:
    """(2-D array with nodata cells holding the nodata value, metadata)."""
    if HAVE_REFERENCE_IO:
        r = _RefRaster(src)
        a = _ref_from_raster(r).compute()
        fill = r.nodata if r.nodata is not None else a.fill_value
        data = np.ma.filled(a, fill)[0]
        meta = dict(nodata=None if r.nodata is None else float(r.nodata), csx=float(r.csx),
                    csy=float(r.csy), left=float(r.left), top=float(r.top),
                    shape=data.shape, dtype=data.dtype.name)
        return data, meta
    a, meta = _fallback_read(src)
    if meta["nodata"] is not None and a.dtype.kind == "f":
        a = np.where(np.isnan(a), a.dtype.type(meta["nodata"]), a)
    return a, meta


def same_grid(src_a: str, src_b: str, tolerance: float = 1e-6) -> bool:
    """``Raster(a).matches(b)`` (/root/reference/src/hydrotools/raster.py:692-714)."""
    if HAVE_REFERENCE_IO:
        return bool(_RefRaster(src_a).matches(src_b, tolerance))
    ma, mb = _fallback_read(src_a)[1], _fallback_read(src_b)[1]
    return (ma["shape"] == mb["shape"]
            and all(np.isclose(ma[k], mb[k], atol=tolerance) for k in ("csx", "csy", "top", "left"))
            and ma["geo_tags"].get(34735) == mb["geo_tags"].get(34735))


def device_io() -> bool:
    """Write results with the device-side tile encoder (hydrotools.cuda._cog) instead of
    ``hydrotools.raster.to_raster``?  Yes when the reference's I/O layer is not importable,
    or when HYDROTOOLS_CUDA_IO=device asks for it."""
    import os
    mode = os.environ.get("HYDROTOOLS_CUDA_IO", "").lower()
    if mode == "device":
        return True
    if mode in ("reference", "host"):
        return False
    return not HAVE_REFERENCE_IO


def write_raster(data, template: str, destination: str, nodata=None,
                 as_cog: bool = True) -> None:
    """Store a (rows, cols) array (numpy, or a CUDA tensor) on the template's grid.
    ``nodata`` cells must already hold ``nodata``.  A CUDA tensor is compressed on the device
    and written as a COG directly when device_io() says so; otherwise it is copied to the
    host and goes the reference's way."""
    import torch
    if isinstance(data, torch.Tensor):
        if data.is_cuda and device_io():
            from . import _cog
            _cog.write_cog(data, _read_ifd(template), destination, nodata=nodata, overviews=as_cog)
            return
        data = data.cpu().numpy()
    if HAVE_REFERENCE_IO and os_environ_mode() != "device":
        import dask.array as da
        from hydrotools.config import CHUNKS  # type: ignore
        arr = da.from_array(data[None], chunks=CHUNKS)
        if nodata is not None:
            arr = da.ma.masked_where(arr == nodata, arr)
            _ref_to_raster(arr, template, destination, as_cog=as_cog, nodata_value=nodata)
        else:
            _ref_to_raster(arr, template, destination, as_cog=as_cog)
        return
    _fallback_write(np.asarray(data), template, destination, nodata)
