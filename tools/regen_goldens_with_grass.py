"""Pin the oracle against the REAL reference (run on a box that has GRASS >= 8.3,
GDAL 3.11, rasterio, dask and the reference package installed -- none of which
exist in the build image, which is why parity is recorded as "unpinned").

    python tools/regen_goldens_with_grass.py [--time]

Runs the reference's own functions
  hydrotools.watershed.flow_direction_accumulation   (src/hydrotools/watershed.py:30-75)
  hydrotools.elevation.slope / aspect / terrain_ruggedness_index
                                                     (src/hydrotools/elevation.py:24-128)
on tests/data-files/dem.tif and on the rank-conditioned variant, and diffs their
rasters against oracle/ (rules R1-R6, G1-G6).  Any difference names the rule to
correct in oracle/*.c; the CUDA kernels follow the oracle.
"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402


def main():
    try:
        from hydrotools.watershed import flow_direction_accumulation
        from hydrotools.elevation import slope, aspect, terrain_ruggedness_index
        from hydrotools.raster import Raster, from_raster, to_raster
    except Exception as e:  # pragma: no cover
        raise SystemExit(f"the reference stack is not importable here: {e}")
    import dask.array as da
    ref_dem = os.environ.get("HT_REF_DEM", "/root/reference/tests/data-files/dem.tif")
    r = Raster(ref_dem)
    dem = np.ma.filled(from_raster(r).compute(), r.nodata)[0].astype(np.float32)
    nd, res = float(r.nodata), float(r.csx)
    with tempfile.TemporaryDirectory() as tmp:
        cases = {"bundled": ref_dem}
        cdem = oracle.condition_by_rank(dem, nd, res, res)
        cpath = os.path.join(tmp, "cdem.tif")
        to_raster(da.ma.masked_equal(da.from_array(cdem[None]), nd), ref_dem, cpath,
                  nodata_value=nd)
        cases["conditioned"] = cpath
        for name, path in cases.items():
            a = np.ma.filled(from_raster(path).compute(), nd)[0].astype(np.float32)
            fdp, fap = os.path.join(tmp, "fd.tif"), os.path.join(tmp, "fa.tif")
            t0 = time.perf_counter()
            flow_direction_accumulation(path, fdp, fap)
            t_ref = time.perf_counter() - t0
            fd = from_raster(fdp).compute()[0]
            fa = from_raster(fap).compute()[0]
            t0 = time.perf_counter()
            efd, efa = oracle.rwatershed(a, nd, res, res)
            t_ora = time.perf_counter() - t0
            valid = ~np.ma.getmaskarray(fd)
            print(f"[{name}] r.watershed: direction mismatches "
                  f"{(np.ma.filled(fd, 0)[valid] != efd[valid]).sum()} of {valid.sum()}, "
                  f"accumulation mismatches {(np.ma.filled(fa, 0)[valid] != efa[valid]).sum()}; "
                  f"reference {t_ref:.2f} s, oracle {t_ora:.2f} s")
            es, ea, et = oracle.gdaldem(a, nd, res, res)
            for fn, exp, tag in ((slope, es, "slope"), (aspect, ea, "aspect"),
                                 (terrain_ruggedness_index, et, "tri")):
                out = os.path.join(tmp, tag + ".tif")
                fn(path, out)
                g = np.ma.filled(from_raster(out).compute(), -9999.0)[0]
                ok = exp != -9999
                print(f"[{name}] gdaldem {tag}: mask equal {np.array_equal(g == -9999, ~ok)}, "
                      f"max abs diff {np.abs(g[ok] - exp[ok]).max():.3g}")


if __name__ == "__main__":
    main()
