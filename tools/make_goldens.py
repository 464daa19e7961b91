"""Generate tests/golden/*.npz (run HERE, where /root/reference is mounted).

Inputs are the reference's own fixtures (tests/data-files/dem.tif, precip.tif;
/root/reference/tests/__init__.py:5-13), decoded with PIL because rasterio is
absent.  Outputs are the CPU oracle's results on them -- they pin the ORACLE
against regressions; they are NOT outputs of real GRASS/GDAL (parity unpinned,
see oracle/rwatershed.c).  tools/regen_goldens_with_grass.py is the script to
run on a box that has the real reference stack.

    python tools/make_goldens.py
"""
import hashlib
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

REF = "/root/reference/tests/data-files"
OUT = os.path.join(ROOT, "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    os.makedirs(OUT, exist_ok=True)
    dem = np.array(Image.open(os.path.join(REF, "dem.tif")))
    precip = np.array(Image.open(os.path.join(REF, "precip.tif")))
    dem_nd, precip_nd = -32767.0, float(np.float32(-3.4028234663852886e+38))
    res = 5.0

    # configs[0]: the bundled DEM through r.watershed -s -a (reference defaults)
    fd, fa = oracle.rwatershed(dem, dem_nd, res, res, positive_only=True)
    _, fa_signed = oracle.rwatershed(dem, dem_nd, res, res, positive_only=False)
    slope, aspect, tri = oracle.gdaldem(dem, dem_nd, res, res)
    slope_pct, _, _ = oracle.gdaldem(dem, dem_nd, res, res, scale=2.0, percent=True)
    # the conditioned variant north_star stipulates for the GPU path
    cdem = oracle.condition_by_rank(dem, dem_nd, res, res)
    cfd, cfa = oracle.rwatershed(cdem, dem_nd, res, res, positive_only=True)
    _, cfa_signed = oracle.rwatershed(cdem, dem_nd, res, res, positive_only=False)
    # weighted accumulation of the bundled precipitation over the conditioned
    # DEM's direction grid (FlowAccumulation.calculate semantics)
    wacc = oracle.flowacc(cfd, precip, precip_nd)
    cnt = oracle.flowacc(cfd)

    np.savez_compressed(
        os.path.join(OUT, "bundled.npz"),
        dem=dem, precip=precip, dem_nodata=dem_nd, precip_nodata=precip_nd, res=res,
        fd=fd, fa=fa.astype(np.float32), neg=np.packbits(fa_signed < 0),
        cfd=cfd, cfa=cfa.astype(np.float32), cneg=np.packbits(cfa_signed < 0))
    # everything else is pinned by hash only (tests recompute it with the oracle)
    with open(os.path.join(OUT, "bundled.sha256"), "w") as f:
        for name, arr in [("fd", fd), ("fa", fa), ("slope", slope), ("aspect", aspect),
                          ("tri", tri), ("slope_pct", slope_pct), ("cdem", cdem),
                          ("cfd", cfd), ("cfa", cfa), ("wacc", wacc), ("cnt", cnt)]:
            f.write(f"{sha(arr)}  {name}\n")
    print("max fa", np.nanmax(fa), "max cfa", np.nanmax(cfa),
          "size", os.path.getsize(os.path.join(OUT, "bundled.npz")))


if __name__ == "__main__":
    main()
