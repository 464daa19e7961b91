"""Golden vectors for the downstream consumers of the direction grid from
the REFERENCE'S OWN numba code.

Runs only where /root/reference exists (the build container).  The reference modules cannot
be imported (dask, rasterio, fiona are absent), so the two numba kernels are cut out of their
files with `ast` and executed unchanged:
  * `_fa_label_task`     /root/reference/src/hydrotools/streams.py:320-388
  * `_upstream_compute`  /root/reference/src/hydrotools/morphology.py:50-153
Inputs: the oracle's direction / accumulation grids of the reference's bundled DEM
(tests/golden/bundled.npz), streams = accumulation above a threshold, reach values = a
coarse class of the accumulation, with a 1-cell NULL frame (the numba code indexes
neighbours without bounds checks).

    python tools/make_stream_goldens.py      -> tests/golden/streams.npz
"""
import ast
import os
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def cut(path, name):
    import numba
    src = open(path).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"np": np, "njit": numba.njit}
    exec(textwrap.dedent(ast.get_source_segment(src, node)), ns)
    return ns[name]


def inputs():
    z = np.load(os.path.join(ROOT, "tests", "golden", "bundled.npz"))
    fd = z["fd"].astype(np.int32)
    fa = z["fa"].astype(np.float64)
    fd[fd == -128] = -1                      # from_raster(...).filled(-1), streams.py:392 (0 at morphology.py:337)
    fd[[0, -1], :] = -1
    fd[:, [0, -1]] = -1
    streams = np.nan_to_num(fa, nan=0.0) >= 150.0
    streams[[0, -1], :] = False
    streams[:, [0, -1]] = False
    reaches = np.where(streams, np.floor(np.log2(np.maximum(fa, 1.0))).astype(np.int32), 0).astype(np.int32)
    dem = z["dem"].astype(np.float64)
    dem[dem == float(z["dem_nodata"])] = np.nan
    return fd, fa, streams, reaches, dem, float(z["res"])


def main():
    fa_label = cut("/root/reference/src/hydrotools/streams.py", "_fa_label_task")
    upstream = cut("/root/reference/src/hydrotools/morphology.py", "_upstream_compute")
    fd, fa, streams, reaches, dem, res = inputs()
    out = {"fd": fd.astype(np.int8), "streams": np.packbits(streams), "shape": np.array(fd.shape),
           "reaches": reaches, "res": np.float64(res)}
    out["labels"] = fa_label(reaches, fd, 0)
    # upstream_stats (morphology.py:292-372): values = elevation on the streams
    fd_off_i = np.array([0, -1, -1, -1, 0, 1, 1, 1, 0], dtype=np.int32)
    fd_off_j = np.array([0, 1, 0, -1, -1, -1, 0, 1, 1], dtype=np.int32)
    values = np.where(streams, np.nan_to_num(dem, nan=0.0), 0.0)
    out["values"] = values
    for k, distance in enumerate((25.0, 200.0)):
        max_window = int(np.ceil(distance / min(res, res))) + 2
        g = np.full(fd.shape, np.nan, np.float32)   # morphology.py:343
        g = upstream(values, streams, fd, g, fd_off_i, fd_off_j, res, res, distance, max_window)
        out[f"upstream_{k}"] = g
        out[f"upstream_{k}_distance"] = np.float64(distance)
    dst = os.path.join(ROOT, "tests", "golden", "streams.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, f"({os.path.getsize(dst) / 1024:.0f} KiB)", "labels", int(out["labels"].max()),
          "stream cells", int(streams.sum()))


if __name__ == "__main__":
    main()
