"""Golden vectors for the moving-window filter from the REFERENCE'S OWN CODE.

Runs only where /root/reference exists (the build container).  The reference
module cannot be imported (dask, rasterio, dask_image, sklearn are absent), so the
numba kernel `apply_filter` (src/hydrotools/interpolate.py:42-64) and the reducers
of `raster_filter` (:147-222) are cut out of the file with `ast` and executed
unchanged; `da.map_overlap(depth, boundary=nan)` + trim (:94-107) is reproduced by
NaN-padding a single chunk by (i_start, j_start) on both sides.

    python tools/make_filter_goldens.py      -> tests/golden/filter.npz
"""
import ast
import os
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src/hydrotools/interpolate.py"
sys.path.insert(0, ROOT)


def _no_jit(*args, **kwargs):
    if args and callable(args[0]):
        return args[0]
    return lambda f: f


class _PyDict:
    @staticmethod
    def empty(key_type=None, value_type=None):
        return {}


def reference_functions():
    import numba
    src = open(REF).read()
    tree = ast.parse(src)
    top = {n.name: n for n in tree.body if isinstance(n, ast.FunctionDef)}
    apply_src = next(ast.get_source_segment(src, n) for n in top["custom_raster_filter"].body
                     if isinstance(n, ast.FunctionDef) and n.name == "apply_filter")
    ns = {"np": np, "njit": numba.njit}
    exec(textwrap.dedent(apply_src), ns)
    reducers = {}
    # if method == "sum": @njit def func(sample): ...   (an if / elif chain)
    node = next(n for n in top["raster_filter"].body if isinstance(n, ast.If))
    while node is not None:
        name = node.test.comparators[0].value
        fsrc = textwrap.dedent(ast.get_source_segment(src, node.body[0]))
        fns = {"np": np, "njit": numba.njit}
        if name == "mode":
            # `Dict.empty(key_type=np.float64, ...)` (interpolate.py:178-181) does not type with
            # the numba of this image, so the reference's source of this one reducer -- and of
            # apply_filter for it -- is executed as plain Python, with a dict (insertion-ordered
            # like numba's typed Dict) standing in for `Dict.empty(...)`
            fns = {"np": np, "njit": _no_jit, "Dict": _PyDict}
        try:
            exec(fsrc, fns)
            reducers[name] = fns["func"]
        except Exception:
            pass
        node = node.orelse[0] if node.orelse and isinstance(node.orelse[0], ast.If) else None
    py_ns = {"np": np, "njit": _no_jit}
    exec(textwrap.dedent(apply_src), py_ns)
    reducers["__py_apply_filter__"] = py_ns["apply_filter"]
    return ns["apply_filter"], reducers


def reference_filter(apply_filter, func, a, kernel):
    """custom_raster_filter's array path for one chunk."""
    kernel = np.ones(kernel, bool) if isinstance(kernel, tuple) else np.array(kernel, dtype=bool)
    i_start = int(np.ceil((kernel.shape[0] - 1) / 2.0))
    i_end = kernel.shape[0] - 1 - i_start
    j_start = int(np.ceil((kernel.shape[1] - 1) / 2.0))
    j_end = kernel.shape[1] - 1 - j_start
    p = np.pad(a[None].astype(np.float32), ((0, 0), (i_start, i_start), (j_start, j_start)),
               constant_values=np.nan)
    out = apply_filter(p, kernel, i_start, j_start, i_end, j_end, func, ())
    return out[0, i_start:i_start + a.shape[0], j_start:j_start + a.shape[1]]


def disc(radius):
    y, x = np.mgrid[-radius:radius + 1, -radius:radius + 1]
    return (x * x + y * y) <= radius * radius


def cases():
    rng = np.random.default_rng(20261019)
    z = np.load(os.path.join(ROOT, "tests", "golden", "bundled.npz"))
    dem = z["dem"].astype(np.float32)
    dem[dem == float(z["dem_nodata"])] = np.nan
    a1 = (rng.random((37, 53)) * 100).astype(np.float32)
    a1[rng.random(a1.shape) < 0.1] = np.nan
    ring = disc(2)
    ring[2, 2] = False  # kernel without its centre
    classes = rng.integers(0, 5, (41, 47)).astype(np.float32)  # repeated values: mode / median ties
    classes[rng.random(classes.shape) < 0.15] = np.nan
    return {
        "classes_3x3": (classes, (3, 3)),
        "classes_4x5": (classes, (4, 5)),
        "classes_disc3": (classes, disc(3)),
        "rand_3x3": (a1, (3, 3)),
        "rand_4x2": (a1, (4, 2)),          # even sizes: i_start != i_end
        "rand_1x5": (a1, (1, 5)),
        "rand_disc3": (a1, disc(3)),
        "rand_ring": (a1, ring),
        "dem_5x5": (dem[100:220, 300:460], (5, 5)),
        "dem_disc6": (dem[100:220, 300:460], disc(6)),
    }


def main():
    apply_filter, reducers = reference_functions()
    methods = ["sum", "min", "max", "diff", "mean", "std", "variance", "median", "mode"]
    out = {}
    for name, (a, kernel) in cases().items():
        out[f"{name}__in"] = a
        out[f"{name}__kernel"] = (np.ones(kernel, bool) if isinstance(kernel, tuple)
                                  else np.asarray(kernel, bool))
        for m in methods:
            if m == "mode" and not name.startswith("classes"):
                continue  # interpreted: only the small class rasters
            af = reducers["__py_apply_filter__"] if m == "mode" else apply_filter
            out[f"{name}__{m}"] = reference_filter(af, reducers[m], a, kernel)
    dst = os.path.join(ROOT, "tests", "golden", "filter.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, f"({os.path.getsize(dst) / 1024:.0f} KiB, {len(out)} arrays)")


if __name__ == "__main__":
    main()
