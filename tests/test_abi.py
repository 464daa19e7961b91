"""CPU: the C-ABI shared library loads (no GPU needed) and exports every entry
point include/hydrotools_b200.h declares; the host layer fails loudly without a
device instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "hydrotools_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ht_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(hc):
    from hydrotools.cuda import _native
    lib = ctypes.CDLL(_native.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 5
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert lib.ht_version() >= 100


def test_every_bound_symbol_is_declared(hc):
    from hydrotools.cuda import _native
    names = set(declared_symbols())
    for n in _native.SIGNATURES:
        assert n in names


def test_error_strings(hc):
    from hydrotools.cuda import _native
    assert _native.lib.ht_error_string(0) == b"ok"
    assert b"align" in _native.lib.ht_error_string(-2)
    with pytest.raises(ValueError):
        _native.check(-1, "x")
    with pytest.raises(_native.NativeError):
        _native.check(-3, "x")


def test_no_cpu_fallback(hc):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hc.terrain(np.zeros((8, 8), np.float32))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "hydro-tools_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "oracle/" not in src.replace(
                    "oracle/gdaldem.c", "").replace("oracle/rwatershed.c", "").replace(
                    "oracle/flowacc.c", ""), f


def test_lazy_arrays_are_materialised():
    """dask interop of the array-level API: anything with .compute() (dask.array, as
    hydrotools.raster.from_raster returns) is accepted; (1, rows, cols) is squeezed."""
    import numpy as np
    from hydrotools.cuda import _device as D

    class Lazy:
        def __init__(self, a):
            self._a = a

        def compute(self):
            return self._a

    a = np.ma.masked_equal(np.arange(12, dtype=np.float32).reshape(1, 3, 4), 5)
    out = D.as_2d(Lazy(a))
    assert isinstance(out, np.ndarray) and not isinstance(out, np.ma.MaskedArray)
    assert out.shape == (3, 4) and out[1, 1] == 5


def test_pipeline_fails_loudly_without_a_device(hc):
    import torch
    with pytest.raises(ValueError):
        hc.FlowPipeline(0, 8)
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hc.FlowPipeline(8, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        list(hc.flow_direction_accumulation_many([np.zeros((8, 8), np.float32)]))
    assert list(hc.flow_direction_accumulation_many([])) == []


def test_stage_b_mode_hook(hc):
    from hydrotools.cuda import _native
    prev = _native.lib.ht_acc_stage_b_mode(-1)
    assert prev in (0, 1, 2)
    assert _native.lib.ht_acc_stage_b_mode(1) == prev
    assert _native.lib.ht_acc_stage_b_mode(7) == 1      # out of range: unchanged
    assert _native.lib.ht_acc_stage_b_mode(prev) == 1
    assert _native.lib.ht_acc_stage_b_mode(-1) == prev
