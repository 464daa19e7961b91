"""Downstream consumers of the direction grid.  PINNED: the golden vectors
are outputs of the reference's own numba code (tools/make_stream_goldens.py cuts
`_fa_label_task`, streams.py:320-388, and `_upstream_compute`, morphology.py:50-153, out of
the reference and runs them).  CPU: the numpy restatement (oracle/streams.py) equals the
goldens.  GPU: the kernels of csrc/ht_streams.cu equal both, bit for bit."""
import os

import numpy as np
import pytest

from oracle import streams as ost

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sg():
    z = np.load(os.path.join(ROOT, "tests", "golden", "streams.npz"))
    g = {k: z[k] for k in z.files}
    shape = tuple(int(v) for v in g["shape"])
    g["streams"] = np.unpackbits(g["streams"])[:shape[0] * shape[1]].reshape(shape).astype(bool)
    return g


def test_oracle_reach_labels_equal_reference(sg):
    lab = ost.reach_labels(sg["reaches"], sg["fd"], 0)
    np.testing.assert_array_equal(lab, sg["labels"])
    assert lab.max() == 719


def test_oracle_upstream_mean_equals_reference(sg):
    for k in (0, 1):
        got = ost.upstream_mean(sg["values"], sg["streams"], sg["fd"], float(sg["res"]), float(sg["res"]),
                                float(sg[f"upstream_{k}_distance"]))
        np.testing.assert_array_equal(got, sg[f"upstream_{k}"])  # NaN positions included


@pytest.mark.gpu
def test_reach_labels_gpu(hc, sg):
    from hydrotools.cuda import streams as S
    lab = S.reach_labels(sg["reaches"], sg["fd"], 0)
    assert lab.dtype == np.uint32
    np.testing.assert_array_equal(lab, sg["labels"])


@pytest.mark.gpu
def test_upstream_mean_gpu(hc, sg):
    from hydrotools.cuda import streams as S
    res = float(sg["res"])
    for k in (0, 1):
        got = S.upstream_mean(sg["values"], sg["streams"], sg["fd"], res, res,
                              float(sg[f"upstream_{k}_distance"]))
        np.testing.assert_array_equal(got, sg[f"upstream_{k}"])


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_streams_random_vs_oracle(hc, seed):
    """Random reach classes / stream masks over a synthetic direction grid, ragged shape,
    anisotropic cells."""
    import oracle
    from hydrotools.cuda import streams as S
    rng = np.random.default_rng(seed)
    rows, cols = 150 + 13 * seed, 210 - 7 * seed
    dem = oracle.synth_dem(seed, rows, cols, with_nulls=True)
    fd, fa = oracle.rwatershed(dem, -32767.0, 10.0, 10.0)
    streams = np.nan_to_num(fa, nan=0.0) >= 30
    reaches = np.where(streams, rng.integers(1, 4, fd.shape), 0).astype(np.int32)
    np.testing.assert_array_equal(S.reach_labels(reaches, fd, 0), ost.reach_labels(reaches, fd, 0))
    vals = rng.random(fd.shape) * 100
    got = S.upstream_mean(vals, streams, fd, 3.0, 4.0, 37.0)
    np.testing.assert_array_equal(got, ost.upstream_mean(vals, streams, fd, 3.0, 4.0, 37.0))
    rr, cc = np.nonzero(np.nan_to_num(fa, nan=0.0) == np.nanmax(fa))
    for (r, c) in [(int(rr[0]), int(cc[0])), (rows // 2, cols // 2)]:
        np.testing.assert_array_equal(S.basin_mask(fd, r, c), ost.basin_mask(fd, r, c))


@pytest.mark.gpu
def test_basin_size_equals_accumulation(hc):
    """Size-independent property: the basin of a cell holds exactly as many cells as its
    r.flowaccumulation count."""
    import oracle
    from hydrotools.cuda import streams as S
    dem = oracle.synth_dem(9, 900, 700)
    fd, _ = hc.flow_direction_accumulation_arrays(dem, csx=10.0, csy=10.0)
    cnt = hc.flow_accumulation_arrays(fd)
    for (r, c) in [(450, 350), (880, 20), (10, 690)]:
        assert int(S.basin_mask(fd, r, c).sum()) == int(cnt[r, c])
