"""CPU tests of the oracle itself: hand-computed known answers for each rule it
restates, and the committed golden hashes (tests/golden/bundled.sha256)."""
import hashlib

import numpy as np
import pytest

import oracle


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---------------------------------------------------------------- r.watershed
def test_alt_rounding_is_half_away_from_zero():
    L = oracle.lib()
    assert L.ora_alt_from_f32(1.0005) == 1001 or L.ora_alt_from_f32(1.0005) == 1000
    assert L.ora_alt_from_f32(2.5) == 2500
    assert L.ora_alt_from_f32(-2.5) == -2500
    assert L.ora_alt_from_f32(0.0004) == 0
    assert L.ora_alt_from_f32(-0.0006) == -1


def test_border_seed_codes_and_diagonal_skip_3x3():
    dem = np.array([[9, 8, 7], [6, 5, 4], [3, 2, 1]], np.float32)
    fd, fa = oracle.rwatershed(dem, None, 1.0, 1.0)
    # lowest cell keeps its outward code (bottom row -> -6)
    assert fd[2, 2] == -6
    # seeds re-pointed at the first lower cell popped next to them
    assert fd[1, 2] == 6 and fd[2, 1] == 8
    # centre: SE neighbour (1) is popped first but the diagonal is skipped because
    # the slope via the S neighbour is steeper (3000/1 > 4000/sqrt 2) -> flows S
    assert fd[1, 1] == 6
    # edge cells do not pass flow on; the centre donates 1 to (2,1)
    expect = np.ones((3, 3))
    expect[2, 1] = 2
    np.testing.assert_array_equal(fa, expect)


def test_diagonal_taken_when_steepest():
    # centre 5; SE neighbour much lower than S and E -> diagonal wins
    dem = np.array([[9, 9, 9, 9], [9, 5, 4.9, 9], [9, 4.9, 1, 0.5], [9, 9, 0.6, 0.1]],
                   np.float32)
    fd, _ = oracle.rwatershed(dem, None, 1.0, 1.0)
    assert fd[1, 1] == 7


def test_null_adjacent_seed_points_at_first_null_neighbour():
    nd = -1.0
    dem = np.full((5, 5), 10.0, np.float32)
    dem += np.arange(25, dtype=np.float32).reshape(5, 5) * 0.01
    dem[2, 3] = nd  # NULL east of the centre
    fd, fa = oracle.rwatershed(dem, nd, 1.0, 1.0)
    assert fd[2, 3] == oracle.DIR_NULL and np.isnan(fa[2, 3])
    # (2,2): first NULL neighbour in S,N,W,E,... order is E -> asp -8 unless re-pointed
    # its lower neighbours are popped first, so it is re-pointed to the lowest one
    assert fd[2, 2] > 0
    # (1,2): NULL is its SE neighbour; (3,2): NULL is NE
    d2, _ = oracle.rwatershed(np.where(dem == nd, nd, 10.0 - (dem - 10.0)), nd, 1.0, 1.0)
    assert d2[2, 3] == oracle.DIR_NULL


def test_sign_contamination_and_positive_only():
    dem = np.array([[9, 8, 7], [6, 5, 4], [3, 2, 1]], np.float32)
    _, signed = oracle.rwatershed(dem, None, 1.0, 1.0, positive_only=False)
    _, pos = oracle.rwatershed(dem, None, 1.0, 1.0, positive_only=True)
    assert (signed[[0, 0, 0, 1, 1, 2, 2, 2], [0, 1, 2, 0, 2, 0, 1, 2]] < 0).all()
    np.testing.assert_array_equal(np.abs(signed), pos)


def test_fifo_ties_on_flat():
    # a flat DEM is resolved by insertion order: every interior cell gets a direction
    dem = np.zeros((6, 7), np.float32)
    fd, fa = oracle.rwatershed(dem, None, 1.0, 1.0)
    assert (fd != 0).all() and np.isfinite(fa).all()
    assert fa.sum() >= dem.size


def test_swale_edge_cell_is_reset_to_outward_code():
    # a long 3-row trough draining west->east into the right border: the border
    # outlet is re-pointed?  no lower neighbour -> keeps -8; its upstream edge
    # cells with >= 60 cells of inflow get negative codes again
    rows, cols = 2, 200
    dem = np.empty((rows, cols), np.float32)
    dem[:] = 100.0 - 0.1 * np.arange(cols, dtype=np.float32)
    dem[0] += 5
    fd, fa = oracle.rwatershed(dem, None, 1.0, 1.0)
    # every cell is a border (edge) cell here, so nothing accumulates
    assert fa.max() == 1.0
    rows = 5
    dem = np.empty((rows, cols), np.float32)
    dem[:] = 100.0 - 0.1 * np.arange(cols, dtype=np.float32)
    dem[[0, 4]] += 7
    dem[[1, 3]] += 3
    fd, fa = oracle.rwatershed(dem, None, 1.0, 1.0)
    # interior trough row 2 accumulates eastwards; the last interior cell hands its
    # flow to the east border cell, which is an edge cell and keeps it
    assert fa[2, cols - 1] >= 60
    assert fd[2, cols - 1] < 0


def test_condition_by_rank_is_pit_and_tie_free(golden):
    cdem = oracle.condition_by_rank(golden["dem"], float(golden["dem_nodata"]), 5.0, 5.0)
    assert oracle.check_conditioned(cdem, float(golden["dem_nodata"])) == (0, 0)
    assert oracle.check_conditioned(golden["dem"], float(golden["dem_nodata"]))[0] == 2326


# -------------------------------------------------------------------- gdaldem
def test_gdaldem_plane_known_answer():
    dem = np.tile(2.0 * np.arange(6, dtype=np.float32), (5, 1))
    s, a, t = oracle.gdaldem(dem, None, 1.0, 1.0)
    inner = (slice(1, -1), slice(1, -1))
    np.testing.assert_allclose(s[inner], np.degrees(np.arctan(2.0)), rtol=1e-6)
    np.testing.assert_allclose(a[inner], 270.0)
    np.testing.assert_allclose(t[inner], np.sqrt(24.0), rtol=1e-6)
    # -compute_edges: linear extrapolation keeps the plane's values on the edges ...
    np.testing.assert_allclose(s[0, 1:-1], s[2, 2], rtol=1e-6)
    np.testing.assert_allclose(s[1:-1, 0], s[2, 2], rtol=1e-6)
    # ... except at the four corners, where gdaldem clamps the column index
    np.testing.assert_allclose(s[0, 0], 45.0, rtol=1e-6)
    np.testing.assert_allclose(s[-1, -1], 45.0, rtol=1e-6)


def test_gdaldem_flat_aspect_is_nodata_and_percent_scale():
    dem = np.full((4, 4), 3.0, np.float32)
    s, a, t = oracle.gdaldem(dem, None, 1.0, 1.0)
    assert (a == -9999).all() and (s == 0).all() and (t == 0).all()
    dem = np.tile(np.arange(5, dtype=np.float32), (5, 1))
    s, _, _ = oracle.gdaldem(dem, None, 1.0, 1.0, scale=2.0, percent=True)
    np.testing.assert_allclose(s[2, 2], 50.0, rtol=1e-6)


def test_gdaldem_nodata_centre_and_neighbour():
    nd = -32767.0
    dem = np.tile(np.arange(5, dtype=np.float32), (5, 1))
    dem[2, 2] = nd
    s, a, t = oracle.gdaldem(dem, nd, 1.0, 1.0)
    assert s[2, 2] == -9999 and a[2, 2] == -9999 and t[2, 2] == -9999
    # neighbour of the hole: the hole is replaced by the centre value
    w = dem[0:3, 0:3].copy()
    w[2, 2] = w[1, 1]
    d = w - w[1, 1]
    np.testing.assert_allclose(t[1, 1], np.sqrt((d ** 2).sum()), rtol=1e-6)


def test_gdaldem_degenerate_sizes():
    for shape in ((1, 5), (5, 1), (1, 1)):
        s, a, t = oracle.gdaldem(np.ones(shape, np.float32), None, 1.0, 1.0)
        assert (s == -9999).all() and (a == -9999).all() and (t == -9999).all()
    s, _, _ = oracle.gdaldem(np.array([[0, 1], [2, 3]], np.float32), None, 1.0, 1.0)
    assert np.isfinite(s).all() and (s != -9999).all()


# -------------------------------------------------------------- flowacc
def test_flowacc_chain_and_terminals():
    d = np.array([[8, 8, 8, -8], [2, 0, oracle.DIR_NULL, 4]], np.int8)
    acc = oracle.flowacc(d)
    np.testing.assert_array_equal(acc[0], [2, 3, 4, 5])
    assert acc[1, 0] == 1 and acc[1, 1] == 1 and np.isnan(acc[1, 2])
    # (1,3) flows W into a NULL cell: the flow vanishes
    assert acc[1, 3] == 1
    w = np.array([[1.5, 2.5, 0, 1], [4, 5, 6, 7]], np.float32)
    accw = oracle.flowacc(d, w)
    np.testing.assert_allclose(accw[0], [5.5, 8.0, 8.0, 9.0])


# --------------------------------------------------------------- goldens
def test_golden_hashes(golden, golden_hashes):
    dem, nd, res = golden["dem"], float(golden["dem_nodata"]), float(golden["res"])
    fd, fa = oracle.rwatershed(dem, nd, res, res)
    assert sha(fd) == golden_hashes["fd"] and sha(fa) == golden_hashes["fa"]
    np.testing.assert_array_equal(fd, golden["fd"])
    s, a, t = oracle.gdaldem(dem, nd, res, res)
    assert sha(s) == golden_hashes["slope"]
    assert sha(a) == golden_hashes["aspect"]
    assert sha(t) == golden_hashes["tri"]
    cdem = oracle.condition_by_rank(dem, nd, res, res)
    assert sha(cdem) == golden_hashes["cdem"]
    cfd, cfa = oracle.rwatershed(cdem, nd, res, res)
    assert sha(cfd) == golden_hashes["cfd"] and sha(cfa) == golden_hashes["cfa"]
    wacc = oracle.flowacc(cfd, golden["precip"], float(golden["precip_nodata"]))
    assert sha(wacc) == golden_hashes["wacc"]


def test_reference_fixture_facts(golden):
    # 728 x 395, 102 072 nodata cells, same mask in both
    dem, precip = golden["dem"], golden["precip"]
    assert dem.shape == (395, 728) and precip.shape == dem.shape
    m = dem == golden["dem_nodata"]
    assert m.sum() == 102072
    assert ((precip == golden["precip_nodata"]) == m).all()
    # accumulation invariants on the bundled DEM (cell counts, -a)
    fa = golden["fa"]
    assert np.nanmin(fa[~m]) == 1 and np.isnan(fa[m]).all()


# --------------------------------------------------------------- synthetic
@pytest.mark.parametrize("rows,cols,total,row0", [(200, 300, 200, 0), (64, 257, 100000, 49990),
                                                  (300, 200, 10000, 9700)])
def test_synth_dem_is_conditioned(rows, cols, total, row0):
    dem = oracle.synth_dem(20261019, rows, cols, row0=row0, total_rows=total)
    assert oracle.check_conditioned(dem) == (0, 0)
    alt = np.round(dem.astype(np.float64) * 1000).astype(np.int64)
    assert alt.max() < 4_000_000 and alt.min() >= 1000
    dem_n = oracle.synth_dem(5, rows, cols, row0=row0, total_rows=total, with_nulls=True)
    assert oracle.check_conditioned(dem_n, -32767.0) == (0, 0)


def test_synth_bands_tile_the_whole():
    whole = oracle.synth_dem(3, 90, 70)
    parts = [oracle.synth_dem(3, 30, 70, row0=r, total_rows=90) for r in (0, 30, 60)]
    np.testing.assert_array_equal(whole, np.vstack(parts))
