"""GPU: row-band sharding on ONE device -- P CudaBand objects driven in-process
through the same table / band-graph functions BandEngine uses, against the
oracle on the whole raster.  (Ranks that wait on each other must not run as
separate launches on one GPU, so the collective itself is covered by the gloo
test on CPU and by the multi-GPU bench.)"""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
ND = -32767.0


def run_bands(dem, cuts, nodata, csx, csy, mode=0, weight=None, delta=False):
    from hydrotools.cuda import distributed as HD
    total, cols = dem.shape
    P = len(cuts) - 1
    bands = []
    for b in range(P):
        lo, hi = cuts[b], cuts[b + 1]
        band = HD.CudaBand(hi - lo, cols, lo, total, csx, csy, nodata,
                           has_above=b > 0, has_below=b < P - 1)
        band.load(dem[lo:hi])
        if weight is not None:
            band.set_weight(weight[lo:hi])
        bands.append(band)
    # halo exchange (2 rows), as BandEngine.exchange_halos does with send/recv
    for b in range(P):
        if b > 0:
            bands[b].set_halo_rows(top=True, rows=bands[b - 1].edge_rows(top=False, depth=2))
        if b < P - 1:
            bands[b].set_halo_rows(top=False, rows=bands[b + 1].edge_rows(top=True, depth=2))
    return accumulate_bands(bands, mode, delta)


def accumulate_bands(bands, mode=0, delta=False, direction=True):
    """The exchange of BandEngine.accumulate with the collective replaced by a
    concatenation (every band lives on this one GPU)."""
    from hydrotools.cuda import distributed as HD, _native as N
    P, cols = len(bands), bands[0].cols
    weighted = mode == N.MODE_WEIGHTED
    if weighted:  # the MAX all-reduce of BandEngine.accumulate
        m = torch.stack([band.weight_max() for band in bands]).max().reshape(1)
        for band in bands:
            band.set_weight_max(m)
    tabs = []
    for band in bands:
        if direction:
            band.direction()
        band.local(mode)
        if delta:
            band.mark_edges(mode)
        band.resolve(True)
        if delta:
            band.tile_phases(mode)
        if not weighted:
            tabs.append(band.band_tables(mode))
    if weighted or delta:
        # the kernels BandEngine uses: tables, band-level forest, hand-back of the imports
        every = torch.cat([band.publish_tables(mode) for band in bands])
        if not weighted:  # ... agree with their torch formulation
            all_d = torch.stack([t[0] for t in tabs])
            all_l = torch.stack([t[1] for t in tabs])
            base, nxt = HD.build_band_graph(all_d, all_l)
            for b in range(P):
                assert torch.equal(every[4 * b:4 * b + 2], tabs[b][0])
                assert torch.equal(every[4 * b + 2:4 * b + 4], tabs[b][1].to(torch.int64))
            base2, nxt2 = bands[0].band_graph(every, P)
            assert torch.equal(base2, base) and torch.equal(nxt2, nxt)
        if delta:
            # stage C of the tiles that do not depend on the neighbour bands runs BEFORE the
            # imports arrive (BandEngine overlaps the two); poison the outputs first
            for band in bands:
                band.acc.fill_(-7.0)
                band.dir.fill_(99)
                band.final_phase(0, mode)
        for b, band in enumerate(bands):
            band.import_flows(every, P, b, mode, delta)
    else:
        all_d = torch.stack([t[0] for t in tabs])
        all_l = torch.stack([t[1] for t in tabs])
        base, nxt = HD.build_band_graph(all_d, all_l)
        inflow = bands[0].forest_resolve(base, nxt, mode).view(P, 2, cols)
        for b, band in enumerate(bands):  # second full resolve
            band.add_inflow(inflow[b, 0], inflow[b, 1], mode)
            band.resolve(False)
    outs = []
    for band in bands:
        if delta:
            band.final_phase(1, mode)
        else:
            band.final(mode)
        fd, fa = band.results()
        outs.append((fd.cpu().numpy(), fa.cpu().numpy()))
    return np.vstack([o[0] for o in outs]), np.vstack([o[1] for o in outs])


@pytest.mark.parametrize("delta", [False, True])
@pytest.mark.parametrize("cuts", [[0, 150, 300], [0, 64, 128, 300], [0, 2, 100, 101 + 64, 300],
                                  [0, 37, 38 + 37, 200, 300], [0, 129, 300]])
def test_bands_equal_oracle(hc, cuts, delta):
    dem = oracle.synth_dem(42, 300, 333, with_nulls=True)
    fd, fa = run_bands(dem, cuts, ND, 10.0, 10.0, delta=delta)
    efd, efa = oracle.rwatershed(dem, ND, 10.0, 10.0)
    np.testing.assert_array_equal(fd, efd)
    np.testing.assert_array_equal(fa, efa)


@pytest.mark.parametrize("delta", [False, True])
def test_bands_long_rivers(hc, delta):
    """Rivers cross every band boundary; meanders re-enter bands."""
    dem = oracle.synth_dem(20261019, 1500, 1100)
    cuts = [0, 200, 500, 777, 1100, 1500]
    fd, fa = run_bands(dem, cuts, None, 10.0, 10.0, delta=delta)
    efd, efa = oracle.rwatershed(dem, None, 10.0, 10.0)
    np.testing.assert_array_equal(fd, efd)
    np.testing.assert_array_equal(fa, efa)


@pytest.mark.parametrize("delta", [False, True])
def test_bands_upward_flow(hc, delta):
    """Flip the DEM so that water runs towards row 0: flow crosses band boundaries
    upwards (virtual slots of the band above)."""
    dem = np.ascontiguousarray(oracle.synth_dem(7, 400, 300, with_nulls=True)[::-1])
    assert oracle.check_conditioned(dem, ND) == (0, 0)
    fd, fa = run_bands(dem, [0, 130, 270, 400], ND, 10.0, 10.0, delta=delta)
    efd, efa = oracle.rwatershed(dem, ND, 10.0, 10.0)
    np.testing.assert_array_equal(fd, efd)
    np.testing.assert_array_equal(fa, efa)


@pytest.mark.parametrize("delta", [False, True])
def test_bands_weighted(hc, delta):
    dem = oracle.synth_dem(3, 500, 400, with_nulls=True)
    w = oracle.synth_weight(3, 500, 400)
    from hydrotools.cuda import _native as N
    _, acc = run_bands(dem, [0, 100, 333, 500], ND, 10.0, 10.0, mode=N.MODE_WEIGHTED, weight=w,
                       delta=delta)
    # weighted accumulation follows r.watershed's graph here (edge cells do not pass
    # flow on), so the expectation is built from the oracle's directions with the
    # edge cells made terminal
    efd, _ = oracle.rwatershed(dem, ND, 10.0, 10.0)
    edge = np.zeros(dem.shape, bool)
    null = dem == ND
    edge[[0, -1], :] = True
    edge[:, [0, -1]] = True
    pad = np.pad(null, 1, constant_values=False)
    for dr in (0, 1, 2):
        for dc in (0, 1, 2):
            edge |= pad[dr:dr + dem.shape[0], dc:dc + dem.shape[1]]
    d = efd.copy()
    d[edge & ~null] = 0
    exp = oracle.flowacc(d, w)
    m = ~np.isnan(exp)
    np.testing.assert_allclose(acc[m], exp[m], rtol=1e-5, atol=0)


@pytest.mark.parametrize("cuts", [[0, 100, 333, 500], [0, 64, 66, 300, 500], [0, 500]])
def test_bands_stored_direction(hc, golden, cuts):
    """Row a3 sharded: accumulation over a STORED direction raster in row bands equals the
    unsharded function and plain oracle.flowacc (r.flowaccumulation semantics: every cell
    passes its flow on unless the receiver is NULL / outside)."""
    from hydrotools.cuda import distributed as HD, _native as N
    dem = oracle.synth_dem(3, 500, 400, with_nulls=True)
    w = oracle.synth_weight(3, 500, 400)
    fd, _ = oracle.rwatershed(dem, ND, 10.0, 10.0)
    total, cols = fd.shape
    P = len(cuts) - 1
    for mode, weight in ((N.MODE_WEIGHTED, w), (N.MODE_COUNT, None)):
        bands = []
        for b in range(P):
            lo, hi = cuts[b], cuts[b + 1]
            band = HD.CudaBand(hi - lo, cols, lo, total, 10.0, 10.0, None,
                               has_above=b > 0, has_below=b < P - 1)
            band.set_direction(fd[lo:hi])
            if weight is not None:
                band.set_weight(weight[lo:hi])
            bands.append(band)
        edges = [band.dir_edge_rows() for band in bands]  # BandEngine.load_direction's all-gather
        for b, band in enumerate(bands):
            band.set_dir_halo(edges[b - 1][1] if b > 0 else None, edges[b + 1][0] if b < P - 1 else None)
            band.pack_direction()
        _, acc = accumulate_bands(bands, mode, delta=P > 1, direction=False)
        exp = oracle.flowacc(fd, weight)
        np.testing.assert_array_equal(np.isnan(acc), np.isnan(exp))
        m = ~np.isnan(exp)
        if weight is None:
            np.testing.assert_array_equal(acc[m], exp[m])
        else:
            np.testing.assert_allclose(acc[m], exp[m], rtol=1e-12, atol=0)
            whole = hc.flow_accumulation_arrays(fd, weight)
            np.testing.assert_array_equal(acc, whole)  # exact fixed point: sharding changes nothing


def test_direction_in_parts_equals_whole(hc):
    """BandEngine computes a band's interior rows while the halo rows travel, then its first
    and last tile rows (ht_d8_part_f32): the packed bytes equal the one-call result."""
    from hydrotools.cuda import distributed as HD
    dem = oracle.synth_dem(11, 700, 500, with_nulls=True)
    lo, hi = 200, 500
    outs = []
    for parts in (False, True):
        band = HD.CudaBand(hi - lo, 500, lo, 700, 10.0, 10.0, ND, has_above=True, has_below=True)
        band.load(dem[lo:hi])
        band.set_halo_rows(top=True, rows=torch.from_numpy(dem[lo - 2:lo]).cuda())
        band.set_halo_rows(top=False, rows=torch.from_numpy(dem[hi:hi + 2]).cuda())
        band.packed.zero_()
        if parts:
            e = band.D8_TILE_ROWS
            band.direction_part(e, band.rows - e)
            band.direction_part(0, e)
            band.direction_part(band.rows - e, band.rows)
        else:
            band.direction()
        outs.append(band.packed.clone())
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("cuts", [[0, 300], [0, 2, 100, 101 + 64, 300], [0, 129, 300]])
def test_stage_b_variants_agree(hc, cuts):
    """The flat and the two-level boundary-graph resolve (csrc/ht_acc.cu) leave identical
    directions, counts and signed counts -- forced on a raster far below the size at which
    the two-level variant is chosen automatically, so that partial super-tiles, bands of two
    rows, band-edge marks, roots and the delta resolve all go through it."""
    from hydrotools.cuda import _native as N
    dem = oracle.synth_dem(11, 300, 520, with_nulls=True, nodata=ND)
    res = {}
    prev = N.lib.ht_acc_stage_b_mode(-1)
    try:
        for mode in (1, 2):
            N.lib.ht_acc_stage_b_mode(mode)
            assert N.lib.ht_acc_stage_b_mode(-1) == mode
            if len(cuts) == 2:
                res[mode] = hc.flow_direction_accumulation_arrays(dem, nodata=ND, csx=10.0, csy=10.0,
                                                                  positive_only=False)
            else:
                res[mode] = run_bands(dem, cuts, ND, 10.0, 10.0, delta=True)
    finally:
        N.lib.ht_acc_stage_b_mode(prev)
    np.testing.assert_array_equal(res[1][0], res[2][0])
    np.testing.assert_array_equal(res[1][1], res[2][1])
    efd, efa = oracle.rwatershed(dem, ND, 10.0, 10.0, positive_only=len(cuts) > 2)
    np.testing.assert_array_equal(res[2][0], efd)
    np.testing.assert_array_equal(res[2][1], efa)
