"""GPU: a stream of rasters through FlowPipeline (upload, kernels and download of
neighbouring rasters overlapped) leaves exactly what the one-raster call and the oracle do."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

ND = -32767.0
ROWS, COLS = 333, 517


def rasters():
    out = []
    for seed in (3, 4, 5):
        out.append(oracle.synth_dem(seed, ROWS, COLS, with_nulls=seed == 4, nodata=ND))
    # rounded to decimetres: ties and pits, the general kernel chain
    raw = oracle.synth_dem(9, ROWS, COLS, nodata=ND)
    out.insert(2, (np.round(raw * 10) / 10).astype(np.float32))
    out.append(oracle.synth_dem(11, ROWS, COLS, nodata=ND))
    return out


@pytest.mark.parametrize("positive_only", [True, False])
@pytest.mark.parametrize("depth", [2, 3])
def test_pipeline_equals_single_calls(hc, depth, positive_only):
    dems = rasters()
    pipe = hc.FlowPipeline(ROWS, COLS, nodata=ND, csx=10.0, csy=10.0, positive_only=positive_only,
                           depth=depth)
    # one page-locked input, the others pageable (staged)
    feed = [torch.from_numpy(dems[0]).pin_memory()] + dems[1:]
    got = [(fd.copy(), fa.copy()) for fd, fa in pipe.run(feed)]
    assert len(got) == len(dems)
    for dem, (fd, fa) in zip(dems, got):
        efd, efa = oracle.rwatershed(dem, ND, 10.0, 10.0, positive_only=positive_only)
        np.testing.assert_array_equal(fd, efd)
        np.testing.assert_array_equal(fa, efa)
        sfd, sfa = hc.flow_direction_accumulation_arrays(dem, nodata=ND, csx=10.0, csy=10.0,
                                                         positive_only=positive_only)
        np.testing.assert_array_equal(fd, sfd)
        np.testing.assert_array_equal(fa, sfa)


def test_pipeline_many_and_outs(hc):
    dems = rasters()[:3]
    res = list(hc.flow_direction_accumulation_many(dems, nodata=ND, csx=10.0, csy=10.0))
    outs = [(torch.empty((ROWS, COLS), dtype=torch.int8, pin_memory=True),
             torch.empty((ROWS, COLS), dtype=torch.float64, pin_memory=True)) for _ in dems]
    pipe = hc.FlowPipeline(ROWS, COLS, nodata=ND, csx=10.0, csy=10.0, own_outputs=False)
    for (fd, fa), (ofd, ofa), (rfd, rfa) in zip(pipe.run(dems, outs), outs, res):
        assert fd.ctypes.data == ofd.numpy().ctypes.data  # landed in the caller's buffers
        np.testing.assert_array_equal(fd, rfd)
        np.testing.assert_array_equal(fa, rfa)
    with pytest.raises(ValueError):
        list(pipe.run([np.zeros((3, 3), np.float32)]))


def test_pipeline_empty(hc):
    assert list(hc.flow_direction_accumulation_many([])) == []
