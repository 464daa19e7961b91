"""GPU: the device generator produces the same bits as the CPU build of
include/ht_synth.h (so oracle and CUDA path see identical inputs)."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,cols,row0,total,nulls", [
    (100, 130, 0, 100, False), (64, 257, 49990, 100000, False), (97, 1001, 300, 10000, True)])
def test_device_generator_matches_cpu(hc, rows, cols, row0, total, nulls):
    from hydrotools.cuda import synth
    buf, ld = synth.synth_dem(20261019, rows, cols, row0=row0, total_rows=total,
                              with_nulls=nulls)
    got = buf[:, :cols].cpu().numpy()
    exp = oracle.synth_dem(20261019, rows, cols, row0=row0, total_rows=total, with_nulls=nulls)
    np.testing.assert_array_equal(got, exp)
    wbuf, _ = synth.synth_weight(20261019, rows, cols, row0=row0)
    np.testing.assert_array_equal(wbuf[:, :cols].cpu().numpy(),
                                  oracle.synth_weight(20261019, rows, cols, row0=row0))


def test_halo_rows_are_generated(hc):
    from hydrotools.cuda import synth
    buf, ld = synth.synth_dem(7, 50, 90, row0=50, total_rows=200, extra_top=2, extra_bot=2)
    exp = oracle.synth_dem(7, 54, 90, row0=48, total_rows=200)
    np.testing.assert_array_equal(buf[:, :90].cpu().numpy(), exp)
