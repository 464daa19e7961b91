import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hydro-tools_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    """Reference fixtures (tests/data-files/dem.tif, precip.tif of the
    reference) + oracle outputs, written by tools/make_goldens.py."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "bundled.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden_hashes():
    out = {}
    with open(os.path.join(ROOT, "tests", "golden", "bundled.sha256")) as f:
        for line in f:
            h, name = line.split()
            out[name] = h
    return out


@pytest.fixture(scope="session")
def hc():
    """The product package (fails loudly if the CUDA library is not built)."""
    import hydrotools.cuda as m
    return m
