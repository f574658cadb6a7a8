import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the C-ABI library is built in-tree (git-ignored): build it when a fresh checkout has none
    from pdspy_b200.csrc import build as _b
    if _b.needs_build():
        try:
            _b.build()
        except Exception as e:                      # no nvcc: the ABI tests will say so
            print("could not build libpdsb.so:", e)


@pytest.fixture(scope="session")
def fixture720():
    """The reference's only fixture (tests/testdata.hdf5), converted by tests/golden/make_golden.py."""
    d = np.load(os.path.join(GOLDEN, "fixture_720.npz"))
    return {k: d[k] for k in d.files}


@pytest.fixture(scope="session")
def grid_golden():
    d = np.load(os.path.join(GOLDEN, "grid_golden.npz"))
    return {k: d[k] for k in d.files}


@pytest.fixture(scope="session")
def live_ref():
    """The reference's own compiled module (oracle/_ref), or None when it is not available
    (neither prebuilt nor buildable because /root/reference is absent)."""
    from oracle import build_ref
    try:
        return build_ref.load()
    except Exception:
        return None


@pytest.fixture(scope="session")
def gpu():
    """libpdsb with a device initialised; GPU tests call through this (the C-ABI)."""
    from pdspy_b200 import _lib
    try:
        return _lib.lib()
    except _lib.PdsbError as e:                     # no CUDA device here: the GPU suite is for the B200 box
        pytest.skip("no CUDA device (%s)" % e)
