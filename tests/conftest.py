import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")
    config.addinivalue_line("markers", "own_lanes: the test picks the forward kernel's lanes itself (not repeated per lane setting)")


def pytest_generate_tests(metafunc):
    # every GPU test runs once per forward-kernel mapping: 1 lane per (chain, block) (the path large ensembles take), 4 lanes,
    # and automatic (8 lanes at the test sizes); results must not depend on it
    if metafunc.definition.get_closest_marker("gpu") is not None and metafunc.definition.get_closest_marker("own_lanes") is None:
        metafunc.fixturenames.append("_fwd_lanes")
        metafunc.parametrize("_fwd_lanes", [1, 4, 0], ids=["lanes1", "lanes4", "lanes_auto"], indirect=True)


@pytest.fixture
def _fwd_lanes(request):
    import dmt_b200
    from dmt_b200 import _lib
    _lib.DEFAULT_FWD_LANES = request.param
    yield request.param
    _lib.DEFAULT_FWD_LANES = 0


@pytest.fixture(scope="session")
def orc():
    from oracle import orc as _orc
    _orc.build()
    return _orc


@pytest.fixture(scope="session")
def olib(orc):
    return orc.load()
