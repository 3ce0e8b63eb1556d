import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))   # tests are allowed to use the oracle (checker only)
sys.path.insert(0, str(ROOT / "tests"))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def orc():
    import pyoracle

    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def tmc():
    """The product binding.  Builds nothing: the library must already be in-tree."""
    import tiny_mc_b200

    tiny_mc_b200.load()
    return tiny_mc_b200


@pytest.fixture(scope="session")
def gpu(tmc):
    """Library initialised on one B200; fails loudly (no skip, no fallback) without one."""
    tmc.init(1)
    yield tmc
    tmc.finalize()
