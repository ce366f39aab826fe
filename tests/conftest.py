import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Builds libekgsim_b200.so + oracle once per session (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()
    import ekgsim_b200
    return ekgsim_b200


@pytest.fixture(scope="session")
def model24():
    import ekgio
    return ekgio.load_model24()


@pytest.fixture(scope="session")
def model24_delay(model24):
    """Reference-pinned activation map of model_24 (oracle, sha256-checked against the golden)."""
    from oracle import oracle
    return oracle.activation(model24["layers"], model24["transfer"])
