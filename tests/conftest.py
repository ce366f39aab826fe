import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests need a CUDA device: skip them (instead of failing) where there is none and the run did not ask
    for them with -m gpu.  The product itself has no CPU fallback; the library says so when called without a device."""
    if "gpu" in (config.getoption("-m") or ""):
        return
    try:
        import ctypes
        lib = ctypes.CDLL(os.path.join(ROOT, "ekgsim_b200", "libekgsim_b200.so"))
        have = lib.ekg_device_count() > 0
    except OSError:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device (run with -m gpu on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


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
