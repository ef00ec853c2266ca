import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gaussian-splatting_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_python.npz"))


@pytest.fixture(scope="session")
def host_lib(tmp_path_factory):
    """Path of a host build of the kernel SOURCES (tests/host_emul): test infrastructure for boxes without a GPU."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "host_emul"))
    import build as host_build
    try:
        return host_build.build_library(str(tmp_path_factory.mktemp("host_emul")))
    except RuntimeError as e:
        if "no host C++ compiler" in str(e):
            pytest.skip(str(e))
        raise


@pytest.fixture()
def on_host(host_lib):
    """The Python layer driven by CPU tensors against the host build for the duration of one test."""
    import build as host_build
    with host_build.python_layer_on_host(host_lib) as dgr:
        yield dgr
