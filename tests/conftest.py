import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def href():
    """The C restatement of the reference CPU path (oracle/libh2ref.so), built on demand."""
    import h2ref
    h2ref.lib()
    return h2ref


@pytest.fixture(scope="session")
def spec():
    """The big-integer specification oracle."""
    import bn254
    return bn254


@pytest.fixture(scope="session")
def h2b():
    """The product package, initialised on cuda:0.  Fails loudly when the CUDA library is missing."""
    import halo2_prover_b200 as pkg
    from halo2_prover_b200 import _ffi
    _ffi.init(0)
    return pkg
