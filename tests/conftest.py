import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle as O
    return O.Oracle()


@pytest.fixture(scope="session")
def orc_libm():
    from oracle import oracle as O
    return O.Oracle(libm=True)


@pytest.fixture(scope="session")
def gpu():
    """The product library on a real device; fails (not skips) if the CUDA extension is missing."""
    import gen_b200
    from gen_b200 import build
    if build.needs_build():
        build.build()
    gen_b200.load()
    return gen_b200
