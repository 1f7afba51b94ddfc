import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    import pyoracle

    return pyoracle.Port()


@pytest.fixture(scope="session")
def ref():
    import pyoracle

    if not pyoracle.Ref.available() and not os.path.isdir("/root/reference"):
        pytest.skip("oracle/_ref/libsqref.so not built and /root/reference absent")
    return pyoracle.Ref()


@pytest.fixture(scope="session")
def sq():
    import squander_b200

    return squander_b200
