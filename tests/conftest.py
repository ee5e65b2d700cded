import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as orc
    orc.build()
    orc.lib()
    return orc


@pytest.fixture(scope="session")
def ctx():
    """One pandrs_b200 device context for the whole GPU session."""
    import pandrs_b200 as pb
    c = pb.Context(device=0)
    yield c
    c.close()
