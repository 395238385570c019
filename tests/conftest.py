import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_binding as ob
    ob.lib()  # builds oracle/libicp_oracle.so if missing
    return ob


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "example_pair.npz")
    return dict(np.load(path))
