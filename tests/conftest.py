import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "hotpath_golden.npz")))


@pytest.fixture(scope="session")
def golden64():
    import numpy as np
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "cs64_golden.npz")))
