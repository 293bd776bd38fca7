import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run under gpurun on a B200)")


@pytest.fixture(scope="session")
def lenna():
    import numpy as np
    from PIL import Image

    return np.asarray(Image.open(os.path.join(ROOT, "tests", "golden", "lenna_512_rgb.png")).convert("RGB"))


@pytest.fixture(scope="session")
def fanlin():
    """The product package (directory name has a hyphen; loaded by path)."""
    import __graft_entry__ as G

    return G.load_package()
