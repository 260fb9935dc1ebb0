import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def built():
    """make sure the in-tree native libraries exist (nvcc cross-compiles without a GPU)"""
    from megaverse_b200 import _build

    _build.build_all()
    import orc

    orc.lib()
    return True
