import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu through gpurun)")


@pytest.fixture(scope="session")
def jm():
    """The product package with its CUDA library built."""
    import dolfinx_materials_b200 as pkg
    from dolfinx_materials_b200 import build

    build.build_library()
    return pkg
