import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built_lib():
    """The C-ABI library, built on demand (nvcc cross-compiles without a GPU)."""
    from manus_b200 import build, _lib

    build.build()
    return _lib.lib()


@pytest.fixture(scope="session")
def raster_ref():
    from oracle.raster_ref import RasterRef

    return RasterRef("f32")


@pytest.fixture(scope="session")
def raster_ref64():
    from oracle.raster_ref import RasterRef

    return RasterRef("f64")
