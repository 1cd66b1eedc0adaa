import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ngp():
    """The product package (directory name contains a hyphen, hence importlib)."""
    import importlib

    return importlib.import_module("instant-ngp_b200")


@pytest.fixture(scope="session")
def built_lib():
    import __graft_entry__ as g

    g.build()
    import importlib

    return importlib.import_module("instant-ngp_b200").load_library()
