import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FIREFOX = "/root/reference/src/test/java/SevenZip/firefox.exe"  # only in the build container


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU parity oracle (test infrastructure)."""
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def corpus():
    from tools import corpus as Cp
    Cp.build()
    return Cp


@pytest.fixture(scope="session")
def lzb():
    """The product package; builds the C-ABI library if it is missing."""
    mod = importlib.import_module("lzma-java_b200")
    if not os.path.exists(mod.SO_PATH):
        importlib.import_module("lzma-java_b200.build").build()
    return mod
