import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FIREFOX = "/root/reference/src/test/java/SevenZip/firefox.exe"  # only in the build container
FIREFOX_XZ = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "firefox.exe.xz")  # travels (make_fixture.py)
FIREFOX_MD5 = "5744fff8e72d105c138dae9e17bb29fe"

# (switch, overrides of the CLI defaults d23 lc3 lp0 pb2 fb128 bt4, length, md5) -- LzmaAloneTest.java:27-38
VECTORS = [
    ("", {}, 138940, "93c6983fcfa73e55099a11ee13139687"),
    ("-eos", {"eos": True}, 138946, "4b9287512dcf72b094abafbd5fbfda85"),
    ("-d0", {"dict_size": 1}, 356822, "385ef9694b5d0640fd372c99cec1d575"),
    ("-fb5", {"fb": 5}, 150508, "81b9ab49744b242c4e5a0274ae5a83d3"),
    ("-fb273", {"fb": 273}, 138711, "44e59bfa0128c6dcfde164598e180e92"),
    ("-lc0", {"lc": 0}, 143351, "8ebbd8dc6c1a1dd2c1803659a4a2b978"),
    ("-lc8", {"lc": 8}, 144829, "f7a9f4ce9c7853c07445b41cca75c58c"),
    ("-lp1", {"lp": 1}, 137620, "27fba851ee64468dc5391d4a0f430ab7"),
    ("-lp4", {"lp": 4}, 141530, "377337634457f7017760e45129760c7d"),
    ("-pb0", {"pb": 0}, 142879, "563da117b34b52358e24d6e5b16d093d"),
    ("-pb4", {"pb": 4}, 140046, "cbbff9f4722065bec54336a7d3d49832"),
    ("-mfbt2", {"mf": 0}, 138877, "126f88731f968265bf163b7f7b5521db"),
]
CLI_DEFAULTS = dict(dict_size=1 << 23, lc=3, lp=0, pb=2, fb=128, mf=1, eos=False)


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


@pytest.fixture(scope="session")
def firefox():
    """The input of the reference's 12 golden vectors (LzmaAloneTest.java:25-38): the reference's own
    file where it exists, else the packed copy under tests/golden/ (identical md5 either way)."""
    import hashlib
    import lzma
    if os.path.exists(FIREFOX):
        data = open(FIREFOX, "rb").read()
    else:
        data = lzma.decompress(open(FIREFOX_XZ, "rb").read())
    assert hashlib.md5(data).hexdigest() == FIREFOX_MD5
    return data
