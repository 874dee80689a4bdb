"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports
every symbol include/lzma_b200.h declares; without a GPU nothing computes and
nothing silently falls back."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "lzma_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lzb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lzb):
    L = lzb.lib()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(L, name), "liblzma_b200.so does not export %s" % name
    assert sorted(n for n, _, _ in lzb.ABI) == declared  # the Python binding covers the whole header


def test_version_and_bound(lzb):
    assert b"sm_100a" in lzb.lib().lzb_version()
    assert lzb.enc_bound(0) == 128
    assert lzb.enc_bound(3 << 20) == (3 << 20) + (1 << 20) + 128


def test_no_cpu_fallback(lzb):
    """Without a CUDA device handle creation must fail loudly."""
    if lzb.lib().lzb_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(lzb.LzbError):
        lzb.Decoder()
    with pytest.raises(lzb.LzbError):
        lzb.Encoder()
    assert "no CPU fallback" in lzb.last_error()


def test_product_does_not_reference_oracle():
    """The oracle is test infrastructure: nothing under the package may import, link or load it."""
    pkg = os.path.join(ROOT, "lzma-java_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "lzma_oracle" not in text and "oracle." not in text and "import oracle" not in text, f
    out = os.popen("ldd %s 2>/dev/null" % os.path.join(pkg, "liblzma_b200.so")).read()
    assert "oracle" not in out
