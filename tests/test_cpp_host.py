"""The C++ mirror of the reference classes (include/lzma_b200.hpp): it must compile and link
against the C-ABI library everywhere, fail loudly without a GPU, and on a B200 behave like
`LzmaAlone e` / `LzmaAlone d` (LzmaAlone.java:190-239)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "lzma_alone")


def _build(lzb):
    src = os.path.join(ROOT, "tests", "cpp", "lzma_alone.cpp")
    so_dir = os.path.dirname(lzb.SO_PATH)
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < max(os.path.getmtime(src), os.path.getmtime(lzb.SO_PATH)):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), src, "-o", BIN,
                               "-L", so_dir, "-l:liblzma_b200.so", "-Wl,-rpath," + so_dir])
    return BIN


def test_cpp_mirror_compiles_links_and_fails_loudly(lzb, tmp_path):
    exe = _build(lzb)
    if lzb.lib().lzb_device_count() > 0:
        pytest.skip("a GPU is present")
    src = tmp_path / "in.bin"
    src.write_bytes(b"hello hello hello")
    r = subprocess.run([exe, "e", str(src), str(tmp_path / "out.lzma")], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cpp_lzma_alone_roundtrip(lzb, oracle, corpus, tmp_path):
    exe = _build(lzb)
    data = corpus.generate(300000, 1, 0, 40, 0).tobytes()
    src, comp, back = tmp_path / "in.bin", tmp_path / "out.lzma", tmp_path / "back.bin"
    src.write_bytes(data)
    subprocess.check_call([exe, "e", str(src), str(comp), "20", "32"])
    assert comp.read_bytes() == oracle.encode(data, oracle.props(dict_size=1 << 20, fb=32), alone=True)
    subprocess.check_call([exe, "d", str(comp), str(back)])
    assert back.read_bytes() == data
    bad = bytearray(comp.read_bytes())
    bad[20] ^= 0xFF
    (tmp_path / "bad.lzma").write_bytes(bytes(bad))
    r = subprocess.run([exe, "d", str(tmp_path / "bad.lzma"), str(back)], capture_output=True, text=True)
    ok, _ = oracle.decode_alone(bytes(bad), out_cap=300273)
    assert (r.returncode == 0) == (ok == 1)
