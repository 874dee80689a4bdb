"""Pins the CPU oracle against everything the reference's own tests hold for
the path: the 12 firefox.exe (length, md5) vectors of
LzmaAloneTest.java:27-38, the range-encoder byte strings of
RangeCoder/EncoderLearningTest.java:31-72, the bit-tree prices of
RangeCoder/BitTreeEncoderLearningTest.java:24-31 and ProbPrices spot values."""
import ctypes as C
import hashlib
import json
import lzma
import os

import numpy as np
import pytest

from conftest import CLI_DEFAULTS, FIREFOX, FIREFOX_XZ, VECTORS  # noqa: F401

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _rc_bits(O, bits):
    arr = (C.c_int * max(len(bits), 1))(*bits)
    out = (C.c_uint8 * 64)()
    n = O.lib().lzo_kat_rc_bits(arr, len(bits), out, 64)
    return " ".join("%02x" % out[i] for i in range(n))


def test_range_encoder_kats(oracle):  # EncoderLearningTest.java:29-49
    assert _rc_bits(oracle, [0, 0, 0]) == "00 00 00 00 00"
    assert _rc_bits(oracle, [1, 1, 1]) == "00 dc f8 3c 00"
    assert _rc_bits(oracle, []) == "00 00 00 00 00"
    assert _rc_bits(oracle, [0]) == "00 00 00 00 00"
    assert _rc_bits(oracle, [1]) == "00 7f ff fc 00"
    assert _rc_bits(oracle, [0, 1] * 5) == "00 56 fa d6 38 2c"
    assert _rc_bits(oracle, [1] * 10) == "00 ff 2e 08 28 00"
    assert _rc_bits(oracle, [0, 1] * 10) == "00 57 0d 5d 83 4f 8e"
    assert _rc_bits(oracle, [1] * 20) == "00 ff fb 88 c9 99"


def test_range_encoder_direct_bits(oracle):  # EncoderLearningTest.java:55-68
    def run(calls):
        v = (C.c_int * len(calls))(*[c[0] for c in calls])
        nb = (C.c_int * len(calls))(*[c[1] for c in calls])
        out = (C.c_uint8 * 64)()
        n = oracle.lib().lzo_kat_rc_direct(v, nb, len(calls), out, 64)
        return " ".join("%02x" % out[i] for i in range(n))
    assert run([(0x1, 2), (0xD, 4)]) == "00 73 ff ff fc"
    assert run([(0x1D, 6)]) == "00 73 ff ff fc"


def test_bittree_prices(oracle):  # BitTreeEncoderLearningTest.java:14-31
    pr = (C.c_int * 8)()
    oracle.lib().lzo_kat_bittree_prices(pr)
    assert list(pr) == [194, 194, 192, 186, 196, 196, 196, 196]


def test_prob_prices_table(oracle):  # ProbPrices.java:8-18, SURVEY App. A #13
    t = (C.c_int * 512)()
    oracle.lib().lzo_kat_prob_prices(t)
    assert (t[0], t[1], t[2], t[3], t[4], t[128], t[256], t[511]) == (0, 576, 512, 480, 448, 128, 64, 0)


def test_packed_fixture_is_the_reference_file():
    """tests/golden/firefox.exe.xz unpacks to the reference's own test input (checked wherever both exist)."""
    packed = lzma.decompress(open(FIREFOX_XZ, "rb").read())
    assert len(packed) == 916960 and hashlib.md5(packed).hexdigest() == "5744fff8e72d105c138dae9e17bb29fe"
    if os.path.exists(FIREFOX):
        assert packed == open(FIREFOX, "rb").read()


@pytest.mark.parametrize("switch,kw,length,md5", VECTORS, ids=[v[0] or "default" for v in VECTORS])
def test_firefox_golden_vectors(oracle, firefox, switch, kw, length, md5):
    data = firefox
    d = dict(CLI_DEFAULTS)
    d.update(kw)
    s = oracle.encode(data, oracle.props(**d), alone=True)
    assert len(s) == length
    assert hashlib.md5(s).hexdigest() == md5
    ok, back = oracle.decode_alone(s)
    assert ok == 1 and back == data
    if d["lc"] + d["lp"] <= 4:  # liblzma's own limit; independent check of the container format
        assert lzma.decompress(s, format=lzma.FORMAT_ALONE) == data


def test_committed_golden_fixtures(oracle, corpus):
    """tests/golden/oracle_vectors.json was produced by tests/golden/make_golden.py with
    the oracle that passed the 12 vectors above; it travels to the GPU box, where the
    reference fixture does not exist."""
    vec = json.load(open(os.path.join(GOLDEN, "oracle_vectors.json")))
    assert len(vec["cases"]) >= 20
    for c in vec["cases"]:
        data = corpus.generate(c["size"], 1, c["cls"], c["config_id"], c["block"])
        assert hashlib.sha256(data.tobytes()).hexdigest() == c["in_sha256"], c
        s = oracle.encode(data, oracle.props(**c["props"]), alone=True)
        assert len(s) == c["out_len"] and hashlib.sha256(s).hexdigest() == c["out_sha256"], c


def test_oracle_decodes_liblzma_streams(oracle, corpus):
    for cls in range(4):
        data = corpus.generate(50000, 1, cls, 9).tobytes()
        filt = [{"id": lzma.FILTER_LZMA1, "dict_size": 1 << 20, "lc": 3, "lp": 0, "pb": 2, "nice_len": 64}]
        s = lzma.compress(data, format=lzma.FORMAT_ALONE, filters=filt)
        ok, back = oracle.decode_alone(s)  # liblzma writes size -1 + end marker
        assert ok == 1 and back == data


def test_oracle_edge_cases(oracle):
    p = oracle.props(dict_size=1 << 16, fb=32)
    assert oracle.encode(b"", p) == b"\x00" * 5  # App. A #16
    for data in [b"a", b"ab", b"abc", b"aaaa", b"\x00" * 1000, bytes(range(256)) * 3]:
        pay = oracle.encode(data, p)
        ok, back = oracle.decode(oracle.props_bytes(p), pay, len(data))
        assert ok == 1 and back == data
    pe = oracle.props(dict_size=1 << 16, fb=32, eos=True)
    pay = oracle.encode(b"hello hello hello", pe)
    ok, back = oracle.decode(oracle.props_bytes(pe), pay, -1)
    assert ok == 1 and back == b"hello hello hello"
    # corrupt / truncated input: reference returns false or garbage, never throws
    pay = oracle.encode(b"hello hello hello hello", p)
    ok, _ = oracle.decode(oracle.props_bytes(p), b"\x00\xff\xff\xff\xff\xff\xff", 100)
    assert ok in (0, 1)
    assert not oracle.lib().lzo_props_valid(C.byref(oracle.props(fb=4)))
    assert not oracle.lib().lzo_props_valid(C.byref(oracle.props(dict_size=0)))
    assert not oracle.lib().lzo_props_valid(C.byref(oracle.props(lc=9)))
    assert oracle.lib().lzo_props_valid(C.byref(oracle.props(dict_size=1 << 29, fb=273, lc=8, lp=4, pb=4)))
