"""GPU parity for the batch decoder (Decoder.java:205-301) through the C ABI:
streams produced by the CPU oracle (bit-identical to the reference encoder)
and by liblzma must decode to the original bytes and agree with the oracle's
decoder on return value and length, including corrupt and truncated input."""
import lzma
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["auto", "shared", "hybrid"])
def dec_mode(request):
    """Every test runs with the library's own choice of where the probability models live, with
    all of them forced into shared memory, and with the matched-literal tables forced into
    global memory (the mode big batches use; lzb_kernels.h DecMode)."""
    old = os.environ.pop("LZB_DEC_MODE", None)
    if request.param != "auto":
        os.environ["LZB_DEC_MODE"] = "0" if request.param == "shared" else "1"
    yield request.param
    os.environ.pop("LZB_DEC_MODE", None)
    if old is not None:
        os.environ["LZB_DEC_MODE"] = old

BASE = dict(dict_size=1 << 20, lc=3, lp=0, pb=2, fb=32, mf=1, eos=False)


def _pack(streams):
    off = np.zeros(len(streams), dtype=np.uint64)
    ln = np.array([len(s) for s in streams], dtype=np.uint64)
    if len(streams) > 1:
        off[1:] = np.cumsum(ln)[:-1]
    return np.frombuffer(b"".join(streams), dtype=np.uint8), off, ln


def _decode_all(lzb, streams, sizes, slack=273):
    arr, off, ln = _pack(streams)
    cap = np.array([s + slack for s in sizes], dtype=np.uint64)
    ooff = np.zeros(len(streams), dtype=np.uint64)
    if len(streams) > 1:
        ooff[1:] = np.cumsum(cap)[:-1]
    dec = lzb.Decoder()
    out, out_len, status = dec.code_batch(arr, off, ln, ooff, cap)
    dec.close()
    return [out[int(o): int(o) + int(l)].tobytes() for o, l in zip(ooff, out_len)], status


def test_decode_oracle_streams_all_classes(lzb, oracle, corpus):
    blocks, streams = [], []
    for cls in range(4):
        for k in range(6):
            size = [1, 2, 777, 4096, 65536, 262144][k]
            b = corpus.generate(size, 1, cls, 2, k).tobytes()
            blocks.append(b)
            streams.append(oracle.encode(b, oracle.props(**BASE), alone=True))
    got, status = _decode_all(lzb, streams, [len(b) for b in blocks])
    assert list(status) == [1] * len(blocks)
    for g, b in zip(got, blocks):
        assert g == b


@pytest.mark.parametrize("kw", [{"lc": 0}, {"lc": 8}, {"lp": 1}, {"lp": 4}, {"pb": 0}, {"pb": 4}, {"pb": 3},
                                {"lc": 4, "lp": 4, "pb": 4}, {"dict_size": 1}, {"dict_size": 4096},
                                {"fb": 273, "dict_size": 1 << 23}, {"mf": 0}])
def test_decode_property_variants(lzb, oracle, corpus, kw):
    p = dict(BASE)
    p.update(kw)
    blocks = [corpus.generate(50000 + 1000 * c, 1, c, 5, 3).tobytes() for c in range(4)]
    streams = [oracle.encode(b, oracle.props(**p), alone=True) for b in blocks]
    got, status = _decode_all(lzb, streams, [len(b) for b in blocks])
    assert list(status) == [1] * 4
    assert got == blocks


def test_decode_end_marker_and_unknown_size(lzb, oracle, corpus):
    p = dict(BASE)
    p["eos"] = True
    blocks = [corpus.generate(30000, 1, c, 6, 0).tobytes() for c in range(4)] + [b""]
    streams = [oracle.encode(b, oracle.props(**p), alone=True) for b in blocks]  # size field = -1
    got, status = _decode_all(lzb, streams, [len(b) for b in blocks])
    assert list(status) == [1] * 5 and got == blocks
    # liblzma's encoder: a different parser, same format
    filt = [{"id": lzma.FILTER_LZMA1, "dict_size": 1 << 20, "lc": 3, "lp": 0, "pb": 2, "nice_len": 64}]
    streams = [lzma.compress(b, format=lzma.FORMAT_ALONE, filters=filt) for b in blocks]
    got, status = _decode_all(lzb, streams, [len(b) for b in blocks])
    assert list(status) == [1] * 5 and got == blocks


def test_decode_mixed_properties_in_one_batch(lzb, oracle, corpus):
    variants = [{}, {"lc": 8}, {"pb": 4}, {"lp": 2, "lc": 2}, {"eos": True}, {"dict_size": 1 << 16}]
    blocks, streams = [], []
    for i, kw in enumerate(variants * 3):
        p = dict(BASE)
        p.update(kw)
        b = corpus.generate(20000 + 37 * i, 1, i % 4, 7, i).tobytes()
        blocks.append(b)
        streams.append(oracle.encode(b, oracle.props(**p), alone=True))
    got, status = _decode_all(lzb, streams, [len(b) for b in blocks])
    assert list(status) == [1] * len(blocks) and got == blocks


def test_decode_corrupt_and_truncated_matches_oracle(lzb, oracle, corpus):
    rng = np.random.default_rng(1234)
    good = [oracle.encode(corpus.generate(20000, 1, c, 8, 0).tobytes(), oracle.props(**BASE), alone=True) for c in range(4)]
    bad = []
    for s in good:
        a = bytearray(s)
        bad.append(bytes(a[: len(a) // 2]))            # truncated: EOF reads as all ones (RangeDecoder.java:23,36)
        for _ in range(3):                              # flipped payload bytes
            b = bytearray(s)
            i = int(rng.integers(13, len(b)))
            b[i] ^= 1 << int(rng.integers(0, 8))
            bad.append(bytes(b))
        b = bytearray(s)
        b[0] = 225                                      # pb = 5: SetDecoderProperties returns false
        bad.append(bytes(b))
        bad.append(bytes(s[:7]))                        # shorter than the header
    sizes = [20000] * len(bad)
    got, status = _decode_all(lzb, bad, sizes)
    for i, s in enumerate(bad):
        ok, ref = oracle.decode_alone(s, out_cap=20000 + 273)
        assert int(status[i]) == (ok if ok >= 0 else lzb.LZB_E_CAPACITY), i
        assert len(got[i]) == len(ref), i
        if ok == 1:
            assert got[i] == ref, i


def test_decode_capacity_error(lzb, oracle, corpus):
    b = corpus.generate(10000, 1, 0, 9, 0).tobytes()
    s = oracle.encode(b, oracle.props(**BASE), alone=True)
    arr, off, ln = _pack([s])
    dec = lzb.Decoder()
    out, out_len, status = dec.code_batch(arr, off, ln, np.zeros(1, dtype=np.uint64), np.array([5000], dtype=np.uint64))
    assert int(status[0]) == lzb.LZB_E_CAPACITY and int(out_len[0]) <= 5000
    dec.close()


def test_decoder_class_mirrors_reference(lzb, oracle, corpus):
    """LzmaAlone.java:220-239 written against the mirror classes."""
    import io
    b = corpus.generate(100000, 1, 0, 10, 0).tobytes()
    p = oracle.props(**BASE)
    s = oracle.encode(b, p, alone=True)
    dec = lzb.Decoder()
    assert not dec.SetDecoderProperties(s[:4])
    assert not dec.SetDecoderProperties(bytes([225, 0, 0, 16, 0]))
    assert dec.SetDecoderProperties(s[:5])
    out = io.BytesIO()
    assert dec.Code(io.BytesIO(s[13:]), out, int.from_bytes(s[5:13], "little"))
    assert out.getvalue() == b
    out = io.BytesIO()
    assert not dec.Code(io.BytesIO(b"\x00" + b"\xff" * 40), out, 1000)
    dec.close()


def test_decode_many_streams(lzb, oracle, corpus):
    """More streams than resident warp slots (148 SMs x 15, x 28 in hybrid mode): exercises the
    ticket queue, and the chunked host pipeline with kernels of several chunks sharing the SMs."""
    n = 5000
    data = corpus.generate(4096, n, corpus.MIXED, 11)
    off = np.arange(n, dtype=np.uint64) * 4096
    ln = np.full(n, 4096, dtype=np.uint64)
    comp, coff, clen = oracle.encode_batch(data, off, ln, oracle.props(**BASE), with_header=True, threads=8)
    dec = lzb.Decoder()
    cap = np.full(n, 4096 + 273, dtype=np.uint64)
    ooff = np.arange(n, dtype=np.uint64) * (4096 + 273)
    out, out_len, status = dec.code_batch(comp, coff, clen, ooff, cap)
    dec.close()
    assert (status == 1).all() and (out_len == 4096).all()
    got = out.reshape(n, 4096 + 273)[:, :4096].reshape(-1)
    assert np.array_equal(got, data)


@pytest.mark.parametrize("chunks", [None, "3"])
def test_decode_progressive_readback(lzb, oracle, corpus, chunks, monkeypatch):
    """Row-shaped outputs of >= 64 KiB take the progressive read-back of lzb_dec_code_batch (strided
    copies issued while the kernels run, driven by the kernel's progress counters).  Streams that
    end early (corrupt, truncated, bad header, capacity) must not stall it, and every byte the
    oracle's decoder writes must arrive.  One launch by default; LZB_DEC_CHUNKS=3 runs three chunks
    with a progress row each."""
    if chunks:
        monkeypatch.setenv("LZB_DEC_CHUNKS", chunks)
    n, size = 1300, 70000
    data = corpus.generate(size, n, corpus.MIXED, 21)
    off = np.arange(n, dtype=np.uint64) * size
    ln = np.full(n, size, dtype=np.uint64)
    comp, coff, clen = oracle.encode_batch(data, off, ln, oracle.props(**BASE), with_header=True, threads=8)
    comp = comp.copy()
    rng = np.random.default_rng(99)
    for i in rng.choice(n, 40, replace=False):      # flipped payload bytes
        comp[int(coff[i]) + 13 + int(rng.integers(0, int(clen[i]) - 13))] ^= 1 << int(rng.integers(0, 8))
    clen = clen.copy()
    for i in rng.choice(n, 20, replace=False):      # truncated
        clen[i] = clen[i] // 2
    comp[int(coff[7])] = 225                        # pb = 5
    clen[11] = 5                                    # shorter than the header
    pitch = size + 273 + 15
    cap = np.full(n, size + 273, dtype=np.uint64)
    ooff = np.arange(n, dtype=np.uint64) * pitch
    ref_out, ref_len, ref_status = oracle.decode_batch(comp, coff, clen, ooff, cap, threads=8)
    dec = lzb.Decoder()
    out, out_len, status = dec.code_batch(comp, coff, clen, ooff, cap)
    dec.close()
    ref_status = np.where(ref_status < 0, lzb.LZB_E_CAPACITY, ref_status)
    assert np.array_equal(status, ref_status)
    assert np.array_equal(out_len, ref_len)
    assert (status == 1).sum() > n - 80
    for i in range(n):
        a, b = int(ooff[i]), int(ooff[i]) + int(out_len[i])
        assert np.array_equal(out[a:b], ref_out[a:b]), i


def test_decode_alone_end_marker_grows_the_output(lzb, oracle):
    """An end-marker stream says nothing about its size; one that expands far beyond the first guess
    (here 100 000 zeros in ~40 bytes) must still decode: the mirror grows the buffer and decodes again
    on LZB_E_CAPACITY, as java/.../Decoder.java does (Decoder.java:219,277-282: outSize < 0)."""
    data = b"\0" * 100000 + b"tail"
    p = dict(BASE)
    p["eos"] = True
    s = oracle.encode(data, oracle.props(**p), alone=True)
    assert len(s) < 200
    ok, back = lzb.decode_alone(s)
    assert ok and back == data
    s2 = lzb.encode_alone(data, eos=True)
    ok, back = lzb.decode_alone(s2)
    assert ok and back == data


def test_decode_batch_leaves_gaps_between_outputs_alone(lzb, oracle, corpus):
    """include/lzma_b200.h promises writes only inside out[out_off[i] .. out_off[i] + out_cap[i]): host
    bytes between the regions (ragged capacities, so not the row-shaped fast path) keep their values."""
    n = 700
    sizes = [3000 + 17 * (i % 40) for i in range(n)]
    blocks = [corpus.generate(sz, 1, i % 4, 31, i).tobytes() for i, sz in enumerate(sizes)]
    streams = [oracle.encode(b, oracle.props(**BASE), alone=True) for b in blocks]
    arr, off, ln = _pack(streams)
    cap = np.array([sz + 273 for sz in sizes], dtype=np.uint64)
    gap = 96
    ooff = np.zeros(n, dtype=np.uint64)
    ooff[1:] = np.cumsum(cap + gap)[:-1]
    ooff += gap
    total = int(ooff[-1] + cap[-1]) + gap
    L = lzb.lib()
    out = np.full(total, 0xA5, dtype=np.uint8)
    out_len = np.zeros(n, dtype=np.uint64)
    status = np.zeros(n, dtype=np.int32)
    dec = lzb.Decoder()
    rc = L.lzb_dec_code_batch(dec._h, arr.ctypes.data, off.ctypes.data, ln.ctypes.data, n, out.ctypes.data, ooff.ctypes.data,
                              cap.ctypes.data, out_len.ctypes.data, status.ctypes.data)
    dec.close()
    assert rc == 1 and (status == 1).all()
    keep = np.ones(total, dtype=bool)
    for i in range(n):
        a = int(ooff[i])
        assert out[a: a + sizes[i]].tobytes() == blocks[i], i
        keep[a: a + int(cap[i])] = False
    assert (out[keep] == 0xA5).all()
