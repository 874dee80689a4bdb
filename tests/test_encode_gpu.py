"""GPU parity for the encoder pipeline (BinTree.java:152-356, Encoder.java:275-1125)
through the C ABI: compressed bytes must be bit-identical to the CPU oracle (which
reproduces the reference's 12 golden vectors) for every block and property set."""
import hashlib
import io
import json
import os

import numpy as np
import pytest

from conftest import CLI_DEFAULTS, VECTORS

pytestmark = pytest.mark.gpu

BASE = dict(dict_size=1 << 20, lc=3, lp=0, pb=2, fb=32, mf=1, eos=False)
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _encoder(lzb, p):
    enc = lzb.Encoder()
    assert enc.SetDictionarySize(p["dict_size"])
    assert enc.SetNumFastBytes(p["fb"])
    assert enc.SetMatchFinder(p["mf"])
    assert enc.SetLcLpPb(p["lc"], p["lp"], p["pb"])
    enc.SetEndMarkerMode(p["eos"])
    return enc


def _gpu_streams(lzb, p, blocks, with_header=True):
    off = np.zeros(len(blocks), dtype=np.uint64)
    ln = np.array([len(b) for b in blocks], dtype=np.uint64)
    if len(blocks) > 1:
        off[1:] = np.cumsum(ln)[:-1]
    arr = np.frombuffer(b"".join(blocks), dtype=np.uint8) if ln.sum() else np.zeros(1, dtype=np.uint8)
    enc = _encoder(lzb, p)
    out, ooff, olen = enc.code_batch(arr, off, ln, with_header=with_header)
    enc.close()
    return [out[int(o): int(o + l)].tobytes() for o, l in zip(ooff, olen)]


def _check(lzb, oracle, p, blocks):
    got = _gpu_streams(lzb, p, blocks)
    for i, b in enumerate(blocks):
        ref = oracle.encode(b, oracle.props(**p), alone=True)
        assert len(got[i]) == len(ref), (i, len(b), len(got[i]), len(ref))
        assert got[i] == ref, (i, len(b))


def test_encode_tiny_and_ragged_blocks(lzb, oracle, corpus):
    blocks = [b"", b"a", b"ab", b"abc", b"abcd", b"aaaaa", b"abcabcabcabc", b"\x00" * 300, bytes(range(256)) * 2]
    for cls in range(4):
        for size in (5, 63, 64, 65, 1000, 4097):
            blocks.append(corpus.generate(size, 1, cls, 20, size).tobytes())
    _check(lzb, oracle, BASE, blocks)


def test_encode_all_classes_64k(lzb, oracle, corpus):
    blocks = [corpus.generate(65536, 1, cls, 21, k).tobytes() for cls in range(4) for k in range(3)]
    _check(lzb, oracle, BASE, blocks)


@pytest.mark.parametrize("kw", [{"fb": 5}, {"fb": 64}, {"fb": 128, "dict_size": 1 << 23}, {"fb": 273, "dict_size": 1 << 23},
                                {"mf": 0}, {"mf": 2}, {"eos": True}, {"lc": 0}, {"lc": 8}, {"lp": 1}, {"lp": 4}, {"pb": 0},
                                {"pb": 4}, {"pb": 4, "fb": 273}, {"dict_size": 1}, {"dict_size": 4096},
                                {"dict_size": 100000, "fb": 48}, {"dict_size": 1 << 22}, {"lc": 4, "lp": 4, "pb": 3}])
def test_encode_property_variants(lzb, oracle, corpus, kw):
    p = dict(BASE)
    p.update(kw)
    blocks = [corpus.generate(40000 + 123 * c, 1, c, 22, 1).tobytes() for c in range(4)]
    _check(lzb, oracle, p, blocks)


def test_encode_degenerate_runs(lzb, oracle):
    blocks = [b"\x00" * 100000, b"ab" * 50000, b"abc" * 30000, (b"x" * 1000 + b"y") * 90,
              bytes([i % 251 for i in range(100000)])]
    _check(lzb, oracle, BASE, blocks)
    p = dict(BASE)
    p.update(fb=273, dict_size=1 << 16)
    _check(lzb, oracle, p, blocks)


def test_encode_committed_golden_fixtures(lzb, corpus):
    """tests/golden/oracle_vectors.json (made by the pinned oracle) against the GPU encoder."""
    vec = json.load(open(os.path.join(GOLDEN, "oracle_vectors.json")))
    for c in vec["cases"]:
        data = corpus.generate(c["size"], 1, c["cls"], c["config_id"], c["block"]).tobytes()
        s = _gpu_streams(lzb, c["props"], [data])[0]
        assert len(s) == c["out_len"] and hashlib.sha256(s).hexdigest() == c["out_sha256"], c


def test_encode_payload_only_and_class_api(lzb, oracle, corpus):
    """LzmaAlone.java:190-218 written against the mirror class."""
    data = corpus.generate(200000, 1, 0, 23, 0).tobytes()
    p = dict(BASE)
    p.update(dict_size=1 << 23, fb=128)
    enc = _encoder(lzb, p)
    assert not enc.SetNumFastBytes(4) and not enc.SetNumFastBytes(274)
    assert not enc.SetDictionarySize(0) and not enc.SetDictionarySize((1 << 29) + 1)
    assert not enc.SetLcLpPb(9, 0, 0) and not enc.SetLcLpPb(0, 5, 0) and not enc.SetLcLpPb(0, 0, 5)
    assert not enc.SetMatchFinder(3) and lzb.Encoder.SetAlgorithm(2)
    out = io.BytesIO()
    enc.WriteCoderProperties(out)
    assert out.getvalue() == oracle.props_bytes(oracle.props(**p))

    class Progress:
        calls = []

        def SetProgress(self, a, b):
            self.calls.append((a, b))

    pr = Progress()
    payload = io.BytesIO()
    enc.Code(io.BytesIO(data), payload, -1, -1, pr)
    enc.close()
    assert payload.getvalue() == oracle.encode(data, oracle.props(**p))
    assert pr.calls and pr.calls[-1][0] >= len(data)  # LzmaBench.java:219-223,366-368


def test_encode_reports_progress_while_it_runs(lzb, oracle, corpus):
    """ICodeProgress (ICodeProgress.java:3-5; Encoder.java:929-933, 1070-1072 call it after every CodeOneBlock):
    running totals arrive while the stream is being coded, monotone in both sizes, bounded by the final ones, and
    the bytes are still the oracle's."""
    data = corpus.generate(1 << 20, 1, corpus.TEXT, 41).tobytes()

    class Progress:
        def __init__(self):
            self.calls = []

        def SetProgress(self, a, b):
            self.calls.append((a, b))

    pr = Progress()
    enc = _encoder(lzb, BASE)
    payload = io.BytesIO()
    enc.Code(io.BytesIO(data), payload, -1, -1, pr)
    enc.close()
    ref = oracle.encode(data, oracle.props(**BASE))
    assert payload.getvalue() == ref
    assert pr.calls[-1] == (len(data), len(ref))
    running = pr.calls[:-1]
    assert len(running) >= 4, running  # 16 reports of 64 KiB leave the kernel; the host polls every 200 us
    assert all(0 < a <= len(data) for a, _ in running)
    assert all(x[0] < y[0] and x[1] <= y[1] for x, y in zip(running, running[1:]))
    assert all(0 < b <= len(ref) + 5 for _, b in running)  # getProcessedSizeAdd counts the coder's 5 pending bytes


def test_encode_1mib_blocks_c3(lzb, oracle, corpus):
    """BASELINE config 3 shape: 1 MiB blocks, dict 1 MiB, fb 64, mixed corpus (8 blocks here)."""
    p = dict(BASE)
    p.update(fb=64)
    n, size = 8, 1 << 20
    data = corpus.generate(size, n, corpus.MIXED, 3)
    off = np.arange(n, dtype=np.uint64) * size
    ln = np.full(n, size, dtype=np.uint64)
    ref, roff, rlen = oracle.encode_batch(data, off, ln, oracle.props(**p), True, 8)
    enc = _encoder(lzb, p)
    out, ooff, olen = enc.code_batch(data, off, ln, with_header=True)
    enc.close()
    assert np.array_equal(olen, rlen)
    for i in range(n):
        assert np.array_equal(out[int(ooff[i]): int(ooff[i] + olen[i])], ref[int(roff[i]): int(roff[i] + rlen[i])]), i


@pytest.mark.parametrize("switch,kw,length,md5", VECTORS, ids=[v[0] or "default" for v in VECTORS])
def test_encode_reference_golden_vectors(lzb, firefox, switch, kw, length, md5):
    """The reference's own parity contract, LzmaAloneTest.java:27-38, straight through the CUDA path:
    `LzmaAlone e firefox.exe <switch>` must produce exactly `length` bytes with this md5 (the Java
    test's constants), and the GPU decoder must give the file back.  The input is the reference's
    fixture (tests/golden/firefox.exe.xz on the GPU box, same md5)."""
    p = dict(CLI_DEFAULTS)
    p.update(kw)
    s = lzb.encode_alone(firefox, **p)
    assert len(s) == length
    assert hashlib.md5(s).hexdigest() == md5
    ok, back = lzb.decode_alone(s)
    assert ok and back == firefox


def test_roundtrip_gpu_encode_gpu_decode(lzb, corpus):
    n, size = 64, 100000
    data = corpus.generate(size, n, corpus.MIXED, 24)
    off = np.arange(n, dtype=np.uint64) * size
    ln = np.full(n, size, dtype=np.uint64)
    enc = _encoder(lzb, BASE)
    out, ooff, olen = enc.code_batch(data, off, ln, with_header=True)
    enc.close()
    dec = lzb.Decoder()
    cap = np.full(n, size + 273, dtype=np.uint64)
    doff = np.arange(n, dtype=np.uint64) * (size + 273)
    dout, dlen, status = dec.code_batch(out, ooff, olen, doff, cap)
    dec.close()
    assert (status == 1).all() and (dlen == size).all()
    assert np.array_equal(dout.reshape(n, size + 273)[:, :size].reshape(-1), data)


def test_encode_8mib_block_max_ratio_settings(lzb, oracle, corpus):
    """BASELINE config 4 shape: one 8 MiB block, dict 8 MiB, fb 273 (the largest block the encoder takes)."""
    p = dict(BASE)
    p.update(dict_size=1 << 23, fb=273)
    data = corpus.generate(1 << 23, 1, corpus.TEXT, 25, 0).tobytes()
    got = _gpu_streams(lzb, p, [data])[0]
    ref = oracle.encode(data, oracle.props(**p), alone=True)
    assert len(got) == len(ref) and got == ref
    ok, back = lzb.decode_alone(got)
    assert ok and back == data


def _stream_equals_oracle(lzb, oracle, p, data):
    enc = _encoder(lzb, p)
    got = enc.code_bytes(data)  # Encoder.Code: payload only
    enc.close()
    ref = oracle.encode(data, oracle.props(**p))
    assert len(got) == len(ref)
    assert got == ref
    dec = lzb.Decoder()
    assert dec.SetDecoderProperties(oracle.props_bytes(oracle.props(**p)))
    ok, back = dec.code_bytes(got, len(data))
    dec.close()
    assert ok and back == bytes(data)


def test_encode_stream_longer_than_8mib_and_than_its_dictionary(lzb, oracle, corpus):
    """Encoder.Code drains any stream to EOF with a sliding window (Encoder.java:1064-1077, InWindow.java:24-63,
    BinTree.java:84-90,164,231): 12 MiB through a 1 MiB dictionary, twelve window lengths, every class."""
    p = dict(BASE)
    data = np.concatenate([corpus.generate(3 << 20, 1, c, 41, c) for c in range(4)])
    _stream_equals_oracle(lzb, oracle, p, data)


def test_encode_lzmabench_buffer_at_d23(lzb, oracle, corpus):
    """LzmaBench's buffer is dictionary + 2 MiB of its own generator's data (LzmaBench.java:329-332): -d23 = 10 MiB."""
    p = dict(BASE)
    p.update(dict_size=1 << 23)
    data = corpus.generate((1 << 23) + (2 << 20), 1, corpus.REPETITIVE, 42, 0)
    _stream_equals_oracle(lzb, oracle, p, data)


def test_encode_distances_beyond_2_23(lzb, oracle, corpus):
    """Dictionaries above 8 MiB (SetDictionarySize takes up to 2^29, Encoder.java:1135-1146): the second
    part of the stream repeats data 9 MiB back, so the chosen matches have distances above 2^23."""
    p = dict(BASE)
    p.update(dict_size=1 << 25)
    a = np.concatenate([corpus.generate(8 << 20, 1, corpus.RANDOM, 43, 0), corpus.generate(1 << 20, 1, corpus.TEXT, 43, 1)])
    data = np.concatenate([a, a[1 << 20: 7 << 20], a[-(1 << 20):]])
    _stream_equals_oracle(lzb, oracle, p, data)
    enc = _encoder(lzb, p)
    counts, pairs = enc.trace_matches(data[(8 << 20):(8 << 20) + (2 << 20) + 4096])  # the tap speaks the wide format too
    enc.close()
    assert counts.sum() == len(pairs)


def test_stream_of_2_30_bytes_is_refused(lzb):
    """The one limit left: a single stream of 2^30 bytes or more (BinTree.Normalize, BinTree.java:358-375, is not built)."""
    import ctypes as C
    enc = lzb.Encoder()
    n = C.c_uint64(0)
    dummy = np.zeros(16, dtype=np.uint8)
    rc = lzb.lib().lzb_enc_code_batch_device(enc._h, dummy.ctypes.data, dummy.ctypes.data, dummy.ctypes.data, 1, 1 << 30,
                                             dummy.ctypes.data, dummy.ctypes.data, dummy.ctypes.data, dummy.ctypes.data, 0, None)
    assert rc == lzb.LZB_E_UNSUPPORTED
    enc.close()


def test_encode_many_tiny_blocks_and_pair_overflow_retry(lzb, oracle, corpus, monkeypatch):
    blocks = [corpus.generate(200 + (i % 97), 1, i % 4, 26, i).tobytes() for i in range(3000)]
    got = _gpu_streams(lzb, BASE, blocks)
    for i in range(0, 3000, 111):
        assert got[i] == oracle.encode(blocks[i], oracle.props(**BASE), alone=True), i
    # one pair slot per input byte is not enough for text: the wave is retried with more room
    monkeypatch.setenv("LZB_PAIR_MUL", "1")
    text = [corpus.generate(60000, 1, 0, 27, k).tobytes() for k in range(3)]
    _check(lzb, oracle, BASE, text)


def test_encode_more_blocks_than_parser_slots(lzb, oracle, corpus, monkeypatch):
    """A wave with more blocks than resident parser slots hands its blocks out longest-expected-first
    (lzb_encode.cu, run_waves); one parser slot per SM forces that on a small batch.  Placement of the
    outputs must not depend on the order: every block == oracle."""
    monkeypatch.setenv("LZB_ENC_WARPS", "1")
    blocks = [corpus.generate(500 + 37 * (i % 50), 1, i % 4, 29, i).tobytes() for i in range(400)] + [b"", b"z"]
    got = _gpu_streams(lzb, BASE, blocks)
    for i, b in enumerate(blocks):
        assert got[i] == oracle.encode(b, oracle.props(**BASE), alone=True), i
    monkeypatch.setenv("LZB_ENC_FIFO", "1")  # the plain ticket order gives the same bytes
    assert _gpu_streams(lzb, BASE, blocks) == got


@pytest.mark.parametrize("kw", [{}, {"fb": 64}, {"fb": 273, "dict_size": 1 << 23}, {"mf": 0}, {"dict_size": 4096}, {"dict_size": 1}])
def test_match_finder_lists_equal_the_instrumented_trace(lzb, oracle, corpus, kw):
    """North-star subsystem 1: the per-position (length, distance) lists of the parallel bt4/bt2
    match finder are identical to the oracle's trace of BinTree.fillMatches0 / Skip for EVERY
    position (BinTree.java:139-356), not just where the parser happened to look."""
    p = dict(BASE)
    p.update(kw)
    for cls in range(4):
        data = corpus.generate(30000 + 1111 * cls, 1, cls, 28, cls).tobytes()
        _, tr = oracle.encode(data, oracle.props(**p), trace=True)
        ref_counts = np.diff(tr["mf_off"].astype(np.int64))
        enc = _encoder(lzb, p)
        counts, pairs = enc.trace_matches(data)
        enc.close()
        assert np.array_equal(counts.astype(np.int64), ref_counts), (kw, cls)
        # the parser truncates a list in place near the end of a chunk (Encoder.java:737-743), which the
        # oracle's tap records before; compare the untouched (length, distance) pairs
        assert np.array_equal(pairs.astype(np.int64), tr["mf_pairs"].astype(np.int64)), (kw, cls)


def test_reference_learning_test_inputs(lzb, oracle):
    """The four inputs of LZMA/EncoderLearningTest.java:35-38 with the class defaults (Encoder.java:26-27,151-158):
    the reference prints a trace and asserts nothing; here GPU == oracle and the round trip holds."""
    inputs = [bytes([99, 100, 98, 100, 100, 100, 100, 100, 100, 100, 100]),
              bytes([100, 101, 102, 103, 104, 105, 101, 102, 101, 102]),
              bytes([100, 101, 102, 103, 101, 104, 101, 101, 101]),
              bytes([100, 100, 100, 101, 100, 100, 100, 101, 100, 100, 100, 101])]
    enc = lzb.Encoder()  # class defaults: dict 1 << 22, fb 32, lc3 lp0 pb2, bt4
    p = oracle.props()
    for data in inputs:
        payload = enc.code_bytes(data)
        assert payload == oracle.encode(data, p)
        ok, back = oracle.decode(oracle.props_bytes(p), payload, len(data))
        assert ok == 1 and back == data
    enc.close()


def _batch_equals_oracle(lzb, oracle, p, data, n, size, threads=8):
    off = np.arange(n, dtype=np.uint64) * size
    ln = np.full(n, size, dtype=np.uint64)
    ref, roff, rlen = oracle.encode_batch(data, off, ln, oracle.props(**p), True, threads)
    enc = _encoder(lzb, p)
    out, ooff, olen = enc.code_batch(data, off, ln, with_header=True)
    enc.close()
    assert np.array_equal(olen, rlen)
    for i in range(n):
        assert np.array_equal(out[int(ooff[i]): int(ooff[i] + olen[i])], ref[int(roff[i]): int(roff[i] + rlen[i])]), i
    dec = lzb.Decoder()
    cap = np.full(n, size + 273, dtype=np.uint64)
    doff = np.arange(n, dtype=np.uint64) * (size + 273)
    dout, dlen, status = dec.code_batch(out, ooff, olen, doff, cap)
    dec.close()
    assert (status == 1).all() and (dlen == size).all()
    assert np.array_equal(dout.reshape(n, size + 273)[:, :size].reshape(-1), data)


def test_encode_4mib_blocks_c5(lzb, oracle, corpus):
    """BASELINE config 5 shape: 4 MiB blocks, dict 4 MiB, fb 32, the mixed class cycle (8 blocks = two of
    each class): every block == oracle, and the GPU decoder gives the corpus back."""
    p = dict(BASE)
    p.update(dict_size=1 << 22, fb=32)
    n, size = 8, 1 << 22
    _batch_equals_oracle(lzb, oracle, p, corpus.generate(size, n, corpus.MIXED, 5), n, size)


def test_encode_8mib_blocks_c4_all_classes(lzb, oracle, corpus):
    """BASELINE config 4 shape on every corpus class: 8 MiB blocks, dict 8 MiB, fb 273 (max-ratio settings)."""
    p = dict(BASE)
    p.update(dict_size=1 << 23, fb=273)
    n, size = 4, 1 << 23
    _batch_equals_oracle(lzb, oracle, p, corpus.generate(size, n, corpus.MIXED, 4), n, size, threads=4)


def test_encode_groups_and_waves(lzb, oracle, corpus, monkeypatch):
    """The match finder works through a batch in groups (its scratch is reused) and the parser in waves (as many
    groups as fit the list pool).  Tiny limits force 9 groups and several waves on a small batch; every block ==
    oracle whatever the cut (lzb_encode.cu, run_encode)."""
    monkeypatch.setenv("LZB_ENC_GROUP", "5")
    monkeypatch.setenv("LZB_ENC_POOL_MB", "3")
    blocks = [corpus.generate(60000 + 7001 * (i % 9), 1, i % 4, 45, i).tobytes() for i in range(43)] + [b"", b"q"]
    got = _gpu_streams(lzb, BASE, blocks)
    for i, b in enumerate(blocks):
        assert got[i] == oracle.encode(b, oracle.props(**BASE), alone=True), i


def test_encode_device_outputs_stay_inside_their_capacity(lzb, oracle, corpus):
    """lzb_enc_code_batch_device writes block i only inside d_out[out_off[i] .. out_off[i] + out_cap[i]): canaries between the
    regions survive, a block whose capacity is too small reports UINT64_MAX and leaves its neighbours alone, and a block longer
    than the declared max_in_len is refused (LZB_E_ARG) instead of overrunning the match finder's scratch."""
    import torch
    dev = torch.device("cuda:0")
    n, size, gap = 24, 30000, 512
    data = corpus.generate(size, n, corpus.MIXED, 46)
    cap = lzb.enc_bound(size) + 13
    caps = np.full(n, cap, dtype=np.int64)
    caps[5] = 100          # far too small
    pitch = cap + gap
    d_in = torch.from_numpy(data).to(dev)
    off = torch.arange(n, dtype=torch.int64, device=dev) * size
    ln = torch.full((n,), size, dtype=torch.int64, device=dev)
    ooff = torch.arange(n, dtype=torch.int64, device=dev) * pitch + gap
    ocap = torch.from_numpy(caps).to(dev)
    d_out = torch.full((n * pitch + gap,), 0xA5, dtype=torch.uint8, device=dev)
    d_len = torch.zeros(n, dtype=torch.int64, device=dev)
    enc = _encoder(lzb, BASE)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        enc.code_batch_device(d_in.data_ptr(), off.data_ptr(), ln.data_ptr(), n, size, d_out.data_ptr(), ooff.data_ptr(),
                              ocap.data_ptr(), d_len.data_ptr(), True, st.cuda_stream)
    torch.cuda.synchronize()
    out = d_out.cpu().numpy()
    lens = d_len.cpu().numpy()
    keep = np.ones(out.size, dtype=bool)
    for i in range(n):
        a = gap + i * pitch
        keep[a: a + int(caps[i])] = False
        if i == 5:
            assert lens[i] == -1  # UINT64_MAX
            continue
        ref = oracle.encode(data[i * size:(i + 1) * size], oracle.props(**BASE), alone=True)
        assert out[a: a + int(lens[i])].tobytes() == ref, i
    assert (out[keep] == 0xA5).all()
    # a block longer than max_in_len
    with pytest.raises(lzb.LzbError) as ei:
        with torch.cuda.stream(st):
            enc.code_batch_device(d_in.data_ptr(), off.data_ptr(), ln.data_ptr(), n, size - 1, d_out.data_ptr(), ooff.data_ptr(),
                                  ocap.data_ptr(), d_len.data_ptr(), True, st.cuda_stream)
    assert ei.value.code == lzb.LZB_E_ARG
    enc.close()
