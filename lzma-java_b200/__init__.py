"""lzma-java_b200 -- B200-native LZMA block codec behind the Encoder/Decoder
surface of rfalke/lzma-java.

This package is a thin ctypes host layer over the C-ABI shared library
``liblzma_b200.so`` (include/lzma_b200.h).  ``Encoder`` and ``Decoder`` mirror
the reference classes method for method (same names, argument meaning and
error behaviour; SevenZip/Compression/LZMA/Encoder.java:1064-1184,
Decoder.java:205-318) so the parity tests read like the reference's own.

There is no CPU fallback: if the CUDA library is missing or no sm_100 device
is present, construction raises.

Import with ``importlib.import_module("lzma-java_b200")`` (the directory name
is fixed by the repo layout and is not a Python identifier).
"""
import ctypes as C
import io
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("LZB_SO") or os.path.join(_HERE, "liblzma_b200.so")  # LZB_SO: A/B builds of the same ABI

LZB_OK = 1
LZB_FALSE = 0
LZB_E_CUDA = -1
LZB_E_ARG = -2
LZB_E_NOMEM = -3
LZB_E_CAPACITY = -4
LZB_E_UNSUPPORTED = -5
HEADER_SIZE = 13

PROGRESS_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_uint64, C.c_uint64)  # lzb_progress_fn
# every symbol include/lzma_b200.h declares: (name, restype, argtypes)
_vp, _u8p, _u64p, _i32p = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
ABI = [
    ("lzb_version", C.c_char_p, []),
    ("lzb_device_count", C.c_int, []),
    ("lzb_last_error", C.c_char_p, []),
    ("lzb_kernel_launches", C.c_uint64, []),
    ("lzb_host_alloc", C.c_void_p, [C.c_size_t]),
    ("lzb_host_free", None, [C.c_void_p]),
    ("lzb_enc_create", C.c_void_p, [C.c_int]),
    ("lzb_enc_destroy", None, [_vp]),
    ("lzb_enc_set_dictionary_size", C.c_int, [_vp, C.c_int32]),
    ("lzb_enc_set_num_fast_bytes", C.c_int, [_vp, C.c_int32]),
    ("lzb_enc_set_match_finder", C.c_int, [_vp, C.c_int32]),
    ("lzb_enc_set_lc_lp_pb", C.c_int, [_vp, C.c_int32, C.c_int32, C.c_int32]),
    ("lzb_enc_set_end_marker_mode", C.c_int, [_vp, C.c_int32]),
    ("lzb_enc_set_algorithm", C.c_int, [C.c_int32]),
    ("lzb_enc_write_coder_properties", C.c_int, [_vp, _u8p]),
    ("lzb_enc_bound", C.c_uint64, [C.c_uint64]),
    ("lzb_enc_set_progress", C.c_int, [_vp, _vp, _vp]),
    ("lzb_enc_code", C.c_int, [_vp, _u8p, C.c_uint64, _u8p, C.c_uint64, _u64p]),
    ("lzb_enc_code_batch", C.c_int, [_vp, _u8p, _u64p, _u64p, C.c_uint32, _u8p, _u64p, _u64p, _u64p, C.c_int32]),
    ("lzb_enc_code_batch_device", C.c_int,
     [_vp, _u8p, _u64p, _u64p, C.c_uint32, C.c_uint64, _u8p, _u64p, _u64p, _u64p, C.c_int32, _vp]),
    ("lzb_enc_trace_matches", C.c_int, [_vp, _u8p, C.c_uint64, _vp, _vp, C.c_uint64, _u64p]),
    ("lzb_dec_create", C.c_void_p, [C.c_int]),
    ("lzb_dec_destroy", None, [_vp]),
    ("lzb_dec_set_decoder_properties", C.c_int, [_vp, _u8p, C.c_uint32]),
    ("lzb_dec_code", C.c_int, [_vp, _u8p, C.c_uint64, _u8p, C.c_uint64, C.c_int64, _u64p]),
    ("lzb_dec_code_batch", C.c_int, [_vp, _u8p, _u64p, _u64p, C.c_uint32, _u8p, _u64p, _u64p, _u64p, _i32p]),
    ("lzb_dec_code_batch_device", C.c_int,
     [_vp, _u8p, _u64p, _u64p, C.c_uint32, _u8p, _u64p, _u64p, _u64p, _i32p, _vp]),
]

_lib = None


class LzbError(IOError):
    """Infrastructure failure (negative C-ABI return); the Java binding maps
    these to IOException."""

    def __init__(self, code, msg):
        super().__init__("lzma_b200 error %d: %s" % (code, msg))
        self.code = code


def lib():
    """Load liblzma_b200.so and declare every prototype.  Raises if the
    library has not been built -- there is no fallback implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError("%s not built: run `python __graft_entry__.py build` (no CPU fallback exists)" % SO_PATH)
        L = C.CDLL(SO_PATH)
        for name, restype, argtypes in ABI:
            fn = getattr(L, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = L
    return _lib


def last_error():
    return lib().lzb_last_error().decode()


def _check(rc):
    if rc < 0:
        raise LzbError(rc, last_error())
    return rc


def _u8(data):
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data, dtype=np.uint8)
    return np.frombuffer(bytes(data), dtype=np.uint8)


def _u64(x):
    return np.ascontiguousarray(x, dtype=np.uint64)


def enc_bound(n):
    return int(lib().lzb_enc_bound(int(n)))


def kernel_launches():
    return int(lib().lzb_kernel_launches())


class Encoder:
    """SevenZip.Compression.LZMA.Encoder (Encoder.java), setters return
    True/False exactly where the reference does."""

    def __init__(self, device=0):
        self._h = lib().lzb_enc_create(device)
        if not self._h:
            raise LzbError(LZB_E_CUDA, last_error())

    def close(self):
        if getattr(self, "_h", None):
            lib().lzb_enc_destroy(self._h)
            self._h = None

    __del__ = close

    # Encoder.java:1135 / 1148 / 1156 / 1169 / 1182 / 1127
    def SetDictionarySize(self, dictionarySize):
        return _check(lib().lzb_enc_set_dictionary_size(self._h, dictionarySize)) == 1

    def SetNumFastBytes(self, numFastBytes):
        return _check(lib().lzb_enc_set_num_fast_bytes(self._h, numFastBytes)) == 1

    def SetMatchFinder(self, matchFinderIndex):
        return _check(lib().lzb_enc_set_match_finder(self._h, matchFinderIndex)) == 1

    def SetLcLpPb(self, lc, lp, pb):
        return _check(lib().lzb_enc_set_lc_lp_pb(self._h, lc, lp, pb)) == 1

    def SetEndMarkerMode(self, endMarkerMode):
        _check(lib().lzb_enc_set_end_marker_mode(self._h, 1 if endMarkerMode else 0))

    @staticmethod
    def SetAlgorithm(algorithm):
        return lib().lzb_enc_set_algorithm(algorithm) == 1

    def WriteCoderProperties(self, outStream):  # Encoder.java:1079
        buf = (C.c_uint8 * 5)()
        _check(lib().lzb_enc_write_coder_properties(self._h, buf))
        outStream.write(bytes(buf))

    def Code(self, inStream, outStream, inSize=-1, outSize=-1, progress=None):
        """Encoder.Code (Encoder.java:1064): drains inStream, writes the payload
        (no header) to outStream.  inSize/outSize are ignored like in the
        reference.  progress.SetProgress(in, out) (ICodeProgress.java:3-5) is called with
        monotone running totals while the parser runs (about every 64 KiB; the reference: every
        >= 4096 bytes, Encoder.java:929-933) and once more with the final sizes."""
        data = inStream.read()
        if progress is None:
            payload = self.code_bytes(data)
        else:
            cb = PROGRESS_FN(lambda _user, n_in, n_out: progress.SetProgress(n_in, n_out))
            _check(lib().lzb_enc_set_progress(self._h, C.cast(cb, C.c_void_p), None))
            try:
                payload = self.code_bytes(data)
            finally:
                _check(lib().lzb_enc_set_progress(self._h, None, None))
        outStream.write(payload)
        if progress is not None:
            progress.SetProgress(len(data), len(payload))

    def code_bytes(self, data):
        a = _u8(data)
        cap = enc_bound(a.size)
        out = np.empty(cap, dtype=np.uint8)
        n = C.c_uint64(0)
        rc = _check(lib().lzb_enc_code(self._h, a.ctypes.data, a.size, out.ctypes.data, cap, C.byref(n)))
        if rc != 1:
            raise LzbError(rc, "encode failed")
        return out[: n.value].tobytes()

    def trace_matches(self, data):
        """Match-finder trace tap -> (counts[n], pairs[k, 2]) with pairs as (length, distance)."""
        a = _u8(data)
        counts = np.zeros(max(a.size, 1), dtype=np.uint32)
        cap = max(1024, 16 * a.size)
        pairs = np.zeros(2 * cap, dtype=np.uint32)
        used = C.c_uint64(0)
        rc = _check(lib().lzb_enc_trace_matches(self._h, a.ctypes.data, a.size, counts.ctypes.data, pairs.ctypes.data, cap,
                                                C.byref(used)))
        if rc != 1:
            raise LzbError(rc, "trace failed")
        return counts[: a.size], pairs[: 2 * used.value].reshape(-1, 2)

    def code_batch(self, in_arr, in_off, in_len, with_header=True):
        """n independent streams from host memory -> (out, out_off, out_len)."""
        in_arr, in_off, in_len = _u8(in_arr), _u64(in_off), _u64(in_len)
        n = in_off.size
        caps = np.array([enc_bound(int(x)) + (HEADER_SIZE if with_header else 0) for x in in_len], dtype=np.uint64)
        out_off = np.zeros(n, dtype=np.uint64)
        if n > 1:
            out_off[1:] = np.cumsum(caps)[:-1]
        out = np.empty(int(caps.sum()), dtype=np.uint8)
        out_len = np.zeros(n, dtype=np.uint64)
        rc = _check(lib().lzb_enc_code_batch(self._h, in_arr.ctypes.data, in_off.ctypes.data, in_len.ctypes.data, n,
                                             out.ctypes.data, out_off.ctypes.data, caps.ctypes.data,
                                             out_len.ctypes.data, 1 if with_header else 0))
        if rc != 1:
            raise LzbError(rc, "batch encode failed")
        return out, out_off, out_len

    def code_batch_device(self, d_in, d_in_off, d_in_len, n, max_in_len, d_out, d_out_off, d_out_cap, d_out_len,
                          with_header=True, stream=None):
        """Device-resident batch; arguments are device pointers (ints)."""
        _check(lib().lzb_enc_code_batch_device(self._h, d_in, d_in_off, d_in_len, n, max_in_len, d_out, d_out_off,
                                               d_out_cap, d_out_len, 1 if with_header else 0, stream))


class Decoder:
    """SevenZip.Compression.LZMA.Decoder (Decoder.java)."""

    def __init__(self, device=0):
        self._h = lib().lzb_dec_create(device)
        if not self._h:
            raise LzbError(LZB_E_CUDA, last_error())

    def close(self):
        if getattr(self, "_h", None):
            lib().lzb_dec_destroy(self._h)
            self._h = None

    __del__ = close

    def SetDecoderProperties(self, properties):  # Decoder.java:303
        p = bytes(properties)
        buf = (C.c_uint8 * max(len(p), 1))(*p)
        return _check(lib().lzb_dec_set_decoder_properties(self._h, buf, len(p))) == 1

    def Code(self, inStream, outStream, outSize):
        """Decoder.Code (Decoder.java:205): returns False on corrupt data."""
        ok, data = self.code_bytes(inStream.read(), outSize)
        if ok:
            outStream.write(data)  # the reference does not flush its window on failure (Decoder.java:281,290)
        return ok

    def code_bytes(self, payload, out_size, out_cap=None):
        """One payload -> (ok, bytes).  With out_size < 0 (decode until the end marker,
        Decoder.java:219,277-282) the output size is unknown: start from a guess and, like the Java
        binding, grow the buffer and decode again while the library reports LZB_E_CAPACITY."""
        a = _u8(payload)
        grow = out_cap is None and out_size < 0
        if out_cap is None:
            out_cap = (out_size if out_size >= 0 else 8 * a.size + 4096) + 273
        while True:
            out = np.empty(max(out_cap, 1), dtype=np.uint8)
            w = C.c_uint64(0)
            rc = lib().lzb_dec_code(self._h, a.ctypes.data, a.size, out.ctypes.data, out_cap, out_size, C.byref(w))
            if rc == LZB_E_CAPACITY and grow and out_cap < (1 << 32) - (1 << 20):
                out_cap = min(out_cap * 4, (1 << 32) - (1 << 20))
                continue
            _check(rc)
            return rc == 1, out[: w.value].tobytes()

    def code_batch(self, in_arr, in_off, in_len, out_off, out_cap):
        """n LzmaAlone streams from host memory -> (out, out_len, status)."""
        in_arr, in_off, in_len = _u8(in_arr), _u64(in_off), _u64(in_len)
        out_off, out_cap = _u64(out_off), _u64(out_cap)
        n = in_off.size
        total = int((out_off + out_cap).max()) if n else 0
        out = np.empty(max(total, 1), dtype=np.uint8)
        out_len = np.zeros(n, dtype=np.uint64)
        status = np.zeros(n, dtype=np.int32)
        _check(lib().lzb_dec_code_batch(self._h, in_arr.ctypes.data, in_off.ctypes.data, in_len.ctypes.data, n,
                                        out.ctypes.data, out_off.ctypes.data, out_cap.ctypes.data,
                                        out_len.ctypes.data, status.ctypes.data))
        return out, out_len, status

    def code_batch_device(self, d_in, d_in_off, d_in_len, n, d_out, d_out_off, d_out_cap, d_out_len, d_status,
                          stream=None):
        """Device-resident batch; arguments are device pointers (ints)."""
        _check(lib().lzb_dec_code_batch_device(self._h, d_in, d_in_off, d_in_len, n, d_out, d_out_off, d_out_cap,
                                               d_out_len, d_status, stream))


def decode_alone(stream_bytes, device=0):
    """LzmaAlone 'd' (LzmaAlone.java:220-239) for one .lzma file in memory."""
    s = bytes(stream_bytes)
    if len(s) < HEADER_SIZE:
        return False, b""
    dec = Decoder(device)
    try:
        if not dec.SetDecoderProperties(s[:5]):
            return False, b""
        size = int.from_bytes(s[5:13], "little", signed=True)
        return dec.code_bytes(s[13:], size)
    finally:
        dec.close()


def encode_alone(data, device=0, dict_size=1 << 23, lc=3, lp=0, pb=2, fb=128, mf=1, eos=False):
    """LzmaAlone 'e' (LzmaAlone.java:190-218) with the CLI's defaults."""
    enc = Encoder(device)
    try:
        ok = (enc.SetDictionarySize(dict_size) and enc.SetNumFastBytes(fb) and enc.SetMatchFinder(mf)
              and enc.SetLcLpPb(lc, lp, pb))
        if not ok:
            raise ValueError("Incorrect encoder parameter")
        enc.SetEndMarkerMode(eos)
        out = io.BytesIO()
        enc.WriteCoderProperties(out)
        size = -1 if eos else len(data)
        out.write(size.to_bytes(8, "little", signed=True))
        out.write(enc.code_bytes(data))
        return out.getvalue()
    finally:
        enc.close()
