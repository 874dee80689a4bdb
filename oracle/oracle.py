"""ctypes binding of the CPU parity oracle (oracle/lzma_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(lzma-java_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liblzma_oracle.so")


class Props(C.Structure):
    """Mirror of lzo_props == the reference's setters (Encoder.java:1127-1184)."""
    _fields_ = [("dict_size", C.c_int32), ("lc", C.c_int32), ("lp", C.c_int32),
                ("pb", C.c_int32), ("fb", C.c_int32), ("mf", C.c_int32),
                ("eos", C.c_int32)]


class Trace(C.Structure):
    _fields_ = [("mf_off", C.POINTER(C.c_uint32)), ("mf_pairs", C.POINTER(C.c_uint32)),
                ("mf_pairs_cap", C.c_uint64), ("mf_pairs_used", C.c_uint64),
                ("mf_overflow", C.c_uint64), ("dec", C.POINTER(C.c_int64)),
                ("dec_cap", C.c_uint64), ("dec_used", C.c_uint64)]


def build(force=False):
    src = os.path.join(_HERE, "lzma_oracle.c")
    hdr = os.path.join(_HERE, "lzma_oracle.h")
    if (force or not os.path.exists(_SO)
            or os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr))):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p = C.POINTER(C.c_uint8)
        L.lzo_props_valid.argtypes = [C.POINTER(Props)]
        L.lzo_props_valid.restype = C.c_int
        L.lzo_write_props.argtypes = [C.POINTER(Props), u8p]
        L.lzo_encode_bound.argtypes = [C.c_size_t]
        L.lzo_encode_bound.restype = C.c_size_t
        L.lzo_encode.argtypes = [C.POINTER(Props), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(Trace)]
        L.lzo_encode.restype = C.c_size_t
        L.lzo_encode_alone.argtypes = [C.POINTER(Props), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.lzo_encode_alone.restype = C.c_size_t
        L.lzo_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int64,
                                 C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.lzo_decode.restype = C.c_int
        L.lzo_decode_alone.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.lzo_decode_alone.restype = C.c_int
        L.lzo_kat_rc_bits.argtypes = [C.POINTER(C.c_int), C.c_int, u8p, C.c_size_t]
        L.lzo_kat_rc_bits.restype = C.c_size_t
        L.lzo_kat_rc_direct.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, u8p, C.c_size_t]
        L.lzo_kat_rc_direct.restype = C.c_size_t
        L.lzo_kat_bittree_prices.argtypes = [C.POINTER(C.c_int)]
        L.lzo_kat_prob_prices.argtypes = [C.POINTER(C.c_int)]
        L.lzo_encode_batch.argtypes = [C.POINTER(Props), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.lzo_encode_batch.restype = C.c_int
        L.lzo_decode_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.lzo_decode_batch.restype = C.c_int
        _lib = L
    return _lib


def props(dict_size=1 << 22, lc=3, lp=0, pb=2, fb=32, mf=1, eos=False):
    """Class defaults of the reference Encoder (Encoder.java:26-27,151-158,172)."""
    return Props(dict_size, lc, lp, pb, fb, mf, 1 if eos else 0)


def _buf(data):
    a = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, dtype=np.uint8)
    return a


def props_bytes(p):
    out = (C.c_uint8 * 5)()
    lib().lzo_write_props(C.byref(p), out)
    return bytes(out)


def encode(data, p, alone=False, trace=False):
    """Encoder.Code payload (or the LzmaAlone file when alone=True)."""
    a = _buf(data)
    n = a.size
    cap = lib().lzo_encode_bound(n) + 64
    out = np.empty(cap, dtype=np.uint8)
    if alone:
        r = lib().lzo_encode_alone(C.byref(p), a.ctypes.data, n, out.ctypes.data, cap)
        if r == C.c_size_t(-1).value:
            raise ValueError("oracle encode failed")
        return out[:r].tobytes()
    tr = None
    if trace:
        tr = Trace()
        mf_off = np.zeros(n + 1, dtype=np.uint32)
        cap_pairs = max(1024, 24 * n)
        mf_pairs = np.zeros(2 * cap_pairs, dtype=np.uint32)
        dec = np.zeros(3 * (n + 1), dtype=np.int64)
        tr.mf_off = mf_off.ctypes.data_as(C.POINTER(C.c_uint32))
        tr.mf_pairs = mf_pairs.ctypes.data_as(C.POINTER(C.c_uint32))
        tr.mf_pairs_cap = cap_pairs
        tr.dec = dec.ctypes.data_as(C.POINTER(C.c_int64))
        tr.dec_cap = n + 1
    r = lib().lzo_encode(C.byref(p), a.ctypes.data, n, out.ctypes.data, cap, C.byref(tr) if tr else None)
    if r == C.c_size_t(-1).value:
        raise ValueError("oracle encode failed")
    payload = out[:r].tobytes()
    if not trace:
        return payload
    if tr.mf_overflow:
        raise ValueError("oracle trace overflow")
    return payload, {
        "mf_off": mf_off,
        "mf_pairs": mf_pairs[: 2 * tr.mf_pairs_used].reshape(-1, 2),
        "decisions": dec[: 3 * tr.dec_used].reshape(-1, 3),
    }


def decode(props5, payload, out_size, slack=273):
    """Decoder.SetDecoderProperties + Decoder.Code -> (ok, bytes)."""
    a = _buf(payload)
    cap = (out_size if out_size >= 0 else 64 * max(a.size, 1) + 4096) + slack
    out = np.empty(max(cap, 1), dtype=np.uint8)
    w = C.c_size_t(0)
    pb = (C.c_uint8 * 5)(*props5)
    r = lib().lzo_decode(pb, a.ctypes.data, a.size, out.ctypes.data, cap, out_size, C.byref(w), None)
    return r, out[: w.value].tobytes()


def decode_alone(stream, out_cap=None):
    a = _buf(stream)
    if out_cap is None:
        size = int.from_bytes(bytes(a[5:13]), "little") if a.size >= 13 else 0
        out_cap = (size if size < (1 << 62) else 64 * a.size + 4096) + 273
    out = np.empty(max(out_cap, 1), dtype=np.uint8)
    w = C.c_size_t(0)
    r = lib().lzo_decode_alone(a.ctypes.data, a.size, out.ctypes.data, out_cap, C.byref(w))
    return r, out[: w.value].tobytes()


def encode_batch(in_arr, in_off, in_len, p, with_header=True, threads=1):
    """cpu_baseline leg: one block per task over a pthread pool.  Returns (out, out_off, out_len)."""
    in_arr = np.ascontiguousarray(in_arr, dtype=np.uint8)
    in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
    in_len = np.ascontiguousarray(in_len, dtype=np.uint64)
    n = in_off.size
    caps = np.array([lib().lzo_encode_bound(int(x)) + 13 + 64 for x in in_len], dtype=np.uint64)
    out_off = np.zeros(n, dtype=np.uint64)
    if n > 1:
        out_off[1:] = np.cumsum(caps)[:-1]
    out = np.empty(int(caps.sum()), dtype=np.uint8)
    out_len = np.zeros(n, dtype=np.uint64)
    rc = lib().lzo_encode_batch(C.byref(p), in_arr.ctypes.data, in_off.ctypes.data, in_len.ctypes.data, n,
                                out.ctypes.data, out_off.ctypes.data, caps.ctypes.data, out_len.ctypes.data,
                                1 if with_header else 0, threads)
    if rc != 0:
        raise ValueError("oracle batch encode failed")
    return out, out_off, out_len


def decode_batch(in_arr, in_off, in_len, out_off, out_cap, threads=1):
    """cpu_baseline leg.  Returns (out, out_len, status)."""
    in_arr = np.ascontiguousarray(in_arr, dtype=np.uint8)
    in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
    in_len = np.ascontiguousarray(in_len, dtype=np.uint64)
    out_off = np.ascontiguousarray(out_off, dtype=np.uint64)
    out_cap = np.ascontiguousarray(out_cap, dtype=np.uint64)
    n = in_off.size
    total = int((out_off + out_cap).max()) if n else 0
    out = np.empty(max(total, 1), dtype=np.uint8)
    out_len = np.zeros(n, dtype=np.uint64)
    status = np.zeros(n, dtype=np.int32)
    lib().lzo_decode_batch(in_arr.ctypes.data, in_off.ctypes.data, in_len.ctypes.data, n, out.ctypes.data,
                           out_off.ctypes.data, out_cap.ctypes.data, out_len.ctypes.data, status.ctypes.data, threads)
    return out, out_len, status
