"""Synthetic corpora (tools/corpus.c) for tests and bench.py -- SURVEY.md 8(d)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcorpus.so")
TEXT, BINARY, RANDOM, REPETITIVE, MIXED = 0, 1, 2, 3, 4
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "corpus.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-pthread", "-o", _SO, src, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.corpus_generate.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_int, C.c_uint64, C.c_int]
        _lib.corpus_generate.restype = None
    return _lib


def generate(block_size, n_blocks, cls=TEXT, config_id=0, first_block=0, threads=None, out=None):
    """n_blocks blocks of block_size bytes, concatenated, as a uint8 array."""
    if threads is None:
        threads = min(os.cpu_count() or 1, 64)
    total = block_size * n_blocks
    if out is None:
        out = np.empty(total, dtype=np.uint8)
    assert out.dtype == np.uint8 and out.size >= total and out.flags["C_CONTIGUOUS"]
    if total:
        _load().corpus_generate(out.ctypes.data, block_size, n_blocks, first_block, cls, config_id, threads)
    return out
