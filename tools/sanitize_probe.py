"""Small encode + decode round trip for compute-sanitizer (memcheck / racecheck) runs."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lzb = importlib.import_module("lzma-java_b200")
from tools import corpus  # noqa: E402


def main():
    sizes = [0, 1, 5, 300, 4097, 20000, 20000, 20000, 20000]
    blocks = [corpus.generate(max(s, 1), 1, i % 4, 50, i)[:s] for i, s in enumerate(sizes)]
    data = np.concatenate(blocks) if sum(sizes) else np.zeros(1, dtype=np.uint8)
    ln = np.array(sizes, dtype=np.uint64)
    off = np.zeros(len(sizes), dtype=np.uint64)
    off[1:] = np.cumsum(ln)[:-1]
    for kw in [dict(dict_size=1 << 16, fb=32, lc=3, lp=0, pb=2), dict(dict_size=1 << 12, fb=273, lc=4, lp=2, pb=4)]:
        enc = lzb.Encoder()
        assert enc.SetDictionarySize(kw["dict_size"]) and enc.SetNumFastBytes(kw["fb"]) and enc.SetLcLpPb(kw["lc"], kw["lp"], kw["pb"])
        out, ooff, olen = enc.code_batch(data, off, ln, with_header=True)
        enc.close()
        dec = lzb.Decoder()
        cap = ln + np.uint64(273)
        doff = np.zeros(len(sizes), dtype=np.uint64)
        doff[1:] = np.cumsum(cap)[:-1]
        dout, dlen, status = dec.code_batch(out, ooff, olen, doff, cap)
        dec.close()
        assert (status == 1).all() and np.array_equal(dlen, ln), (status, dlen)
        for i, s in enumerate(sizes):
            assert np.array_equal(dout[int(doff[i]): int(doff[i]) + s], blocks[i]), i
    print("sanitize probe OK")


if __name__ == "__main__":
    main()
