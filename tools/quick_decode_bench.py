"""Throwaway-style developer probe: decode throughput of N oracle-encoded text streams with
inputs resident in HBM.  (bench.py is the contract benchmark; this exists for ncu captures.)"""
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lzb = importlib.import_module("lzma-java_b200")
from oracle import oracle as O  # noqa: E402
from tools import corpus  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
    cls = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    data = corpus.generate(size, n, cls, 2)
    off = np.arange(n, dtype=np.uint64) * size
    ln = np.full(n, size, dtype=np.uint64)
    t = time.time()
    comp, coff, clen = O.encode_batch(data, off, ln, O.props(dict_size=1 << 20, fb=32), True, os.cpu_count())
    print("oracle encode %.1fs, ratio %.3f" % (time.time() - t, clen.sum() / data.size))
    # repack compressed streams contiguously
    coff2 = np.zeros(n, dtype=np.uint64)
    coff2[1:] = np.cumsum(clen)[:-1]
    packed = np.concatenate([comp[int(o): int(o + l)] for o, l in zip(coff, clen)])
    dev = torch.device("cuda:0")
    d_in = torch.from_numpy(packed).to(dev)
    meta = torch.from_numpy(np.concatenate([coff2, clen, off + np.arange(n, dtype=np.uint64) * 0, ln + 0]).astype(np.int64)).to(dev)
    cap = size + 288
    ooff = torch.arange(n, dtype=torch.int64, device=dev) * cap
    ocap = torch.full((n,), cap, dtype=torch.int64, device=dev)
    d_out = torch.empty(n * cap, dtype=torch.uint8, device=dev)
    d_len = torch.zeros(n, dtype=torch.int64, device=dev)
    d_status = torch.zeros(n, dtype=torch.int32, device=dev)
    dec = lzb.Decoder()
    side = torch.cuda.Stream()  # a non-default stream: handle 0 would mean "the decoder's own stream"
    st = side.cuda_stream
    torch.cuda.synchronize()
    best = None
    for it in range(iters + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(side):
            e0.record()
            dec.code_batch_device(d_in.data_ptr(), meta.data_ptr(), meta.data_ptr() + 8 * n, n, d_out.data_ptr(),
                                  ooff.data_ptr(), ocap.data_ptr(), d_len.data_ptr(), d_status.data_ptr(), st)
            e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if it:
            best = ms if best is None else min(best, ms)
        print("iter %d: %.2f ms  %.2f GB/s out" % (it, ms, n * size / ms / 1e6))
    assert (d_status == 1).all() and (d_len == size).all()
    got = d_out.view(n, cap)[:, :size].reshape(-1).cpu().numpy()
    assert np.array_equal(got, data)
    print("OK n=%d size=%d cls=%d best %.2f ms = %.2f GB/s uncompressed" % (n, size, cls, best, n * size / best / 1e6))


if __name__ == "__main__":
    main()
