"""Developer probe: encode throughput of N blocks with inputs resident in HBM (for ncu captures)."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lzb = importlib.import_module("lzma-java_b200")
from tools import corpus  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
    cls = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    iters = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    fb = int(sys.argv[5]) if len(sys.argv) > 5 else 64
    dict_size = int(sys.argv[6]) if len(sys.argv) > 6 else 1 << 20
    data = corpus.generate(size, n, cls, 3)
    dev = torch.device("cuda:0")
    d_in = torch.from_numpy(data).to(dev)
    cap = lzb.enc_bound(size) + 13
    off = torch.arange(n, dtype=torch.int64, device=dev) * size
    ln = torch.full((n,), size, dtype=torch.int64, device=dev)
    ooff = torch.arange(n, dtype=torch.int64, device=dev) * cap
    ocap = torch.full((n,), cap, dtype=torch.int64, device=dev)
    d_out = torch.empty(n * cap, dtype=torch.uint8, device=dev)
    d_len = torch.zeros(n, dtype=torch.int64, device=dev)
    enc = lzb.Encoder()
    assert enc.SetDictionarySize(dict_size) and enc.SetNumFastBytes(fb) and enc.SetLcLpPb(3, 0, 2) and enc.SetMatchFinder(1)
    side = torch.cuda.Stream()
    torch.cuda.synchronize()
    best = None
    for it in range(iters + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(side):
            e0.record()
            enc.code_batch_device(d_in.data_ptr(), off.data_ptr(), ln.data_ptr(), n, size, d_out.data_ptr(), ooff.data_ptr(),
                                  ocap.data_ptr(), d_len.data_ptr(), True, side.cuda_stream)
            e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if it:
            best = ms if best is None else min(best, ms)
        print("iter %d: %.1f ms  %.3f GB/s in" % (it, ms, n * size / ms / 1e6), flush=True)
    csum = int(d_len.sum().item())
    best = best if best is not None else ms
    print("OK n=%d size=%d cls=%d fb=%d best %.1f ms = %.3f GB/s, ratio %.3f" % (n, size, cls, fb, best, n * size / best / 1e6, csum / (n * size)))


if __name__ == "__main__":
    main()
