"""Encode then decode one BASELINE configuration on the GPU (device-resident timing, one pass each),
verifying the round trip.  python tools/run_config.py N SIZE CLS FB DICT [check_blocks]"""
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lzb = importlib.import_module("lzma-java_b200")
from tools import corpus  # noqa: E402


def main():
    n, size, cls, fb, dict_size = [int(x) for x in sys.argv[1:6]]
    check = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    dev = torch.device("cuda:0")
    host = torch.empty(n * size, dtype=torch.uint8).pin_memory()
    corpus.generate(size, n, cls, 5, out=host.numpy())
    d_in = host.to(dev)
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
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(side):
        e0.record()
        enc.code_batch_device(d_in.data_ptr(), off.data_ptr(), ln.data_ptr(), n, size, d_out.data_ptr(), ooff.data_ptr(),
                              ocap.data_ptr(), d_len.data_ptr(), True, side.cuda_stream)
        e1.record()
    torch.cuda.synchronize()
    enc_ms = e0.elapsed_time(e1)
    enc.close()
    csum = int(d_len.sum().item())
    print("encode: %d x %d B cls %d fb %d dict %d: %.1f ms = %.1f MB/s, ratio %.3f" % (n, size, cls, fb, dict_size, enc_ms,
                                                                                      n * size / enc_ms / 1e3, csum / (n * size)), flush=True)
    if check:
        from oracle import oracle as O
        t = time.time()
        m = min(check, n)
        hoff = np.arange(m, dtype=np.uint64) * size
        hlen = np.full(m, size, dtype=np.uint64)
        r_out, r_off, r_len = O.encode_batch(host.numpy()[: m * size], hoff, hlen, O.props(dict_size=dict_size, fb=fb), True, os.cpu_count())
        g = d_out.cpu().numpy()
        gl = d_len.cpu().numpy()
        for i in range(m):
            assert gl[i] == r_len[i] and np.array_equal(g[i * cap: i * cap + int(gl[i])], r_out[int(r_off[i]): int(r_off[i] + r_len[i])]), i
        dt = time.time() - t
        print("oracle check of %d blocks OK (CPU %d threads: %.1f MB/s)" % (m, os.cpu_count(), m * size / dt / 1e6), flush=True)
    # decode what was just encoded (streams stay where the encoder wrote them)
    dcap = size + 288
    doff = torch.arange(n, dtype=torch.int64, device=dev) * dcap
    dcapt = torch.full((n,), dcap, dtype=torch.int64, device=dev)
    d_dec = torch.empty(n * dcap, dtype=torch.uint8, device=dev)
    d_dlen = torch.zeros(n, dtype=torch.int64, device=dev)
    d_status = torch.zeros(n, dtype=torch.int32, device=dev)
    dec = lzb.Decoder()
    best = None
    for it in range(2):
        with torch.cuda.stream(side):
            e0.record()
            dec.code_batch_device(d_out.data_ptr(), ooff.data_ptr(), d_len.data_ptr(), n, d_dec.data_ptr(), doff.data_ptr(),
                                  dcapt.data_ptr(), d_dlen.data_ptr(), d_status.data_ptr(), side.cuda_stream)
            e1.record()
        torch.cuda.synchronize()
        best = e0.elapsed_time(e1) if best is None else min(best, e0.elapsed_time(e1))
    dec.close()
    assert bool((d_status == 1).all()) and bool((d_dlen == size).all())
    assert torch.equal(d_dec.view(n, dcap)[:, :size].reshape(-1), d_in)
    print("decode: %.1f ms = %.1f MB/s, round trip bit-exact" % (best, n * size / best / 1e3), flush=True)


if __name__ == "__main__":
    main()
