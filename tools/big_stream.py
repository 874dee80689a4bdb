"""Row f.4 evidence: one long stream through Encoder.Code on the GPU against the oracle.
python tools/big_stream.py <MiB> <dict> [fb]      (e.g. 64 8388608 ; 64 1048576)"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lzb = importlib.import_module("lzma-java_b200")
from oracle import oracle as O  # noqa: E402  (checker)
from tools import corpus  # noqa: E402


def main():
    mib, dict_size = int(sys.argv[1]), int(sys.argv[2])
    fb = int(sys.argv[3]) if len(sys.argv) > 3 else 32
    part = (mib << 20) // 4
    data = np.concatenate([corpus.generate(part, 1, c, 44, c) for c in range(4)])
    enc = lzb.Encoder()
    assert enc.SetDictionarySize(dict_size) and enc.SetNumFastBytes(fb)
    t0 = time.time()
    got = enc.code_bytes(data)
    t_gpu = time.time() - t0
    enc.close()
    t0 = time.time()
    ref = O.encode(data, O.props(dict_size=dict_size, fb=fb))
    t_cpu = time.time() - t0
    dec = lzb.Decoder()
    assert dec.SetDecoderProperties(O.props_bytes(O.props(dict_size=dict_size, fb=fb)))
    ok, back = dec.code_bytes(got, len(data))
    dec.close()
    print("BIG %d MiB dict %d fb %d: gpu %d B in %.1f s, oracle %d B in %.1f s (1 thread), equal=%s, gpu round trip=%s" %
          (mib, dict_size, fb, len(got), t_gpu, len(ref), t_cpu, got == ref, ok and back == data.tobytes()), flush=True)
    assert got == ref and ok and back == data.tobytes()


if __name__ == "__main__":
    main()
