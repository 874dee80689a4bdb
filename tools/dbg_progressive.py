import sys, os, importlib, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lzb = importlib.import_module("lzma-java_b200")
from oracle import oracle
from tools import corpus
BASE = dict(dict_size=1 << 20, lc=3, lp=0, pb=2, fb=32, mf=1, eos=False)
n, size = 1300, 70000
data = corpus.generate(size, n, corpus.MIXED, 21)
off = np.arange(n, dtype=np.uint64) * size
ln = np.full(n, size, dtype=np.uint64)
comp, coff, clen = oracle.encode_batch(data, off, ln, oracle.props(**BASE), with_header=True, threads=8)
comp = comp.copy()
rng = np.random.default_rng(99)
flipped = rng.choice(n, 40, replace=False)
for i in flipped:
    comp[int(coff[i]) + 13 + int(rng.integers(0, int(clen[i]) - 13))] ^= 1 << int(rng.integers(0, 8))
clen = clen.copy()
trunc = rng.choice(n, 20, replace=False)
for i in trunc:
    clen[i] = clen[i] // 2
comp[int(coff[7])] = 225
clen[11] = 5
for pitch_extra, capx in ((15, 273), (0, 273)):
    pitch = size + capx + pitch_extra
    cap = np.full(n, size + capx, dtype=np.uint64)
    ooff = np.arange(n, dtype=np.uint64) * pitch
    ref_out, ref_len, ref_status = oracle.decode_batch(comp, coff, clen, ooff, cap, threads=8)
    dec = lzb.Decoder()
    out, out_len, status = dec.code_batch(comp, coff, clen, ooff, cap)
    dec.close()
    d = np.nonzero(status != ref_status)[0]
    print("pitch_extra", pitch_extra, "status diffs", len(d), [(int(i), int(status[i]), int(ref_status[i]), int(out_len[i]), int(ref_len[i]), i in flipped, i in trunc) for i in d[:10]])
    d = np.nonzero(out_len != ref_len)[0]
    print(" len diffs", len(d), d[:10])
    bad = [i for i in range(n) if not np.array_equal(out[int(ooff[i]):int(ooff[i]) + int(ref_len[i])], ref_out[int(ooff[i]):int(ooff[i]) + int(ref_len[i])])]
    print(" byte diffs", len(bad), bad[:10])
