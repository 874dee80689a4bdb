"""Generates tests/golden/oracle_vectors.json: (input sha256, props, LzmaAlone
length, sha256) for synthetic blocks, produced by the CPU oracle AFTER it
reproduced the reference's 12 firefox.exe vectors in this container (the
script refuses to run otherwise).  The fixture travels to the GPU box, where
/root/reference does not exist."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O  # noqa: E402
from tools import corpus  # noqa: E402
from test_oracle_golden import CLI_DEFAULTS, VECTORS  # noqa: E402

FIREFOX = "/root/reference/src/test/java/SevenZip/firefox.exe"


def main():
    data = open(FIREFOX, "rb").read()
    for _, kw, length, md5 in VECTORS:
        d = dict(CLI_DEFAULTS)
        d.update(kw)
        s = O.encode(data, O.props(**d), alone=True)
        assert len(s) == length and hashlib.md5(s).hexdigest() == md5, "oracle is not pinned; refusing"
    cases = []
    base = dict(dict_size=1 << 20, lc=3, lp=0, pb=2, fb=32, mf=1, eos=False)
    variants = [
        {}, {"fb": 64}, {"fb": 273, "dict_size": 1 << 23}, {"dict_size": 1 << 22}, {"dict_size": 1 << 23, "fb": 128},
        {"fb": 5}, {"mf": 0}, {"eos": True}, {"lc": 0}, {"lc": 8}, {"lp": 1}, {"lp": 4}, {"pb": 0}, {"pb": 4},
        {"dict_size": 1}, {"dict_size": 4096}, {"dict_size": 100000, "fb": 48},
    ]
    k = 0
    for vi, v in enumerate(variants):
        for cls in ([0, 1, 2, 3] if vi < 3 else [vi % 4]):
            size = [65536, 40000, 20000, 100000][cls] if vi else 262144
            p = dict(base)
            p.update(v)
            blk = corpus.generate(size, 1, cls, 100, k)
            s = O.encode(blk, O.props(**p), alone=True)
            ok, back = O.decode_alone(s)
            assert ok == 1 and back == blk.tobytes()
            cases.append({"cls": cls, "config_id": 100, "block": k, "size": size, "props": p,
                          "in_sha256": hashlib.sha256(blk.tobytes()).hexdigest(), "out_len": len(s),
                          "out_sha256": hashlib.sha256(s).hexdigest()})
            k += 1
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_vectors.json")
    json.dump({"generator": "tests/golden/make_golden.py", "oracle": "oracle/lzma_oracle.c",
               "pinned_by": "12 firefox.exe vectors of LzmaAloneTest.java:27-38", "cases": cases}, open(out, "w"), indent=1)
    print(len(cases), "cases ->", out)


if __name__ == "__main__":
    main()
