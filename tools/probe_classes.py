"""Developer probe: encoder phase times per corpus class (LZB_ENC_TIMING lines on stderr)."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["LZB_ENC_TIMING"] = "1"
lzb = importlib.import_module("lzma-java_b200")
from tools import corpus  # noqa: E402


def run(n, size, cls, fb, dict_size, iters=1):
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
    for it in range(iters + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(side):
            e0.record()
            enc.code_batch_device(d_in.data_ptr(), off.data_ptr(), ln.data_ptr(), n, size, d_out.data_ptr(), ooff.data_ptr(),
                                  ocap.data_ptr(), d_len.data_ptr(), True, side.cuda_stream)
            e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("PROBE n=%d size=%d cls=%d fb=%d iter %d: %.1f ms  %.1f MB/s ratio %.3f" %
              (n, size, cls, fb, it, ms, n * size / ms / 1e3, int(d_len.sum().item()) / (n * size)), flush=True)
    enc.close()


if __name__ == "__main__":
    # spec: n,size,cls,fb,dict[,iters][;ENV=VALUE...]   (environment knobs apply to that spec only)
    for spec in sys.argv[1:]:
        parts = spec.split(";")
        saved = {}
        for kv in parts[1:]:
            k, v = kv.split("=")
            saved[k] = os.environ.get(k)
            os.environ[k] = v
        nums = [int(x) for x in parts[0].split(",")]
        print("SPEC", spec, flush=True)
        run(*nums[:5], iters=nums[5] if len(nums) > 5 else 1)
        for k, v in saved.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v
