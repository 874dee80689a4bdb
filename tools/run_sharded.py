"""BASELINE configs[4] (strong scaling): one synthetic corpus cut into fixed-size blocks, the blocks sharded over the
ranks of one node (rank r of W takes [r*B/W, (r+1)*B/W), no data-path collective; the ranks only
all_gather the per-block compressed sizes to place their output in the container).  Each rank encodes and then decodes
its own blocks on its own GPU; times are CUDA events, the maximum over ranks.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/run_sharded.py N_BLOCKS BLOCK_SIZE CLS FB DICT
  (N = 1 also runs without torchrun)
"""
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lzb = importlib.import_module("lzma-java_b200")
blocks = importlib.import_module("lzma-java_b200.blocks")
from tools import corpus  # noqa: E402


def main():
    n_total, size, cls, fb, dict_size = [int(x) for x in sys.argv[1:6]]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("gloo")  # carries 8 bytes per block and the timing reduction, nothing else
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    lo, hi = blocks.shard_range(n_total, rank, world)
    n = hi - lo
    host = torch.empty(n * size, dtype=torch.uint8).pin_memory()
    corpus.generate(size, n, cls, 5, first_block=lo, threads=max(1, (os.cpu_count() or 1) // world), out=host.numpy())
    d_in = host.to(dev)
    cap = lzb.enc_bound(size) + 13
    off = torch.arange(n, dtype=torch.int64, device=dev) * size
    ln = torch.full((n,), size, dtype=torch.int64, device=dev)
    ooff = torch.arange(n, dtype=torch.int64, device=dev) * cap
    ocap = torch.full((n,), cap, dtype=torch.int64, device=dev)
    d_out = torch.empty(n * cap, dtype=torch.uint8, device=dev)
    d_len = torch.zeros(n, dtype=torch.int64, device=dev)
    enc = lzb.Encoder(local)
    assert enc.SetDictionarySize(dict_size) and enc.SetNumFastBytes(fb) and enc.SetLcLpPb(3, 0, 2) and enc.SetMatchFinder(1)
    side = torch.cuda.Stream(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()

    def reduce_max(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    barrier()
    with torch.cuda.stream(side):
        e0.record()
        enc.code_batch_device(d_in.data_ptr(), off.data_ptr(), ln.data_ptr(), n, size, d_out.data_ptr(), ooff.data_ptr(),
                              ocap.data_ptr(), d_len.data_ptr(), True, side.cuda_stream)
        e1.record()
    barrier()
    enc_ms = reduce_max(e0.elapsed_time(e1))
    enc.close()
    # the one cross-rank exchange: compressed sizes -> every block's offset in the container
    csize = blocks.gather_sizes(d_len.cpu().numpy(), n_total, rank, world)
    offsets = blocks.exclusive_scan(csize)

    dcap = size + 288
    doff = torch.arange(n, dtype=torch.int64, device=dev) * dcap
    dcapt = torch.full((n,), dcap, dtype=torch.int64, device=dev)
    d_dec = torch.empty(n * dcap, dtype=torch.uint8, device=dev)
    d_dlen = torch.zeros(n, dtype=torch.int64, device=dev)
    d_status = torch.zeros(n, dtype=torch.int32, device=dev)
    dec = lzb.Decoder(local)
    dec_ms = None
    for _ in range(2):
        barrier()
        with torch.cuda.stream(side):
            e0.record()
            dec.code_batch_device(d_out.data_ptr(), ooff.data_ptr(), d_len.data_ptr(), n, d_dec.data_ptr(), doff.data_ptr(),
                                  dcapt.data_ptr(), d_dlen.data_ptr(), d_status.data_ptr(), side.cuda_stream)
            e1.record()
        barrier()
        t = reduce_max(e0.elapsed_time(e1))
        dec_ms = t if dec_ms is None else min(dec_ms, t)
    dec.close()
    assert bool((d_status == 1).all()) and bool((d_dlen == size).all())
    assert torch.equal(d_dec.view(n, dcap)[:, :size].reshape(-1), d_in), "round trip differs"
    if rank == 0:
        total = n_total * size
        print(json.dumps({"workload": "%d x %d B blocks, class %d, fb %d, dict %d, sharded over %d GPU(s)" % (n_total, size, cls, fb, dict_size, world),
                          "n_gpus": world, "encode_MBps": total / enc_ms / 1e3, "encode_ms": enc_ms,
                          "decode_MBps": total / dec_ms / 1e3, "decode_ms": dec_ms,
                          "compressed_ratio": float(csize.sum()) / total, "container_payload_end": int(offsets[-1] + csize[-1]),
                          "parity": "every rank's decoded blocks == its corpus blocks"}), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
