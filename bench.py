#!/usr/bin/env python
"""bench.py -- the contract benchmark (see DESIGN.md "measurement").

Headline workload = BASELINE.json configs[1]: batch decode of 4096 independent
256 KiB .lzma streams (dict 1 MiB, lc3 lp0 pb2) per GPU.  One step = one pass
of the batch decoder over the whole batch.
  value     uncompressed MB/s (1e6 B/s) over all ranks, inputs resident in HBM,
            CUDA events on the launching stream, max over ranks
  e2e       the same batch through the host-buffer C-ABI call
            (lzb_dec_code_batch): pinned host input, H2D + kernels + D2H
  roofline  algorithmic bytes (sum N + sum C) / decode-call duration vs the
            measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline  the CPU oracle (a C port of the reference; no JVM exists on the
            box) decoding the same streams on all host cores
  encode    extra object: BASELINE.json configs[2] (1 MiB blocks, dict 1 MiB,
            fb 64, mixed corpus; 2048 blocks = 2 GiB per GPU) measured the same way
  c5        extra object: BASELINE.json configs[4], STRONG scaling: one 8 GiB corpus
            of 2048 x 4 MiB blocks (dict 4 MiB, fb 32) sharded over the ranks by
            blocks.shard_range; compressed sizes all-gathered, every rank copies
            its payloads to the scanned offsets of ONE host container (shared
            pinned memory), then decodes its blocks back out of that container
The compressed streams are produced by this repo's GPU encoder (bit-identical
to the reference encoder, tests/test_encode_gpu.py) during untimed set-up, and
the decoded bytes are compared with the corpus inside the run.

`--impl reference` times the reference's algorithm on the host cores only
(the oracle port; the Java reference cannot run here: no JVM), decode and encode.
"""
import argparse
import importlib
import json
import mmap
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DEC = dict(size=256 << 10, dict_size=1 << 20, fb=32, cls=0, config_id=2)   # configs[1]
ENC = dict(size=1 << 20, dict_size=1 << 20, fb=64, cls=4, config_id=3)     # configs[2]
C5 = dict(size=4 << 20, dict_size=1 << 22, fb=32, cls=4, config_id=5, blocks=2048)  # configs[4]
METRIC = "LZMA batch decode, uncompressed input MB/s (bit-exact vs reference)"
# dram__bytes_read.sum + dram__bytes_write.sum of one launch at the bench shape (profiles/, one ncu --set full capture each)
NCU_DRAM_BYTES_DECODE = 8.091469e9 + 1.174024e9   # profiles/r02_decode_hybrid_ncu.txt, 4096 x 256 KiB text streams
NCU_DRAM_BYTES_PARSE_PER_INPUT_BYTE = (6.117494e9 + 0.684309e9) / (2072 * 131072)  # profiles/r02_parse_full14_ncu.txt (lists position-ordered)


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def pack(out, off, ln):
    """Gather variable-length records into one contiguous array -> (packed, new offsets)."""
    noff = np.zeros(len(ln), dtype=np.uint64)
    if len(ln) > 1:
        noff[1:] = np.cumsum(ln)[:-1]
    packed = np.empty(int(ln.sum()), dtype=np.uint8)
    for i in range(len(ln)):
        packed[int(noff[i]): int(noff[i] + ln[i])] = out[int(off[i]): int(off[i] + ln[i])]
    return packed, noff


def cpu_sample_blocks(n, want):
    """`want` block indices spread over [0, n) in runs of four, so that every corpus class (block index mod 4)
    is sampled equally."""
    want = max(4, min(n, want) // 4 * 4)
    if want >= n:
        return np.arange(n)
    groups = want // 4
    starts = (np.arange(groups) * (n // 4) // groups) * 4
    return (starts[:, None] + np.arange(4)[None, :]).reshape(-1)


# --------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path (the C port: no JVM exists on the box) on all host cores:
    the headline decode on every stream of the workload, plus `encode` (C3) and `c5` on bounded class-balanced samples."""
    if rank != 0:
        return
    from oracle import oracle as O
    from tools import corpus
    threads = os.cpu_count() or 1
    n = args.streams
    size = DEC["size"]
    data = corpus.generate(size, n, DEC["cls"], DEC["config_id"])
    off = np.arange(n, dtype=np.uint64) * size
    ln = np.full(n, size, dtype=np.uint64)
    comp, coff, clen = O.encode_batch(data, off, ln, O.props(dict_size=DEC["dict_size"], fb=DEC["fb"]), True, threads)
    cap = np.full(n, size + 273, dtype=np.uint64)
    ooff = np.arange(n, dtype=np.uint64) * (size + 273)
    for _ in range(args.warmup):
        O.decode_batch(comp, coff, clen, ooff, cap, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out, olen, status = O.decode_batch(comp, coff, clen, ooff, cap, threads)
    dt = time.perf_counter() - t0
    assert (status == 1).all() and np.array_equal(out.reshape(n, size + 273)[:, :size].reshape(-1), data)
    v = args.steps * n * size / dt / 1e6
    sample = "all %d streams of the workload per step (C port of the reference decoder, %d pthreads; no JVM on the box)" % (n, threads)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "MB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args.streams, world),
            "cpu_baseline": {"value": v, "unit": "MB/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    del data, comp, out
    if not args.no_encode:
        line["encode"] = reference_encode(args, O, corpus, threads, ENC, args.enc_blocks,
                                          max(4 * threads, args.ref_enc_blocks), "C3")
    if not args.no_c5:
        line["c5"] = reference_encode(args, O, corpus, threads, C5, C5["blocks"], max(2 * threads, 32), "C5", decode=True)
    print(json.dumps(line), flush=True)


def reference_encode(args, O, corpus, threads, W, n_total, m, name, decode=False):
    """Oracle encoder (and decoder) on `m` class-balanced blocks of the workload, all host cores, one block per task."""
    size = W["size"]
    pick = cpu_sample_blocks(n_total, m)
    m = len(pick)
    data = np.empty(m * size, dtype=np.uint8)
    for j, b in enumerate(pick):
        corpus.generate(size, 1, W["cls"], W["config_id"], first_block=int(b), out=data[j * size:(j + 1) * size])
    off = np.arange(m, dtype=np.uint64) * size
    ln = np.full(m, size, dtype=np.uint64)
    p = O.props(dict_size=W["dict_size"], fb=W["fb"])
    steps = max(1, args.enc_steps)
    t0 = time.perf_counter()
    for _ in range(steps):
        comp, coff, clen = O.encode_batch(data, off, ln, p, True, threads)
    dt = time.perf_counter() - t0
    res = {"workload": "%s: %d of %d blocks of %d B per step, class-balanced (dict %d, fb %d, bt4, lc3 lp0 pb2)"
                       % (name, m, n_total, size, W["dict_size"], W["fb"]),
           "value": steps * m * size / dt / 1e6, "unit": "MB/s", "steps": steps, "ms_per_step": dt / steps * 1e3,
           "cores": threads, "kind": "port", "compressed_ratio": float(clen.sum()) / (m * size)}
    if decode:
        cap = np.full(m, size + 273, dtype=np.uint64)
        ooff = np.arange(m, dtype=np.uint64) * (size + 273)
        t0 = time.perf_counter()
        out, olen, status = O.decode_batch(comp, coff, clen, ooff, cap, threads)
        dt = time.perf_counter() - t0
        assert (status == 1).all() and np.array_equal(out.reshape(m, size + 273)[:, :size].reshape(-1), data)
        res["decode"] = {"value": m * size / dt / 1e6, "unit": "MB/s"}
    return res


def workload_config(streams, world):
    return {"workload": "C2: batch decode of %d independent 256 KiB .lzma streams per GPU (dict 1 MiB, lc3 lp0 pb2, "
                        "encoder fb 32, text-like synthetic corpus)" % streams,
            "streams_per_gpu": streams, "stream_bytes": DEC["size"], "dict": DEC["dict_size"], "lc": 3, "lp": 0, "pb": 2,
            "sharding": "independent streams, %d rank(s), no collective" % world,
            "l2_policy": "inputs larger than L2 (compressed batch > 300 MB, output 1 GiB per step)"}


# --------------------------------------------------------------------------- B200 arm
def gpu_encode(lzb, torch, dev, data_host, n, size, dict_size, fb, stream):
    """Untimed set-up: compress the corpus with the GPU encoder -> device tensors + host copies."""
    enc = lzb.Encoder(dev.index)
    assert enc.SetDictionarySize(dict_size) and enc.SetNumFastBytes(fb) and enc.SetLcLpPb(3, 0, 2) and enc.SetMatchFinder(1)
    cap = lzb.enc_bound(size) + lzb.HEADER_SIZE
    d_in = data_host.to(dev, non_blocking=True)
    off = torch.arange(n, dtype=torch.int64, device=dev) * size
    ln = torch.full((n,), size, dtype=torch.int64, device=dev)
    ooff = torch.arange(n, dtype=torch.int64, device=dev) * cap
    ocap = torch.full((n,), cap, dtype=torch.int64, device=dev)
    d_out = torch.empty(n * cap, dtype=torch.uint8, device=dev)
    d_len = torch.zeros(n, dtype=torch.int64, device=dev)
    torch.cuda.synchronize(dev)
    with torch.cuda.stream(stream):
        enc.code_batch_device(d_in.data_ptr(), off.data_ptr(), ln.data_ptr(), n, size, d_out.data_ptr(), ooff.data_ptr(),
                              ocap.data_ptr(), d_len.data_ptr(), True, stream.cuda_stream)
    torch.cuda.synchronize(dev)
    return enc, dict(d_in=d_in, off=off, ln=ln, ooff=ooff, ocap=ocap, d_out=d_out, d_len=d_len, cap=cap)


def timed_steps(torch, dev, stream, steps, warmup, fn, barrier):
    """W untimed + K timed steps; per-step CUDA events on the launching stream."""
    for _ in range(warmup):
        with torch.cuda.stream(stream):
            fn()
    torch.cuda.synchronize(dev)
    barrier()
    torch.cuda.synchronize(dev)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    with torch.cuda.stream(stream):
        evs[0].record()
        for i in range(steps):
            fn()
            evs[i + 1].record()
    torch.cuda.synchronize(dev)
    barrier()
    per = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
    return evs[0].elapsed_time(evs[steps]), per


class Ranks:
    """The process group of the contract: barrier and reductions of timings only (gloo; the data path has no
    collective -- independent streams; C5 all-gathers 8 bytes per block through blocks.gather_sizes)."""

    def __init__(self, world):
        self.world, self.dist, self.torch = world, None, None
        if world > 1:
            import torch
            import torch.distributed as dist
            self.dist, self.torch = dist, torch
            dist.init_process_group("gloo")

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def _red(self, x, op):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._red(x, self.dist.ReduceOp.MAX) if self.dist else x

    def sum(self, x):
        return self._red(x, self.dist.ReduceOp.SUM) if self.dist else x

    def bcast_obj(self, obj):
        if self.dist is None:
            return obj
        box = [obj]
        self.dist.broadcast_object_list(box, src=0)
        return box[0]

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def run_b200(args, rank, world, local_rank):
    import torch
    lzb = importlib.import_module("lzma-java_b200")
    lzb.lib()  # fails loudly when the CUDA library is missing: there is no fallback
    from tools import corpus
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    R = Ranks(world)
    barrier, max_over_ranks = R.barrier, R.max

    stream = torch.cuda.Stream(dev)  # a non-default stream: handle 0 would select the codec's own stream
    threads = max(1, (os.cpu_count() or 1) // world)
    launches0 = lzb.kernel_launches()

    # ---------------- set-up: corpus + GPU encode (untimed)
    n, size = args.streams, DEC["size"]
    host = torch.empty(n * size, dtype=torch.uint8).pin_memory()
    corpus.generate(size, n, DEC["cls"], DEC["config_id"], first_block=rank * n, threads=threads, out=host.numpy())
    enc, E = gpu_encode(lzb, torch, dev, host, n, size, DEC["dict_size"], DEC["fb"], stream)
    clen = E["d_len"].cpu().numpy().astype(np.uint64)
    assert (clen > 0).all() and (clen < 2 ** 62).all(), "GPU encode failed"
    comp_host, coff = pack(E["d_out"].cpu().numpy(), (np.arange(n, dtype=np.uint64) * E["cap"]), clen)
    d_data = E["d_in"]
    del E
    enc.close()
    total_c = int(clen.sum())

    comp_pinned = torch.from_numpy(comp_host).pin_memory()
    d_comp = comp_pinned.to(dev)
    d_coff = torch.from_numpy(coff.astype(np.int64)).to(dev)
    d_clen = torch.from_numpy(clen.astype(np.int64)).to(dev)
    cap = size + 288
    d_ooff = torch.arange(n, dtype=torch.int64, device=dev) * cap
    d_ocap = torch.full((n,), cap, dtype=torch.int64, device=dev)
    d_out = torch.empty(n * cap, dtype=torch.uint8, device=dev)
    d_olen = torch.zeros(n, dtype=torch.int64, device=dev)
    d_status = torch.zeros(n, dtype=torch.int32, device=dev)
    dec = lzb.Decoder(dev.index)

    def step_dec():
        dec.code_batch_device(d_comp.data_ptr(), d_coff.data_ptr(), d_clen.data_ptr(), n, d_out.data_ptr(), d_ooff.data_ptr(),
                              d_ocap.data_ptr(), d_olen.data_ptr(), d_status.data_ptr(), stream.cuda_stream)

    # ---------------- headline: device-resident decode
    l0 = lzb.kernel_launches()
    with ClockSampler(local_rank) as clk:
        total_ms, per = timed_steps(torch, dev, stream, args.steps, args.warmup, step_dec, barrier)
    timed_launches = (lzb.kernel_launches() - l0) * args.steps // (args.steps + args.warmup)
    assert bool((d_status == 1).all()) and bool((d_olen == size).all()), "decode status"
    assert torch.equal(d_out.view(n, cap)[:, :size].reshape(-1), d_data), "decoded bytes differ from the corpus"
    total_ms = max_over_ranks(total_ms)
    value = world * args.steps * n * size / (total_ms * 1e-3) / 1e6
    kernel_ms = statistics.mean(per)
    peak, peak_src = hbm_peak()
    achieved = (n * size + total_c) / (kernel_ms * 1e-3) / 1e9
    del d_out

    # ---------------- e2e: host buffers through the C-ABI batch call (H2D + kernels + D2H)
    out_pinned = torch.empty(n * cap, dtype=torch.uint8).pin_memory()
    ooff_h = (np.arange(n, dtype=np.uint64) * cap)
    ocap_h = np.full(n, cap, dtype=np.uint64)
    olen_h = np.zeros(n, dtype=np.uint64)
    status_h = np.zeros(n, dtype=np.int32)
    L = lzb.lib()

    def step_e2e():
        rc = L.lzb_dec_code_batch(dec._h, comp_pinned.data_ptr(), coff.ctypes.data, clen.ctypes.data, n, out_pinned.data_ptr(),
                                  ooff_h.ctypes.data, ocap_h.ctypes.data, olen_h.ctypes.data, status_h.ctypes.data)
        assert rc == 1, lzb.last_error()

    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    torch.cuda.synchronize(dev)
    barrier()
    out_pinned.zero_()  # the timed steps must produce every byte again
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize(dev)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    assert (status_h == 1).all() and (olen_h == size).all()
    # the WHOLE output of the host-buffer path (progressive read-back) against the corpus
    assert np.array_equal(out_pinned.numpy().reshape(n, cap)[:, :size], host.numpy().reshape(n, size)), "e2e output differs"
    e2e = world * args.steps * n * size / e2e_s / 1e6
    dec.close()

    # ---------------- CPU baseline (rank 0, N=1 only): the oracle on all host cores, same streams
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as O
        t0 = time.perf_counter()
        o_out, o_len, o_status = O.decode_batch(comp_host, coff, clen, ooff_h, ocap_h, os.cpu_count())
        dt = time.perf_counter() - t0
        assert (o_status == 1).all() and np.array_equal(o_out.reshape(n, cap)[:, :size].reshape(-1), host.numpy())
        cpu = {"value": n * size / dt / 1e6, "unit": "MB/s", "cores": os.cpu_count(), "kind": "port",
               "sample": "all %d streams of the step, once (C port of the reference decoder, one pthread per core; "
                         "the Java reference cannot run: no JVM on the box)" % n}
        del o_out
    del out_pinned, comp_pinned, d_comp, d_data, host

    # ---------------- encode extra (configs[2]) and the sharded corpus (configs[4])
    encode = None
    if not args.no_encode:
        encode = bench_encode(args, lzb, torch, corpus, dev, stream, rank, world, threads, R, peak)
    c5 = None
    if not args.no_c5:
        c5 = bench_c5(args, lzb, torch, corpus, dev, stream, rank, world, local_rank, threads, R, peak)

    line = {"metric": METRIC, "value": value, "unit": "MB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": workload_config(n, world),
            "e2e": {"value": e2e, "unit": "MB/s", "h2d_bytes_per_step": total_c + 32 * n, "d2h_bytes_per_step": n * cap + 12 * n},
            "gpu_launches": int(timed_launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_DRAM_BYTES_DECODE if n == 4096 else None,
                         "traffic_source": "profiles/r02_decode_hybrid_ncu.txt (dram__bytes_read+write, one ncu --set full capture of this launch shape)",
                         "peak_source": peak_src, "kernel": "lzb_decode_kernel<kDecHybrid>",
                         "algorithmic_bytes_per_launch": n * size + total_c, "kernel_ms": kernel_ms,
                         "note": "serial range-decoder chains: issue/latency bound, not HBM bound (profiles/)"},
            "clocks": clk.summary(), "compressed_ratio": total_c / (n * size),
            "parity": "decoded bytes == corpus for every stream, device-resident and through the host-buffer call (whole 1 GiB "
                      "compared); GPU-encoded streams (bit-identical to the oracle)"}
    if cpu:
        line["cpu_baseline"] = cpu
    if encode:
        line["encode"] = encode
    if c5:
        line["c5"] = c5
    line["gpu_launches_total"] = int(lzb.kernel_launches() - launches0)
    if rank == 0:
        print(json.dumps(line), flush=True)
    R.close()


def gpu_roundtrip(lzb, torch, dev, stream, E, n, size):
    """Decode every compressed block of E on the GPU and compare with the input that produced it."""
    dcap = size + 288
    doff = torch.arange(n, dtype=torch.int64, device=dev) * dcap
    dcapt = torch.full((n,), dcap, dtype=torch.int64, device=dev)
    d_dec = torch.empty(n * dcap, dtype=torch.uint8, device=dev)
    d_dlen = torch.zeros(n, dtype=torch.int64, device=dev)
    d_status = torch.zeros(n, dtype=torch.int32, device=dev)
    dec = lzb.Decoder(dev.index)
    with torch.cuda.stream(stream):
        dec.code_batch_device(E["d_out"].data_ptr(), E["ooff"].data_ptr(), E["d_len"].data_ptr(), n, d_dec.data_ptr(),
                              doff.data_ptr(), dcapt.data_ptr(), d_dlen.data_ptr(), d_status.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize(dev)
    dec.close()
    ok = bool((d_status == 1).all()) and bool((d_dlen == size).all()) and \
        torch.equal(d_dec.view(n, dcap)[:, :size].reshape(-1), E["d_in"])
    return ok


def bench_encode(args, lzb, torch, corpus, dev, stream, rank, world, threads, R, peak):
    n, size = args.enc_blocks, ENC["size"]
    host = torch.empty(n * size, dtype=torch.uint8).pin_memory()
    corpus.generate(size, n, ENC["cls"], ENC["config_id"], first_block=rank * n, threads=threads, out=host.numpy())
    enc, E = gpu_encode(lzb, torch, dev, host, n, size, ENC["dict_size"], ENC["fb"], stream)  # doubles as warm-up

    def step():
        enc.code_batch_device(E["d_in"].data_ptr(), E["off"].data_ptr(), E["ln"].data_ptr(), n, size, E["d_out"].data_ptr(),
                              E["ooff"].data_ptr(), E["ocap"].data_ptr(), E["d_len"].data_ptr(), True, stream.cuda_stream)

    l0 = lzb.kernel_launches()
    total_ms, per = timed_steps(torch, dev, stream, args.enc_steps, 0, step, R.barrier)
    launches = lzb.kernel_launches() - l0
    total_ms = R.max(total_ms)
    clen = E["d_len"].cpu().numpy().astype(np.uint64)
    total_c = int(clen.sum())
    value = world * args.enc_steps * n * size / (total_ms * 1e-3) / 1e6
    achieved = (n * size + total_c) / (statistics.mean(per) * 1e-3) / 1e9
    # every block: the GPU decoder gives the corpus back (size-independent property, all n blocks)
    assert gpu_roundtrip(lzb, torch, dev, stream, E, n, size), "encode round trip differs from the corpus"

    # e2e through the host-buffer call
    cap = E["cap"]
    out_pinned = torch.empty(n * cap, dtype=torch.uint8).pin_memory()
    off_h = np.arange(n, dtype=np.uint64) * size
    len_h = np.full(n, size, dtype=np.uint64)
    ooff_h = np.arange(n, dtype=np.uint64) * cap
    ocap_h = np.full(n, cap, dtype=np.uint64)
    olen_h = np.zeros(n, dtype=np.uint64)
    L = lzb.lib()
    torch.cuda.synchronize(dev)
    R.barrier()
    e2e_steps = 1
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        rc = L.lzb_enc_code_batch(enc._h, host.data_ptr(), off_h.ctypes.data, len_h.ctypes.data, n, out_pinned.data_ptr(),
                                  ooff_h.ctypes.data, ocap_h.ctypes.data, olen_h.ctypes.data, 1)
        assert rc == 1, lzb.last_error()
    e2e_s = R.max(time.perf_counter() - t0)
    assert np.array_equal(olen_h, clen)
    enc.close()

    res = {"workload": "C3: block encode, %d x 1 MiB blocks per GPU, dict 1 MiB, fb 64, bt4, lc3 lp0 pb2, mixed "
                       "text/binary/random/repetitive corpus" % n,
           "value": value, "unit": "MB/s", "steps": args.enc_steps, "ms_per_step": total_ms / args.enc_steps,
           "e2e": {"value": world * e2e_steps * n * size / e2e_s / 1e6, "unit": "MB/s", "steps": e2e_steps,
                   "h2d_bytes_per_step": n * size + 32 * n, "d2h_bytes_per_step": total_c + 8 * n},
           "gpu_launches": int(launches),
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": NCU_DRAM_BYTES_PARSE_PER_INPUT_BYTE * n * size,
                        "traffic_source": "profiles/r02_parse_full14_ncu.txt: dram bytes of lzb_parse_kernel per input byte "
                                          "(2072 x 128 KiB of this corpus mix), scaled to this launch",
                        "kernel": "lzb_parse_kernel (dominant), lzb_mf_long_kernel, lzb_mf_tree_kernel, lzb_mf_link_kernel"},
           "compressed_ratio": total_c / (n * size),
           "parity": "every block: GPU decode of the GPU-encoded stream == corpus"}
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        pick = cpu_sample_blocks(n, min(n, max(args.enc_cpu_blocks, 128 * cores)))  # all 2048 blocks from 16 cores up
        m = len(pick)
        sub = np.ascontiguousarray(host.numpy().reshape(n, size)[pick]).reshape(-1)
        t0 = time.perf_counter()
        r_out, r_off, r_len = O.encode_batch(sub, np.arange(m, dtype=np.uint64) * size, np.full(m, size, dtype=np.uint64),
                                             O.props(dict_size=ENC["dict_size"], fb=ENC["fb"]), True, cores)
        dt = time.perf_counter() - t0
        g = out_pinned.numpy()
        assert np.array_equal(olen_h[pick], r_len), "GPU encoder: compressed lengths differ from the oracle"
        same = all(np.array_equal(g[int(ooff_h[b]): int(ooff_h[b] + olen_h[b])], r_out[int(r_off[j]): int(r_off[j] + r_len[j])])
                   for j, b in enumerate(pick))
        assert same, "GPU encoder output differs from the oracle"
        res["cpu_baseline"] = {"value": m * size / dt / 1e6, "unit": "MB/s", "cores": cores, "kind": "port",
                               "sample": "%d of %d blocks, class-balanced (C port of the reference encoder, one pthread per core)" % (m, n)}
        res["parity"] += "; compressed bytes of %d of %d blocks (class-balanced) == oracle in this run" % (m, n)
    return res


def shared_container(torch, R, rank, world, nbytes, tag):
    """ONE host buffer every rank can write: pinned memory at N=1, a /dev/shm mapping registered with CUDA otherwise."""
    if world == 1:
        t = torch.empty(max(nbytes, 1), dtype=torch.uint8).pin_memory()
        return t, (lambda: None)
    path = R.bcast_obj("/dev/shm/lzb_c5_%s_%d_%s" % (os.environ.get("MASTER_PORT", "0"), os.getpid(), tag) if rank == 0 else None)
    if rank == 0:
        with open(path, "wb") as f:
            f.truncate(max(nbytes, 1))
    R.barrier()
    f = open(path, "r+b")
    mm = mmap.mmap(f.fileno(), max(nbytes, 1))
    t = torch.frombuffer(mm, dtype=torch.uint8)
    rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), t.numel(), 0)
    assert "success" in str(rc).lower() or str(rc) == "0", "cudaHostRegister failed: %s" % rc

    def release():
        torch.cuda.cudart().cudaHostUnregister(t.data_ptr())
        R.barrier()
        if rank == 0:
            os.unlink(path)
    return t, release


def bench_c5(args, lzb, torch, corpus, dev, stream, rank, world, local_rank, threads, R, peak):
    """configs[4]: 2048 x 4 MiB blocks (8 GiB), sharded over the ranks (strong scaling), encode and decode."""
    blocks = importlib.import_module("lzma-java_b200.blocks")
    n_total, size = args.c5_blocks, C5["size"]
    lo, hi = blocks.shard_range(n_total, rank, world)
    n = hi - lo
    host = torch.empty(n * size, dtype=torch.uint8).pin_memory()
    corpus.generate(size, n, C5["cls"], C5["config_id"], first_block=lo, threads=threads, out=host.numpy())
    enc = lzb.Encoder(dev.index)
    assert enc.SetDictionarySize(C5["dict_size"]) and enc.SetNumFastBytes(C5["fb"]) and enc.SetLcLpPb(3, 0, 2) and enc.SetMatchFinder(1)
    cap = lzb.enc_bound(size) + lzb.HEADER_SIZE
    off = torch.arange(n, dtype=torch.int64, device=dev) * size
    ln = torch.full((n,), size, dtype=torch.int64, device=dev)
    ooff = torch.arange(n, dtype=torch.int64, device=dev) * cap
    ocap = torch.full((n,), cap, dtype=torch.int64, device=dev)
    d_in = torch.empty(n * size, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n * cap, dtype=torch.uint8, device=dev)
    d_len = torch.zeros(n, dtype=torch.int64, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lzb.kernel_launches()

    # ---- encode, end to end: host corpus -> H2D -> kernels -> sizes gathered -> payloads at their container offsets
    torch.cuda.synchronize(dev)
    R.barrier()
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        d_in.copy_(host, non_blocking=True)
        e0.record()
        enc.code_batch_device(d_in.data_ptr(), off.data_ptr(), ln.data_ptr(), n, size, d_out.data_ptr(), ooff.data_ptr(),
                              ocap.data_ptr(), d_len.data_ptr(), True, stream.cuda_stream)
        e1.record()
    torch.cuda.synchronize(dev)
    enc_ms = R.max(e0.elapsed_time(e1))
    mine = d_len.cpu().numpy()
    assert (mine > 0).all() and (mine < 2 ** 62).all(), "GPU encode failed"
    # the one cross-rank exchange: 8 bytes per block -> every block's offset in the container
    csize = blocks.gather_sizes(mine, n_total, rank, world)
    offsets = blocks.exclusive_scan(csize)
    total_c = int(csize.sum())
    container, release = shared_container(torch, R, rank, world, total_c, "a")
    with torch.cuda.stream(stream):
        for i in range(n):
            a = int(offsets[lo + i])
            container[a: a + int(mine[i])].copy_(d_out[i * cap: i * cap + int(mine[i])], non_blocking=True)
    torch.cuda.synchronize(dev)
    R.barrier()
    enc_e2e_s = R.max(time.perf_counter() - t0)
    enc.close()
    launches = lzb.kernel_launches() - l0
    del d_out

    # ---- decode back OUT OF THE CONTAINER: every rank takes its block range from the shared index
    dcap = size + 288
    my_off = offsets[lo:hi].copy()
    my_len = csize[lo:hi].copy()
    ooff_h = np.arange(n, dtype=np.uint64) * dcap
    ocap_h = np.full(n, dcap, dtype=np.uint64)
    olen_h = np.zeros(n, dtype=np.uint64)
    status_h = np.zeros(n, dtype=np.int32)
    out_pinned = torch.empty(n * dcap, dtype=torch.uint8).pin_memory()
    dec = lzb.Decoder(dev.index)
    L = lzb.lib()
    dec_e2e_s = None
    for it in range(2):  # first pass warms the handle's buffers
        out_pinned.zero_()
        torch.cuda.synchronize(dev)
        R.barrier()
        t0 = time.perf_counter()
        rc = L.lzb_dec_code_batch(dec._h, container.data_ptr(), my_off.ctypes.data, my_len.ctypes.data, n, out_pinned.data_ptr(),
                                  ooff_h.ctypes.data, ocap_h.ctypes.data, olen_h.ctypes.data, status_h.ctypes.data)
        assert rc == 1, lzb.last_error()
        dec_e2e_s = R.max(time.perf_counter() - t0)
    assert (status_h == 1).all() and (olen_h == size).all()
    assert np.array_equal(out_pinned.numpy().reshape(n, dcap)[:, :size], host.numpy().reshape(n, size)), "C5 round trip differs"
    del out_pinned
    # device-resident decode of the same slice of the container
    a, b = (int(offsets[lo]), int(offsets[hi - 1] + csize[hi - 1])) if n else (0, 0)
    d_comp = container[a:b].to(dev)
    d_coff = torch.from_numpy((my_off - np.uint64(a)).astype(np.int64)).to(dev)
    d_clen = torch.from_numpy(my_len.astype(np.int64)).to(dev)
    d_ooff = torch.arange(n, dtype=torch.int64, device=dev) * dcap
    d_ocap = torch.full((n,), dcap, dtype=torch.int64, device=dev)
    d_dec = torch.empty(n * dcap, dtype=torch.uint8, device=dev)
    d_olen = torch.zeros(n, dtype=torch.int64, device=dev)
    d_status = torch.zeros(n, dtype=torch.int32, device=dev)

    def step_dec():
        dec.code_batch_device(d_comp.data_ptr(), d_coff.data_ptr(), d_clen.data_ptr(), n, d_dec.data_ptr(), d_ooff.data_ptr(),
                              d_ocap.data_ptr(), d_olen.data_ptr(), d_status.data_ptr(), stream.cuda_stream)

    dec_total_ms, _ = timed_steps(torch, dev, stream, 2, 1, step_dec, R.barrier)
    dec_ms = R.max(dec_total_ms) / 2
    assert bool((d_status == 1).all()) and torch.equal(d_dec.view(n, dcap)[:, :size].reshape(-1), d_in)
    dec.close()
    release()

    total = n_total * size
    res = {"workload": "C5: one %d x 4 MiB corpus (%.1f GiB), dict 4 MiB, fb 32, mixed classes, blocks [r*B/N, (r+1)*B/N) per rank"
                       % (n_total, total / 2 ** 30),
           "n_gpus": world, "scaling": "strong", "blocks_per_gpu": n, "unit": "MB/s",
           "encode": {"value": total / enc_ms / 1e3, "ms": enc_ms, "e2e": total / enc_e2e_s / 1e6,
                      "roofline_frac": (total + total_c) / (enc_ms * 1e-3) / 1e9 / (world * peak),
                      "e2e_includes": "H2D of the shard, kernels, all_gather of %d sizes, D2H of every payload to its scanned "
                                      "offset in the shared container" % n_total},
           "decode": {"value": total / dec_ms / 1e3, "ms": dec_ms, "e2e": total / dec_e2e_s / 1e6,
                      "roofline_frac": (total + total_c) / (dec_ms * 1e-3) / 1e9 / (world * peak),
                      "e2e_includes": "host-buffer call reading the rank's streams from the shared container, progressive read-back"},
           "container_bytes": total_c, "compressed_ratio": total_c / total, "gpu_launches": int(launches),
           "parity": "every rank: blocks decoded out of the shared container == its corpus blocks (host-buffer and device paths)"}
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        pick = cpu_sample_blocks(n, max(2 * cores, 32))
        m = len(pick)
        sub = np.ascontiguousarray(host.numpy().reshape(n, size)[pick]).reshape(-1)
        p = O.props(dict_size=C5["dict_size"], fb=C5["fb"])
        t0 = time.perf_counter()
        r_out, r_off, r_len = O.encode_batch(sub, np.arange(m, dtype=np.uint64) * size, np.full(m, size, dtype=np.uint64), p, True, cores)
        t_enc = time.perf_counter() - t0
        assert np.array_equal(r_len, csize[lo:hi][pick]), "C5: compressed lengths differ from the oracle"
        g = container.numpy() if world == 1 else None
        if g is not None:
            for j, bidx in enumerate(pick):
                o = int(offsets[lo + bidx])
                assert np.array_equal(g[o: o + int(r_len[j])], r_out[int(r_off[j]): int(r_off[j] + r_len[j])]), "C5 block %d != oracle" % bidx
        ocap_s = np.full(m, size + 273, dtype=np.uint64)
        t0 = time.perf_counter()
        _, _, st = O.decode_batch(r_out, r_off, r_len, np.arange(m, dtype=np.uint64) * (size + 273), ocap_s, cores)
        t_dec = time.perf_counter() - t0
        assert (st == 1).all()
        res["cpu_baseline"] = {"encode": m * size / t_enc / 1e6, "decode": m * size / t_dec / 1e6, "unit": "MB/s", "cores": cores,
                               "kind": "port", "sample": "%d of %d blocks, class-balanced, one pthread per core; blocks == oracle" % (m, n_total)}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=4096, help="decode streams per GPU (configs[1]: 4096)")
    ap.add_argument("--enc-blocks", type=int, default=2048, help="1 MiB blocks per GPU for the encode extra (configs[2] fixes the block size, not the count)")
    ap.add_argument("--enc-steps", type=int, default=2)
    ap.add_argument("--enc-cpu-blocks", type=int, default=128)
    ap.add_argument("--ref-enc-blocks", type=int, default=256, help="reference arm: C3 blocks per encode step (at least 4 per core)")
    ap.add_argument("--c5-blocks", type=int, default=C5["blocks"], help="4 MiB blocks of the sharded corpus (configs[4]: 2048 = 8 GiB)")
    ap.add_argument("--no-encode", action="store_true")
    ap.add_argument("--no-c5", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3  # timing rules: W >= 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
