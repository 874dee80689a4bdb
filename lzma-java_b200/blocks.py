"""Block container and sharding for the block codec.

`Encoder.Code` emits a payload only, and a concatenation of block streams is not
self-delimiting, so a multi-block file needs an index.  The container is

    magic "LZB1" | u32 n_blocks | u64 block_size | u64 total_size
    n_blocks x (u64 compressed_size, u64 uncompressed_size)
    block 0 .lzma | block 1 .lzma | ...

where every block is a standalone LzmaAlone file (5 property bytes + LE64 size
+ payload, LzmaAlone.java:208-217) that the reference's own `LzmaAlone d` can
decode after being cut out.

Sharding: blocks are independent streams, so rank r of W takes the contiguous
range [r*B/W, (r+1)*B/W).  The only cross-rank datum is the per-block
compressed size (8 bytes per block); ranks exchange it with one all_gather on
whatever torch.distributed backend is initialised and derive every output
offset with an exclusive scan.  There is no data-path collective.
"""
import struct

import numpy as np

MAGIC = b"LZB1"
_HDR = struct.Struct("<4sIQQ")


def split(total_len, block_size):
    """(offsets, lengths) of the fixed-size blocks of a buffer (the last one may be short)."""
    if total_len == 0:
        return np.zeros(0, dtype=np.uint64), np.zeros(0, dtype=np.uint64)
    n = (total_len + block_size - 1) // block_size
    off = np.arange(n, dtype=np.uint64) * np.uint64(block_size)
    ln = np.full(n, block_size, dtype=np.uint64)
    ln[-1] = total_len - (n - 1) * block_size
    return off, ln


def shard_range(n_blocks, rank, world):
    """Contiguous block range of `rank`."""
    return (rank * n_blocks) // world, ((rank + 1) * n_blocks) // world


def exclusive_scan(sizes):
    sizes = np.asarray(sizes, dtype=np.uint64)
    off = np.zeros(sizes.size, dtype=np.uint64)
    if sizes.size > 1:
        off[1:] = np.cumsum(sizes)[:-1]
    return off


def gather_sizes(local_sizes, n_blocks, rank, world):
    """All ranks learn every block's compressed size: one all_gather of 8 bytes per block."""
    local_sizes = np.asarray(local_sizes, dtype=np.int64)
    if world == 1:
        return local_sizes.astype(np.uint64)
    import torch
    import torch.distributed as dist
    lo, hi = shard_range(n_blocks, rank, world)
    assert local_sizes.size == hi - lo
    width = max(shard_range(n_blocks, r, world)[1] - shard_range(n_blocks, r, world)[0] for r in range(world))
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    mine = torch.zeros(width, dtype=torch.int64, device=dev)
    mine[: local_sizes.size] = torch.from_numpy(local_sizes).to(dev)
    parts = [torch.zeros(width, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(parts, mine)
    out = np.zeros(n_blocks, dtype=np.uint64)
    for r in range(world):
        a, b = shard_range(n_blocks, r, world)
        out[a:b] = parts[r][: b - a].cpu().numpy().astype(np.uint64)
    return out


def pack(streams, block_size, usizes):
    """Container bytes from per-block .lzma streams."""
    n = len(streams)
    total = int(sum(int(u) for u in usizes))
    parts = [_HDR.pack(MAGIC, n, block_size, total)]
    index = np.empty((n, 2), dtype="<u8")
    index[:, 0] = [len(s) for s in streams]
    index[:, 1] = usizes
    parts.append(index.tobytes())
    parts.extend(bytes(s) for s in streams)
    return b"".join(parts)


def unpack(container):
    """-> (block_size, total_size, csize[n], usize[n], offsets[n] into `container`)."""
    buf = memoryview(container)
    if len(buf) < _HDR.size:
        raise ValueError("container too short")
    magic, n, block_size, total = _HDR.unpack_from(buf, 0)
    if magic != MAGIC:
        raise ValueError("bad container magic")
    idx_end = _HDR.size + 16 * n
    if len(buf) < idx_end:
        raise ValueError("truncated container index")
    index = np.frombuffer(buf, dtype="<u8", count=2 * n, offset=_HDR.size).reshape(n, 2)
    csize, usize = index[:, 0].astype(np.uint64), index[:, 1].astype(np.uint64)
    offsets = exclusive_scan(csize) + np.uint64(idx_end)
    if n and int(offsets[-1] + csize[-1]) > len(buf):
        raise ValueError("truncated container payload")
    if int(usize.sum()) != total:
        raise ValueError("container index does not add up")
    return block_size, total, csize, usize, offsets


def encode_buffer(encoder, data, block_size):
    """Compress a buffer block by block on the encoder's GPU -> container bytes."""
    a = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, dtype=np.uint8)
    off, ln = split(a.size, block_size)
    if off.size == 0:
        return pack([], block_size, [])
    out, ooff, olen = encoder.code_batch(a, off, ln, with_header=True)
    streams = [out[int(o): int(o + l)] for o, l in zip(ooff, olen)]
    return pack([s.tobytes() for s in streams], block_size, ln)


def decode_buffer(decoder, container):
    """Inverse of encode_buffer on the decoder's GPU.  Raises ValueError on a corrupt block
    (where the reference's Decoder.Code would have returned false)."""
    block_size, total, csize, usize, offsets = unpack(container)
    if csize.size == 0:
        return b""
    arr = np.frombuffer(bytes(container), dtype=np.uint8)
    cap = usize + np.uint64(273)
    ooff = exclusive_scan(cap)
    out, out_len, status = decoder.code_batch(arr, offsets, csize, ooff, cap)
    if not (status == 1).all() or not np.array_equal(out_len, usize):
        bad = int(np.flatnonzero((status != 1) | (out_len != usize))[0])
        raise ValueError("Error in data stream (block %d, status %d)" % (bad, int(status[bad])))
    return b"".join(out[int(o): int(o + l)].tobytes() for o, l in zip(ooff, usize))
