"""Worker of tests/test_sharding.py: world_size-N run of the host-side sharding logic on the gloo backend (CPU).
Each rank makes the chunks of its shard with zlib (sync-flushed, i.e. byte-aligned with an empty stored block, exactly
the framing b2d_deflate_chunks produces), the product code gathers them onto rank 0, and rank 0 checks the stream.
argv: <result.json> [slices]  -- with a slice count the slice-major partition and the per-slice gather of config 4
(sharding.slice_ranges / gather_slice, bench.py config4_leg) run instead of the one-range gather."""
import json
import os
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

import b2d_loader


def compress_range(data, chunk, n_chunks, lo, hi):
    payload, sizes, crcs, lens = bytearray(), [], [], []
    for c in range(lo, hi):
        piece = data[c * chunk:(c + 1) * chunk]
        z = zlib.compressobj(6, zlib.DEFLATED, -15)
        last = c == n_chunks - 1
        body = z.compress(piece) + (z.flush(zlib.Z_FINISH) if last else z.flush(zlib.Z_SYNC_FLUSH))
        payload += body
        sizes.append(len(body))
        crcs.append(zlib.crc32(piece))
        lens.append(len(piece))
    return payload, sizes, crcs, lens


def main_slices(out_path, n_slices):
    """config 4's exchange: per slice k every rank compresses its sub-range, the slice's sizes are all-gathered, the
    payloads go to their FINAL offsets of the stream on rank 0 (everything in front of slice k is known by then)."""
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    b2d = b2d_loader.load()
    from importlib import import_module
    sharding = import_module("b2deflate.sharding")
    chunk = 1 << 15
    n_chunks = 41                                   # slices and sub-ranges of unequal sizes
    data = b2d.corpus("mixed", 0xDEF1A7E + 4, n_chunks * chunk - 777).tobytes()
    ranges = sharding.slice_ranges(n_chunks, rank, world, n_slices)
    stream = torch.zeros(len(data) + n_chunks * 64 + 4096, dtype=torch.uint8) if rank == 0 else None
    base = 0
    g_sizes, g_crcs, g_lens = [], [], []
    for lo, hi in ranges:
        payload, sizes, crcs, lens = compress_range(data, chunk, n_chunks, lo, hi)
        t_payload = torch.from_numpy(np.frombuffer(bytes(payload), dtype=np.uint8).copy()) if payload else torch.zeros(0, dtype=torch.uint8)
        meta = sharding.all_gather_sizes(torch.tensor(sizes + crcs + lens, dtype=torch.int64))
        totals = []
        for m in meta:                              # rank order inside the slice
            k = m.numel() // 3
            totals.append(int(m[:k].sum()))
            g_sizes += m[:k].tolist(); g_crcs += m[k:2 * k].tolist(); g_lens += m[2 * k:].tolist()
        base += sharding.gather_slice(t_payload, totals[rank], totals, stream, base)
    if rank == 0:
        whole = stream[:base].numpy().tobytes()
        d = zlib.decompressobj(-15)
        ok = d.decompress(whole) == data and d.eof
        # the index that came with the all-gathers is in stream order: chunk c sits at the prefix sum of the sizes
        off = np.concatenate([[0], np.cumsum(g_sizes)])
        per_chunk = all(zlib.decompressobj(-15).decompress(whole[off[c]:off[c + 1]]) == data[c * chunk:(c + 1) * chunk] for c in range(n_chunks))
        crc = sharding.combine_crcs(b2d.crc32_combine, g_crcs, g_lens)
        owned = sorted(c for r in range(world) for lo, hi in sharding.slice_ranges(n_chunks, r, world, n_slices) for c in range(lo, hi))
        res = {"ok": bool(ok), "per_chunk_ok": bool(per_chunk), "crc_ok": crc == zlib.crc32(data), "n_sizes": len(g_sizes),
               "sum_sizes": int(sum(g_sizes)), "stream_len": base, "owned_once": owned == list(range(n_chunks)), "world": world}
        with open(out_path, "w") as f:
            json.dump(res, f)
    dist.barrier()
    dist.destroy_process_group()


def main():
    out_path = sys.argv[1]
    if len(sys.argv) > 2:
        return main_slices(out_path, int(sys.argv[2]))
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    b2d = b2d_loader.load()
    from importlib import import_module
    sharding = import_module("b2deflate.sharding")
    chunk = 1 << 16
    n_chunks = 37                                   # odd on purpose: ranks get 19 and 18
    data = b2d.corpus("mixed", 0xDEF1A7E, n_chunks * chunk - 1234).tobytes()     # ragged last chunk
    lo, hi = sharding.unit_range(n_chunks, rank, world)
    payload, sizes, crcs, lens = bytearray(), [], [], []
    for c in range(lo, hi):
        piece = data[c * chunk:(c + 1) * chunk]
        z = zlib.compressobj(6, zlib.DEFLATED, -15)
        last = c == n_chunks - 1
        body = z.compress(piece) + (z.flush(zlib.Z_FINISH) if last else z.flush(zlib.Z_SYNC_FLUSH))
        payload += body
        sizes.append(len(body))
        crcs.append(zlib.crc32(piece))
        lens.append(len(piece))
    t_payload = torch.from_numpy(np.frombuffer(bytes(payload), dtype=np.uint8).copy())
    t_sizes = torch.tensor(sizes, dtype=torch.int64)
    stream, all_sizes = sharding.gather_stream(t_payload, t_sizes)
    # CRCs and lengths ride an all-gather too
    meta = sharding.all_gather_sizes(torch.tensor([v for pair in zip(crcs, lens) for v in pair], dtype=torch.int64))
    if rank == 0:
        flat = torch.cat(meta).tolist()
        g_crcs, g_lens = flat[0::2], flat[1::2]
        crc = sharding.combine_crcs(b2d.crc32_combine, g_crcs, g_lens)
        whole = stream.numpy().tobytes()
        ok = zlib.decompress(whole, -15) == data
        res = {"ok": bool(ok), "crc_ok": crc == zlib.crc32(data), "n_sizes": int(all_sizes.numel()),
               "sum_sizes": int(all_sizes.sum()), "stream_len": len(whole), "ranges": [sharding.unit_range(n_chunks, r, world) for r in range(world)]}
        with open(out_path, "w") as f:
            json.dump(res, f)
    else:
        assert stream is None
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
