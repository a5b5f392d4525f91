"""Worker of tests/test_sharding.py: world_size-2 run of the host-side sharding logic on the gloo backend (CPU).
Each rank makes the chunks of its shard with zlib (sync-flushed, i.e. byte-aligned with an empty stored block, exactly
the framing b2d_deflate_chunks produces), the product code gathers them onto rank 0, and rank 0 checks the stream."""
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


def main():
    out_path = sys.argv[1]
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
