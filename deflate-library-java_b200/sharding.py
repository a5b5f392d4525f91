"""Host-side partitioning of chunks / members across the GPUs of one box, and the one exchange step of the path.

The codec's units are independent (gzip members; 1 MiB chunks whose history is reset and which end byte-aligned in an
empty stored block), so rank r of N simply takes the contiguous unit range [r*U/N, (r+1)*U/N) -- no data-path collective.
Only the compress direction has an exchange: the per-chunk compressed sizes are all-gathered (every rank can then
compute the global exclusive scan) and the payloads are sent to rank 0, which writes them at the scanned offsets --
NCCL send/recv over NVLink on GPU tensors, gloo on CPU tensors in the tests.  Per-chunk CRC-32s ride the same
all-gather and are folded with crc32_combine (x^(8*len) mod P) on the host.

(The reference is single-threaded and has no counterpart: DeflaterOutputStream.java:119-137 writes sequentially.)
"""
import torch
import torch.distributed as dist


def unit_range(n_units, rank, world):
    """Contiguous range of units for `rank`: [lo, hi).  Sizes differ by at most one."""
    base, rem = divmod(n_units, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_sizes(local_sizes, group=None):
    """local_sizes: 1-D int64 tensor (per-chunk compressed sizes of this rank; lengths may differ per rank).
    -> list of 1-D int64 CPU tensors, one per rank."""
    world = dist.get_world_size(group)
    dev = local_sizes.device
    n_local = torch.tensor([local_sizes.numel()], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    width = max(counts + [1])
    padded = torch.zeros(width, dtype=torch.int64, device=dev)
    padded[:local_sizes.numel()] = local_sizes
    gathered = [torch.zeros(width, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(gathered, padded, group=group)
    return [g[:c].cpu() for g, c in zip(gathered, counts)]


def gather_stream(local_payload, local_sizes, group=None, dst=0):
    """Gathers the compressed chunks of every rank onto rank `dst`, in rank order.

    local_payload: 1-D uint8 tensor (this rank's chunks back to back), local_sizes: 1-D int64 tensor (their sizes).
    -> on dst: (stream uint8 tensor on local_payload's device, global chunk sizes int64 CPU tensor); elsewhere (None, sizes)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    sizes = all_gather_sizes(local_sizes, group)
    totals = [int(s.sum().item()) for s in sizes]
    all_sizes = torch.cat(sizes) if sizes else torch.zeros(0, dtype=torch.int64)
    if world == 1:
        return local_payload[:totals[0]], all_sizes
    if rank == dst:
        stream = torch.empty(sum(totals), dtype=torch.uint8, device=local_payload.device)
        offs = [0]
        for t in totals:
            offs.append(offs[-1] + t)
        stream[offs[rank]:offs[rank + 1]] = local_payload[:totals[rank]]
        reqs = [dist.irecv(stream[offs[r]:offs[r + 1]], src=r, group=group) for r in range(world) if r != dst and totals[r]]
        for q in reqs:
            q.wait()
        return stream, all_sizes
    if totals[rank]:
        dist.send(local_payload[:totals[rank]].contiguous(), dst=dst, group=group)
    return None, all_sizes


def combine_crcs(crc_combine, crcs, lens, crc=0):
    """Folds per-unit CRC-32s (in stream order) into the CRC-32 of the concatenation.  crc_combine = b2d_crc32_combine."""
    for c, n in zip(crcs, lens):
        crc = crc_combine(crc, int(c) & 0xFFFFFFFF, int(n))
    return crc
