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


def slice_ranges(n_units, rank, world, n_slices):
    """Slice-major partition for a pipelined gather: the stream is cut into `n_slices` contiguous slices, and inside
    slice k rank r takes the contiguous sub-range unit_range(len(slice k), r, world).  A rank thus owns n_slices
    contiguous ranges instead of one; in return everything before slice k's payloads is known once slice k's sizes are
    all-gathered, so slice k can travel to its final place on GPU 0 while slice k + 1 is still being compressed.
    -> [(lo, hi)] per slice."""
    out = []
    for k in range(n_slices):
        lo, hi = unit_range(n_units, k, n_slices)
        a, b = unit_range(hi - lo, rank, world)
        out.append((lo + a, lo + b))
    return out


def gather_slice(local_payload, local_total, all_totals, stream, base, group=None, dst=0):
    """One exchange step of the pipelined gather: every rank's payload of this slice goes to rank `dst`, in rank order,
    starting at stream[base].  all_totals: the slice's compressed byte count of every rank (host ints, from the size
    all-gather).  All transfers of the step are posted as ONE batch (ncclGroupStart/End underneath), so they run side
    by side instead of one after the other.  -> bytes the slice occupies in the stream."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    offs = [0]
    for t in all_totals:
        offs.append(offs[-1] + int(t))
    if rank == dst:
        if local_total:
            stream[base + offs[rank]:base + offs[rank + 1]].copy_(local_payload[:local_total], non_blocking=True)
        ops = [dist.P2POp(dist.irecv, stream[base + offs[r]:base + offs[r + 1]], r, group)
               for r in range(world) if r != dst and all_totals[r]]
    else:
        ops = [dist.P2POp(dist.isend, local_payload[:local_total], dst, group)] if local_total else []
    if ops:
        for q in dist.batch_isend_irecv(ops):
            q.wait()
    return offs[-1]


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
        ops = [dist.P2POp(dist.irecv, stream[offs[r]:offs[r + 1]], r, group) for r in range(world) if r != dst and totals[r]]
        if ops:                                      # one batch: the transfers run side by side, not serialised
            for q in dist.batch_isend_irecv(ops):
                q.wait()
        return stream, all_sizes
    if totals[rank]:
        for q in dist.batch_isend_irecv([dist.P2POp(dist.isend, local_payload[:totals[rank]].contiguous(), dst, group)]):
            q.wait()
    return None, all_sizes


def combine_crcs(crc_combine, crcs, lens, crc=0):
    """Folds per-unit CRC-32s (in stream order) into the CRC-32 of the concatenation.  crc_combine = b2d_crc32_combine."""
    for c, n in zip(crcs, lens):
        crc = crc_combine(crc, int(c) & 0xFFFFFFFF, int(n))
    return crc
