#!/usr/bin/env python3
"""bench.py -- the BASELINE.json metric ("deflate/inflate GB/s uncompressed ... % HBM peak") on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b2d|reference] [--size-mib M]

Workload (config.workload): BASELINE.json configs[1] -- batch inflate of 4096 independent 256 KiB gzip-style members
(1 GiB uncompressed per GPU; raw DEFLATE bodies made by zlib level 6 from the G_TEXT corpus, CRC-32 of every
member's output computed as GzipInputStream does) is the headline `value`; configs[2] -- chunked dynamic-Huffman
deflate of 1 GiB of G_MIXED data in 1 MiB chunks with sync-flush markers + CRC-32 -- is reported in the same JSON
line under "deflate".  One "step" = one pass over the whole batch.  Multi-GPU (torchrun, one rank per GPU): members /
chunks are sharded by rank with no data-path collective (weak scaling: every rank decodes its own 1 GiB); the
deflate leg gathers compressed sizes (all_gather) and payloads (NCCL send/recv over NVLink) onto GPU 0.

`value`   : device-resident throughput (inputs and outputs in HBM), CUDA events on the launching stream.
`e2e`     : the same work through the host-pointer C-ABI call (b2d_inflate_batch / b2d_deflate_chunks) with pinned
            host buffers; H2D + kernels + D2H inside the timed region (the call blocks until the result is on the host).
`roofline`: inflate kernel alone: algorithmic bytes (compressed read + uncompressed written) / CUDA-event time,
            against MEASURED_PEAKS.json hbm_gbs.
`cpu_baseline`: the oracle (C restatement of the reference's Open.java / Lz77Huffman.java loops -- no JVM exists in
            this image, so the reference itself cannot run) on the host cores, bounded sample, rank 0.
--impl reference: times that oracle with all host threads on the same workload (bounded sample per step).
"""
import argparse
import concurrent.futures as cf
import ctypes
import json
import os
import struct
import subprocess
import sys
import threading
import time
import zlib

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SEED = 0xDEF1A7E
GZ_HEADER = bytes([0x1F, 0x8B, 8, 0, 0, 0, 0, 0, 0, 3])
MEMBER_BYTES = 256 * 1024
CHUNK_BYTES = 1 << 20


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b2d", choices=["b2d", "reference"])
    ap.add_argument("--size-mib", type=int, default=1024, help="uncompressed MiB per GPU (1024 = BASELINE configs)")
    ap.add_argument("--members", default="zlib", choices=["zlib", "gpu"], help="who encodes the inflate inputs")
    ap.add_argument("--no-deflate", action="store_true", help="skip the deflate leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def measured_traffic(kernel, n_members):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/traffic.json); only valid
    for the workload it was captured on (the default 4096-member batch)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if n_members != 4096 or not os.path.exists(p):
        return None
    with open(p) as f:
        return json.load(f).get(kernel, {}).get("dram_bytes_per_launch")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------ data
def make_members(b2d, n_members, seed0, pool):
    """-> (list of gzip members, np.uint8 uncompressed blob).  Body: zlib level 6, raw (wbits -15); 10-byte header
    (no optional fields, OS = Unix) and CRC-32 + ISIZE trailer as GzipOutputStream.java:62-70 writes them."""
    raw = np.empty(n_members * MEMBER_BYTES, dtype=np.uint8)
    L = b2d.lib()

    def one(i):
        view = raw[i * MEMBER_BYTES:(i + 1) * MEMBER_BYTES]
        L.b2d_corpus_text(seed0 + i, view.ctypes.data, MEMBER_BYTES)
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = c.compress(view.data) + c.flush()
        return GZ_HEADER + body + struct.pack("<II", zlib.crc32(view.data), MEMBER_BYTES)

    return list(pool.map(one, range(n_members))), raw


def make_mixed(b2d, n_bytes, seed, pool):
    """G_MIXED in independent 16 MiB pieces (seed + piece index) so it generates on all cores."""
    out = np.empty(n_bytes, dtype=np.uint8)
    L = b2d.lib()
    piece = 16 << 20

    def one(k):
        a = k * piece
        b = min(n_bytes, a + piece)
        L.b2d_corpus_mixed(seed + k, out[a:b].ctypes.data, b - a)

    list(pool.map(one, range((n_bytes + piece - 1) // piece)))
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, ln in self.lines:
            if t0 is not None and not (t0 - 0.05 <= t <= t1 + 0.15):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------ CPU legs (oracle = restated reference)
def cpu_inflate(O, members, threads):
    """Oracle inflate over `members` with `threads` host threads (ctypes releases the GIL).  -> seconds."""
    L = O.lib()
    outs = [ctypes.create_string_buffer(MEMBER_BYTES) for _ in range(threads)]
    bodies = [m[len(GZ_HEADER):-8] for m in members]
    crcs = [struct.unpack("<I", m[-8:-4])[0] for m in members]

    def work(k):
        ol, ic = ctypes.c_size_t(0), ctypes.c_size_t(0)
        for i in range(k, len(members), threads):
            # GzipInputStream.java:66-90: inflate the body, CRC-32 of the output against the trailer.  The reference takes
            # its CRC from the JDK (java.util.zip.CRC32, an intrinsic); zlib's fast crc32 stands in for it here rather
            # than the oracle's bit-serial one, so the CPU arm is not handicapped.
            st = L.oracle_inflate(bodies[i], len(bodies[i]), outs[k], MEMBER_BYTES, ctypes.byref(ol), ctypes.byref(ic))
            assert st == 0 and ol.value == MEMBER_BYTES
            assert zlib.crc32(memoryview(outs[k])) == crcs[i]
    t = time.perf_counter()
    with cf.ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, range(threads)))
    return time.perf_counter() - t


def cpu_deflate(O, data, n_chunks, threads, strategy):
    """Oracle DeflaterOutputStream (64 KiB blocks) over independent 1 MiB chunks.  -> (seconds, compressed bytes)."""
    L = O.lib()
    cap = L.oracle_deflate_bound(CHUNK_BYTES, 65536)
    outs = [ctypes.create_string_buffer(cap) for _ in range(threads)]
    strat = (ctypes.c_int * 1)(strategy)
    sizes = [0] * n_chunks
    base = data.ctypes.data

    def work(k):
        for c in range(k, n_chunks, threads):
            sizes[c] = L.oracle_deflate(ctypes.c_char_p(base + c * CHUNK_BYTES), CHUNK_BYTES, strat, 1, 65536, 32768, 0,
                                        outs[k], cap)
    t = time.perf_counter()
    with cf.ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, range(threads)))
    return time.perf_counter() - t, sum(sizes)


# ------------------------------------------------------------------ main
def emit(line):
    """The ONE JSON line goes to the real stdout; everything else a library prints there (NCCL's version banner,
    for one) was redirected to stderr in main()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    args = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = host_threads()
    n_members = args.size_mib * (1 << 20) // MEMBER_BYTES
    n_chunks = args.size_mib
    import b2d_loader
    b2d = b2d_loader.load()

    if args.impl == "reference":
        if rank != 0:
            return 0
        return run_reference(args, b2d, cores, n_members)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b2d arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    b2d.init(local_rank)
    L = b2d.lib()
    threads = max(1, cores // world)
    pool = cf.ThreadPoolExecutor(threads)
    hbm_peak, peak_src = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---------------- inflate inputs (this rank's shard: members [rank*n, (rank+1)*n)) ----------------
    t_prep = time.perf_counter()
    if args.members == "zlib":
        members, raw = make_members(b2d, n_members, SEED + rank * n_members, pool)
    else:
        raw = np.empty(n_members * MEMBER_BYTES, dtype=np.uint8)
        list(pool.map(lambda i: L.b2d_corpus_text(SEED + rank * n_members + i, raw[i * MEMBER_BYTES:].ctypes.data,
                                                  MEMBER_BYTES), range(n_members)))
        members = []
        for i in range(n_members):   # one complete stream per member (reference framing), made by the GPU encoder
            view = raw[i * MEMBER_BYTES:(i + 1) * MEMBER_BYTES]
            body = bytes(b2d.deflate_chunks(view, b2d.make_opts(framing=b2d.FRAMING_REFERENCE, mode=b2d.MODE_DYNAMIC)))
            members.append(GZ_HEADER + body + struct.pack("<II", zlib.crc32(view.data), MEMBER_BYTES))
    mem_off = np.zeros(n_members + 1, dtype=np.int64)             # gzip members back to back
    mem_off[1:] = np.cumsum([len(m) for m in members])
    comp_total = int(mem_off[-1])
    in_off = mem_off.copy()                                        # DEFLATE bodies for the device-resident call: member i's
    in_off[:-1] += len(GZ_HEADER)                                  # range runs from its body to the next body (the decoder
    body_len = np.array([len(m) - len(GZ_HEADER) - 8 for m in members], dtype=np.int64)   # stops at BFINAL)
    trailer_crc = np.array([struct.unpack("<I", m[-8:-4])[0] for m in members], dtype=np.uint32)
    out_off = np.arange(n_members + 1, dtype=np.int64) * MEMBER_BYTES
    out_total = n_members * MEMBER_BYTES
    h_blob = torch.empty(comp_total + 64, dtype=torch.uint8).pin_memory()
    h_blob[:comp_total] = torch.from_numpy(np.frombuffer(b"".join(members), dtype=np.uint8).copy())
    h_out = torch.empty(out_total, dtype=torch.uint8).pin_memory()
    prep_s = time.perf_counter() - t_prep

    d_blob = h_blob.to(dev)
    d_in_off = torch.from_numpy(in_off).to(dev)
    d_out_off = torch.from_numpy(out_off).to(dev)
    d_out = torch.zeros(out_total, dtype=torch.uint8, device=dev)
    d_out_len = torch.zeros(n_members, dtype=torch.int64, device=dev)
    d_cons = torch.zeros(n_members, dtype=torch.int64, device=dev)
    d_crc = torch.zeros(n_members, dtype=torch.int32, device=dev)
    d_status = torch.zeros(n_members, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    sp = ctypes.c_void_p(stream.cuda_stream)

    def inflate_dev(flags):
        r = L.b2d_inflate_batch_dev(d_blob.data_ptr(), d_in_off.data_ptr(), n_members, d_out.data_ptr(),
                                    d_out_off.data_ptr(), d_out_len.data_ptr(), d_cons.data_ptr(), d_crc.data_ptr(),
                                    d_status.data_ptr(), flags, sp)
        if r != 0:
            raise RuntimeError(f"b2d_inflate_batch_dev: {b2d.status_name(r)}")

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1) / 1e3)

    # warm-up + correctness at full size (every byte, every member)
    for _ in range(max(3, args.warmup)):
        inflate_dev(b2d.INFLATE_CRC32)
    torch.cuda.synchronize()
    assert int(d_status.abs().sum().item()) == 0, "inflate: a member failed"
    assert bool((d_out_len == MEMBER_BYTES).all().item())
    assert bool((d_cons.cpu() == torch.from_numpy(body_len)).all().item()), "inflate: consumed != body length"
    d_raw = torch.from_numpy(raw).to(dev)
    assert torch.equal(d_out, d_raw), "inflate: output differs from the original data"
    crc_host = d_crc.cpu().numpy().view(np.uint32)
    assert np.array_equal(crc_host, trailer_crc), "inflate: CRC-32 differs from the gzip trailers"
    del d_raw

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    tc0 = time.perf_counter()
    # the headline step: inflate kernel + CRC-32 of every member's output
    step_s = timed(lambda: inflate_dev(b2d.INFLATE_CRC32), args.steps) / args.steps
    # the dominant kernel alone (roofline)
    kern_s = timed(lambda: inflate_dev(0), args.steps) / args.steps
    tc1 = time.perf_counter()
    clocks = sampler.stop(tc0, tc1)
    total_uncomp = sum_over_ranks(float(out_total))
    value = total_uncomp / step_s / 1e9
    algo_bytes = comp_total + out_total
    achieved = algo_bytes / kern_s / 1e9

    # e2e: host pointers through b2d_inflate_batch (H2D of the compressed blob, kernels, D2H of the output)
    h_in_off = mem_off.astype(np.uint64)
    h_out_off = out_off.astype(np.uint64)
    h_len = np.zeros(n_members, np.uint64); h_cons = np.zeros(n_members, np.uint64)
    h_st = np.zeros(n_members, np.int32)

    def inflate_host():                                  # whole gzip members: header + body + trailer checks
        r = L.b2d_gunzip_batch(h_blob.data_ptr(), h_in_off.ctypes.data, n_members, h_out.data_ptr(), h_out_off.ctypes.data,
                               h_len.ctypes.data, h_cons.ctypes.data, h_st.ctypes.data)
        if r != 0:
            raise RuntimeError(f"b2d_gunzip_batch: {b2d.status_name(r)}")

    for _ in range(2):
        inflate_host()
    assert not h_st.any() and np.array_equal(h_out.numpy()[:1 << 24], raw[:1 << 24])
    assert np.array_equal(h_cons.astype(np.int64), mem_off[1:] - mem_off[:-1])
    e2e_steps = max(3, args.steps // 2)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        inflate_host()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    e2e_val = total_uncomp / e2e_s / 1e9
    h2d = comp_total + 2 * (n_members + 1) * 8
    d2h = out_total + n_members * (8 + 8 + 4 + 4)

    line = {
        "metric": "batch inflate GB/s uncompressed (BASELINE: deflate/inflate GB/s uncompressed; deflate leg in \"deflate\")",
        "value": round(value, 3), "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": round(step_s * 1e3, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"BASELINE configs[1]: batch inflate of {n_members} independent 256 KiB members per GPU "
                               f"({args.size_mib} MiB uncompressed per GPU), gzip members, bodies by "
                               f"{'zlib level 6' if args.members == 'zlib' else 'the GPU encoder'} over G_TEXT, CRC-32 per member checked against the trailer",
                   "members_per_gpu": n_members, "member_bytes": MEMBER_BYTES,
                   "compressed_bytes_per_gpu": comp_total, "ratio": round(out_total / comp_total, 4),
                   "sharding": f"members by rank, no collective ({world} rank(s))",
                   "l2": "inputs larger than L2 (compressed blob + 1 GiB output > 126 MB), no flush needed",
                   "e2e_call": "b2d_gunzip_batch (host pointers, pinned): header checks, H2D, inflate + CRC-32 kernels, D2H, trailer checks",
                   "e2e_timer": "host clock around the blocking C-ABI call (ends with a stream sync), max over ranks",
                   "prep_s": round(prep_s, 1)},
        "e2e": {"value": round(e2e_val, 3), "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": round(e2e_s * 1e3, 3)},
        # our kernels launched inside the timed regions of the inflate leg: inflate_kernel + crc32_kernel per headline
        # step, inflate_kernel per roofline step, and per e2e step one (inflate_kernel, crc32_kernel) pair per slice
        # (<= 4 slices of >= 1024 members); the deflate leg adds its own below
        "gpu_launches": 2 * args.steps + args.steps + 2 * min(4, max(1, n_members // 1024)) * e2e_steps,
        "roofline": {"bound": "hbm", "kernel": "b2d::inflate_kernel", "achieved": round(achieved, 2), "peak": hbm_peak,
                     "unit": "GB/s", "frac": round(achieved / hbm_peak, 5),
                     "traffic": measured_traffic("inflate_kernel", n_members), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": round(kern_s * 1e3, 4),
                     "note": "latency/issue-bound by construction (serial Huffman decode per member); see DESIGN.md"},
        "clocks": clocks,
    }

    # ---------------- deflate leg (configs[2]) ----------------
    if not args.no_deflate:
        line["deflate"] = deflate_leg(args, b2d, L, torch, dist, dev, rank, world, pool, timed, barrier, max_over_ranks,
                                      sum_over_ranks, stream, sp, hbm_peak, n_chunks)

    # ---------------- CPU baseline (rank 0, N = 1) ----------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        O.build()
        n_st = min(n_members, 512)
        st_s = cpu_inflate(O, members[:n_st], 1)
        mt_s = cpu_inflate(O, members, cores)
        line["cpu_baseline"] = {
            "value": round(out_total / mt_s / 1e9, 4), "unit": "GB/s", "cores": cores, "kind": "port",
            "single_thread": round(n_st * MEMBER_BYTES / st_s / 1e9, 4),
            "sample": f"oracle_inflate (C restatement of decomp/Open.java; no JVM in this image) + zlib crc32 vs the gzip trailer over all {n_members} members "
                      f"with {cores} threads; single_thread over the first {n_st} members"}
        if "deflate" in line:
            line["deflate"]["cpu_baseline"] = line["deflate"].pop("_cpu")(O, cores)
    elif "deflate" in line:
        line["deflate"].pop("_cpu", None)
    if "deflate" in line:
        line["gpu_launches"] += line["deflate"]["gpu_launches"]
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    b2d.shutdown()
    return 0


def deflate_leg(args, b2d, L, torch, dist, dev, rank, world, pool, timed, barrier, max_over_ranks, sum_over_ranks,
                stream, sp, hbm_peak, n_chunks):
    n_bytes = n_chunks * CHUNK_BYTES
    data = make_mixed(b2d, n_bytes, SEED + 1000 * rank, pool)
    h_in = torch.from_numpy(data).pin_memory()
    d_in = h_in.to(dev)
    bound = b2d.deflate_bound(n_bytes, CHUNK_BYTES)
    d_out = torch.empty(bound, dtype=torch.uint8, device=dev)
    d_total = torch.zeros(1, dtype=torch.int64, device=dev)
    d_clen = torch.zeros(n_chunks, dtype=torch.int64, device=dev)
    d_ccrc = torch.zeros(n_chunks, dtype=torch.int32, device=dev)
    opts = b2d.make_opts(chunk_bytes=CHUNK_BYTES, block_bytes=65536, mode=b2d.MODE_AUTO, is_last=int(rank == world - 1))

    d_bits = torch.zeros(n_bytes // 65536, dtype=torch.int32, device=dev)     # restart index: bit offset of every block

    def deflate_dev():
        r = L.b2d_deflate_chunks_indexed_dev(d_in.data_ptr(), n_bytes, ctypes.byref(opts), d_out.data_ptr(), bound,
                                             d_total.data_ptr(), d_clen.data_ptr(), d_ccrc.data_ptr(), d_bits.data_ptr(), sp)
        if r != 0:
            raise RuntimeError(f"b2d_deflate_chunks_indexed_dev: {b2d.status_name(r)}")

    for _ in range(3):
        deflate_dev()
    torch.cuda.synchronize()
    comp_len = int(d_total.item())
    # correctness at full size: decode the stream on the GPU (chunk-indexed) and compare; zlib-decode a sample of chunks
    clen = d_clen.cpu().numpy()
    coff = np.zeros(n_chunks + 1, dtype=np.int64); coff[1:] = np.cumsum(clen)
    assert int(coff[-1]) == comp_len
    d_coff = torch.from_numpy(coff).to(dev)
    d_ooff = (torch.arange(n_chunks + 1, dtype=torch.int64, device=dev) * CHUNK_BYTES)
    d_dec = torch.zeros(n_bytes, dtype=torch.uint8, device=dev)
    d_ol = torch.zeros(n_chunks, dtype=torch.int64, device=dev); d_ic = torch.zeros_like(d_ol)
    d_st = torch.zeros(n_chunks, dtype=torch.int32, device=dev); d_c2 = torch.zeros_like(d_st)
    def inflate_chunks():
        r = L.b2d_inflate_batch_dev(d_out.data_ptr(), d_coff.data_ptr(), n_chunks, d_dec.data_ptr(), d_ooff.data_ptr(),
                                    d_ol.data_ptr(), d_ic.data_ptr(), d_c2.data_ptr(), d_st.data_ptr(),
                                    b2d.INFLATE_CHUNK_INDEXED | b2d.INFLATE_CRC32, sp)
        assert r == 0
    for _ in range(3):
        inflate_chunks()
    torch.cuda.synchronize()
    # configs[3]'s decompress direction: the chunk-indexed stream decoded with one warp per 1 MiB chunk
    unchunk_s = timed(inflate_chunks, max(3, args.steps // 2)) / max(3, args.steps // 2)
    # ... and with the block index: one warp per 64 KiB block, references replayed per chunk afterwards
    d_dec.zero_()
    d_cst = torch.zeros(n_chunks, dtype=torch.int32, device=dev)

    def inflate_blocks():
        r = L.b2d_inflate_chunks_dev(d_out.data_ptr(), d_coff.data_ptr(), n_chunks, d_bits.data_ptr(), CHUNK_BYTES, 65536,
                                     n_bytes, d_dec.data_ptr(), d_c2.data_ptr(), d_cst.data_ptr(), b2d.INFLATE_CRC32, sp)
        assert r == 0
    for _ in range(3):
        inflate_blocks()
    torch.cuda.synchronize()
    assert int(d_cst.abs().sum().item()) == 0 and torch.equal(d_dec, d_in) and torch.equal(d_c2, d_ccrc), "block-indexed decode differs"
    unblock_s = timed(inflate_blocks, max(3, args.steps // 2)) / max(3, args.steps // 2)
    assert int(d_st.abs().sum().item()) == 0 and torch.equal(d_dec, d_in), "deflate: GPU round trip differs"
    assert torch.equal(d_c2, d_ccrc), "deflate: chunk CRCs differ from the CRCs of the decoded chunks"
    h_comp = d_out[:comp_len].cpu().numpy()
    for c in range(0, n_chunks, max(1, n_chunks // 16)):
        d = zlib.decompressobj(-15)
        got = d.decompress(h_comp[coff[c]:coff[c + 1]].tobytes())
        assert got == data[c * CHUNK_BYTES:(c + 1) * CHUNK_BYTES].tobytes(), f"deflate: zlib decode of chunk {c} differs"
    del d_dec

    step_s = timed(deflate_dev, args.steps) / args.steps
    total_in = sum_over_ranks(float(n_bytes))
    value = total_in / step_s / 1e9

    # the same input with adaptive block splitting (pieces of 16 KiB, the GPU counterpart of BinarySplit)
    opts_split = b2d.make_opts(chunk_bytes=CHUNK_BYTES, block_bytes=65536, mode=b2d.MODE_AUTO, is_last=int(rank == world - 1),
                               split_min_bytes=16384)
    d_total_s = torch.zeros(1, dtype=torch.int64, device=dev)

    def deflate_split_dev():
        r = L.b2d_deflate_chunks_dev(d_in.data_ptr(), n_bytes, ctypes.byref(opts_split), d_out.data_ptr(), bound,
                                     d_total_s.data_ptr(), d_clen.data_ptr(), d_ccrc.data_ptr(), sp)
        if r != 0:
            raise RuntimeError(f"b2d_deflate_chunks_dev (split): {b2d.status_name(r)}")
    split_steps = max(3, args.steps // 2)
    for _ in range(2):
        deflate_split_dev()
    split_s = timed(deflate_split_dev, split_steps) / split_steps
    split_len = int(d_total_s.item())
    clen_s = d_clen.cpu().numpy()
    h_split = d_out[:split_len].cpu().numpy()
    o = 0
    for c in range(n_chunks):                                # zlib-decode a sample of the split stream's chunks
        if c % max(1, n_chunks // 16) == 0:
            got = zlib.decompressobj(-15).decompress(h_split[o:o + int(clen_s[c])].tobytes())
            assert got == data[c * CHUNK_BYTES:(c + 1) * CHUNK_BYTES].tobytes(), f"deflate (split): zlib decode of chunk {c} differs"
        o += int(clen_s[c])
    assert o == split_len
    del h_split
    deflate_dev()                                            # d_out / d_clen hold the unsplit stream again for what follows
    torch.cuda.synchronize()

    # multi-GPU: the one exchange step of the path -- chunk sizes all-gathered, payloads sent to GPU 0 (NCCL over NVLink)
    gather_ms = None
    if world > 1:
        from importlib import import_module
        sharding = import_module("b2deflate.sharding")
        for _ in range(2):                                   # warm-up (NCCL connection set-up), then timed
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(stream)
            whole, all_sizes = sharding.gather_stream(d_out[:comp_len], d_clen)
            g1.record(stream)
            barrier()
            gather_ms = max_over_ranks(g0.elapsed_time(g1))
        if rank == 0:
            assert whole.numel() == int(all_sizes.sum().item()) and torch.equal(whole[:comp_len], d_out[:comp_len])
        del whole

    # e2e through the host-pointer call
    h_out = torch.empty(bound, dtype=torch.uint8).pin_memory()
    idx = np.zeros(n_chunks, np.uint64)

    def deflate_host():
        c = ctypes.c_uint32(0)
        r = L.b2d_deflate_chunks(h_in.data_ptr(), n_bytes, ctypes.byref(opts), h_out.data_ptr(), bound, ctypes.byref(c),
                                 idx.ctypes.data)
        if r < 0:
            raise RuntimeError(f"b2d_deflate_chunks: {b2d.status_name(int(r))}")
        return int(r), c.value

    for _ in range(2):
        n_out, crc = deflate_host()
    assert n_out == comp_len and np.array_equal(h_out.numpy()[:n_out], h_comp)
    if n_bytes <= (1 << 30):
        assert crc == zlib.crc32(data.data)
    e2e_steps = max(3, args.steps // 2)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        deflate_host()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)

    res = {
        "workload": f"BASELINE configs[2]: chunked dynamic-Huffman deflate of {n_chunks} MiB G_MIXED per GPU, 1 MiB chunks "
                    "+ sync-flush markers, 64 KiB blocks, mode auto, CRC-32 per chunk",
        "value": round(value, 3), "unit": "GB/s", "ms_per_step": round(step_s * 1e3, 3),
        "compressed_bytes_per_gpu": comp_len, "ratio": round(n_bytes / comp_len, 4),
        "e2e": {"value": round(total_in / e2e_s / 1e9, 3), "unit": "GB/s", "h2d_bytes_per_step": n_bytes,
                "d2h_bytes_per_step": comp_len + n_chunks * 12 + 8, "ms_per_step": round(e2e_s * 1e3, 3)},
        # chains, match, parse, node_hist, huffman, split_decide, layout, scan, emit, block_bits, crc32 (+ one
        # cudaMemsetAsync, not ours) per call
        "gpu_launches_per_step": 11,
        # deflate steps (+ block_bits_kernel), chunk-indexed decode (inflate + crc32), block-indexed decode (units + resolve + crc32), e2e
        "gpu_launches": 11 * args.steps + (2 + 3) * max(3, args.steps // 2) + 10 * max(1, min(4, n_chunks // 256)) * e2e_steps
                        + 10 * (split_steps + 2) + 11,
        "roofline": {"bound": "hbm", "achieved": round((n_bytes + comp_len) / step_s / 1e9, 2), "peak": hbm_peak,
                     "unit": "GB/s", "frac": round((n_bytes + comp_len) / step_s / 1e9 / hbm_peak, 5),
                     "note": "whole pipeline (10 kernels + memset); algorithmic bytes = input read + compressed written"},
        "gather_to_gpu0_ms": gather_ms,
        "inflate_chunk_indexed": {"value": round(sum_over_ranks(float(n_bytes)) / unchunk_s / 1e9, 3), "unit": "GB/s",
                                  "ms_per_step": round(unchunk_s * 1e3, 3),
                                  "note": f"decode of this stream, one warp per 1 MiB chunk ({n_chunks} units per GPU: latency-bound below ~4096 units)"},
        "inflate_block_indexed": {"value": round(sum_over_ranks(float(n_bytes)) / unblock_s / 1e9, 3), "unit": "GB/s",
                                  "ms_per_step": round(unblock_s * 1e3, 3),
                                  "note": "b2d_inflate_chunks_dev: one warp per 64 KiB block (Huffman decode), then one warp per chunk replays the back-references"},
        "adaptive_split": {"value": round(sum_over_ranks(float(n_bytes)) / split_s / 1e9, 3), "unit": "GB/s",
                           "ms_per_step": round(split_s * 1e3, 3), "compressed_bytes_per_gpu": split_len,
                           "ratio": round(n_bytes / split_len, 4), "bytes_vs_unsplit": round(split_len / comp_len, 5),
                           "note": "split_min_bytes = 16 KiB: every 64 KiB span becomes the cheapest partition of its tree of "
                                   "pieces (SURVEY 8f N3, comp/BinarySplit.java); 10 launches per call"},
    }

    def cpu(O, cores):
        n_s = min(n_chunks, 8 * cores)
        st_s, _ = cpu_deflate(O, data, min(n_chunks, 16), 1, O.RLE_DYNAMIC)
        mt_s, rle_bytes = cpu_deflate(O, data, n_s, cores, O.RLE_DYNAMIC)
        n_f = min(n_chunks, 2 * cores)
        f_s, full_bytes = cpu_deflate(O, data, n_f, cores, O.FULL_DYNAMIC)
        gpu_same = int(clen[:n_f].sum())
        return {"value": round(n_s * CHUNK_BYTES / mt_s / 1e9, 4), "unit": "GB/s", "cores": cores, "kind": "port",
                "single_thread": round(min(n_chunks, 16) * CHUNK_BYTES / st_s / 1e9, 4),
                "sample": f"oracle_deflate RLE_DYNAMIC (the DeflaterOutputStream default, 64 KiB blocks) over the first {n_s} "
                          f"chunks with {cores} threads; FULL_DYNAMIC over the first {n_f} chunks for the ratio bar",
                "full_dynamic_GBps": round(n_f * CHUNK_BYTES / f_s / 1e9, 5),
                "ratio_rle_dynamic": round(n_s * CHUNK_BYTES / rle_bytes, 4),
                "ratio_full_dynamic": round(n_f * CHUNK_BYTES / (full_bytes + 5 * n_f), 4),
                "gpu_bytes_vs_full_dynamic_same_chunks": round(gpu_same / (full_bytes + 5 * n_f), 5)}
    res["_cpu"] = cpu
    return res


def run_reference(args, b2d, cores, n_members):
    """The reference's own CPU path on the host cores.  No JVM exists in this image, so the
    reference's Java cannot run; the arm is the oracle port (C restatement of decomp/Open.java), all host threads,
    each step a bounded sample of the b2d arm's workload."""
    from oracle import oracle as O
    O.build()
    n_s = min(n_members, 64 * cores)
    pool = cf.ThreadPoolExecutor(cores)
    members, raw = make_members(b2d, n_s, SEED, pool)
    for _ in range(min(args.warmup, 1)):
        cpu_inflate(O, members, cores)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_inflate(O, members, cores)
    step_s = t / args.steps
    val = n_s * MEMBER_BYTES / step_s / 1e9
    sample = (f"oracle_inflate (C restatement of decomp/Open.java; JVM absent) + zlib crc32 vs the gzip trailer over the first {n_s} of {n_members} "
              f"members per step, {cores} threads")
    emit({
        "impl": "reference", "metric": "batch inflate GB/s uncompressed", "value": round(val, 4), "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(step_s * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"BASELINE configs[1]: batch inflate of independent 256 KiB gzip members (zlib level 6 over G_TEXT), GzipInputStream semantics; "
                               f"bounded sample of {n_s} members per step"},
        "cpu_baseline": {"value": round(val, 4), "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(val, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0})
    return 0


if __name__ == "__main__":
    sys.exit(main())
