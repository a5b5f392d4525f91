#!/usr/bin/env python3
"""bench.py -- the BASELINE.json metric ("deflate/inflate GB/s uncompressed ... % HBM peak") on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b2d|reference] [--size-mib M]
                    [--config4-gib G] [--edge-gib E] [--no-deflate] [--no-config4] [--no-edge] [--no-cpu-baseline]

One JSON line.  Both arms print the SAME `metric` ("batch inflate GB/s uncompressed") and `config.workload`.

Legs of the b2d arm (one "step" = one pass over the whole batch):
  headline  BASELINE configs[1]: batch inflate of 4096 independent 256 KiB gzip members per GPU (1 GiB uncompressed per
            GPU; bodies by zlib level 6 over G_TEXT; CRC-32 per member as GzipInputStream computes it).  Weak scaling:
            members are sharded by rank, no collective.
              value    device-resident (inputs and outputs in HBM), CUDA events on the launching stream
              e2e      b2d_gunzip_batch with pinned HOST buffers: H2D + kernels + D2H inside the timed region, next
                       to e2e.ceiling_gbs = what the box's PCIe moves for the same bytes (measured in this run)
              roofline inflate_kernel alone: (compressed read + uncompressed written) / CUDA-event time vs the measured
                       HBM peak
  deflate   configs[2]: chunked dynamic-Huffman deflate of 1 GiB G_MIXED per GPU (1 MiB chunks, sync-flush markers,
            CRC-32), device + e2e, plus the decode of that stream (chunk-indexed / block-indexed) and adaptive splitting
  config4   configs[3] as SURVEY 8(d) wrote it: ONE 8 GiB G_MIXED input = 8192 chunks shared by the N ranks (strong
            scaling).  Timed step = compress -> size all-gather -> payloads gathered onto GPU 0 at their final offsets
            (NCCL, batched, pipelined per slice under the next slice's kernels), then the same stream decoded with the
            same sharding.  Rank 0 verifies the GATHERED stream: whole-stream GPU decode vs every rank's input CRCs and a
            strided zlib sample.
  config5   configs[4]: 4 GiB per case shared by the N ranks -- random bytes (stored blocks), zeros (length 258 /
            distance 1 runs), fixed-Huffman-only text -- compress, decode both ways, byte-exact.
  cpu_baseline (rank 0, N = 1): the oracle (C restatement of Open.java / Lz77Huffman.java; no JVM in this image) on the
            host cores, plus system zlib as the java.util.zip context.
--impl reference: times that oracle with all host threads on the same workload, rank 0 only; it loads oracle/ only.
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
# libb2deflate's host pipelines run a dozen streams side by side; CUDA's default of 8 hardware queues would serialise
# streams that share one.  Read at context creation, so it has to be in the environment before torch touches CUDA.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "16")

import numpy as np  # noqa: E402

SEED = 0xDEF1A7E
GZ_HEADER = bytes([0x1F, 0x8B, 8, 0, 0, 0, 0, 0, 0, 3])
MEMBER_BYTES = 256 * 1024
CHUNK_BYTES = 1 << 20
BLOCK_BYTES = 1 << 16
METRIC = "batch inflate GB/s uncompressed"


def workload_name(n_members, size_mib):
    return (f"BASELINE configs[1]: batch inflate of {n_members} independent 256 KiB gzip members per GPU ({size_mib} MiB "
            "uncompressed per GPU), bodies by zlib level 6 over G_TEXT, CRC-32 per member checked against the trailer "
            "(GzipInputStream semantics)")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b2d", choices=["b2d", "reference"])
    ap.add_argument("--size-mib", type=int, default=1024, help="uncompressed MiB per GPU (1024 = BASELINE configs 2 and 3)")
    ap.add_argument("--config4-gib", type=float, default=8.0, help="total GiB of the config-4 leg (shared by all ranks)")
    ap.add_argument("--edge-gib", type=float, default=4.0, help="total GiB per case of the config-5 leg (shared by all ranks)")
    ap.add_argument("--members", default="zlib", choices=["zlib", "gpu"], help="who encodes the inflate inputs")
    ap.add_argument("--no-deflate", action="store_true", help="skip the deflate leg")
    ap.add_argument("--no-config4", action="store_true")
    ap.add_argument("--no-edge", action="store_true")
    ap.add_argument("--no-single-stream", action="store_true")
    ap.add_argument("--single-stream-mib", type=int, default=256, help="uncompressed MiB of the single-stream leg (rank 0)")
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


# ------------------------------------------------------------------ data (CL = any library exporting b2d_corpus_*)
def make_members(CL, n_members, seed0, pool):
    """-> (list of gzip members, np.uint8 uncompressed blob).  Body: zlib level 6, raw (wbits -15); 10-byte header
    (no optional fields, OS = Unix) and CRC-32 + ISIZE trailer as GzipOutputStream.java:62-70 writes them."""
    raw = np.empty(n_members * MEMBER_BYTES, dtype=np.uint8)

    def one(i):
        view = raw[i * MEMBER_BYTES:(i + 1) * MEMBER_BYTES]
        CL.b2d_corpus_text(seed0 + i, view.ctypes.data, MEMBER_BYTES)
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = c.compress(view.data) + c.flush()
        return GZ_HEADER + body + struct.pack("<II", zlib.crc32(view.data), MEMBER_BYTES)

    return list(pool.map(one, range(n_members))), raw


PIECE = 16 << 20


def fill_corpus(CL, kind, out, seed, first_piece, pool):
    """G_<kind> in independent 16 MiB pieces (seed + global piece index), so that it generates on all cores and a rank
    can make any piece-aligned part of a large input on its own."""
    n = out.size

    def one(k):
        a = k * PIECE
        b = min(n, a + PIECE)
        getattr(CL, "b2d_corpus_" + kind)(seed + first_piece + k, out[a:b].ctypes.data, b - a)

    list(pool.map(one, range((n + PIECE - 1) // PIECE)))
    return out


def make_mixed(CL, n_bytes, seed, pool):
    return fill_corpus(CL, "mixed", np.empty(n_bytes, dtype=np.uint8), seed, 0, pool)


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
def cpu_inflate(O, members, threads, use_zlib=False):
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
            if use_zlib:
                assert zlib.crc32(zlib.decompress(bodies[i], -15)) == crcs[i]
                continue
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


def cpu_zlib_deflate(data, n_chunks, threads, level):
    sizes = [0] * n_chunks

    def work(k):
        for c in range(k, n_chunks, threads):
            z = zlib.compressobj(level, zlib.DEFLATED, -15)
            sizes[c] = len(z.compress(data[c * CHUNK_BYTES:(c + 1) * CHUNK_BYTES].data)) + len(z.flush(zlib.Z_SYNC_FLUSH))
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


class Env:
    """What the legs share: torch, the library, rank/world, streams, timing helpers and the launch counter."""

    def __init__(self, args, b2d, torch, dist, rank, world, local_rank, cores):
        self.args, self.b2d, self.torch, self.dist = args, b2d, torch, dist
        self.rank, self.world, self.local_rank, self.cores = rank, world, local_rank, cores
        self.dev = torch.device("cuda", local_rank)
        self.L = b2d.lib()
        self.threads = max(1, cores // world)
        self.pool = cf.ThreadPoolExecutor(self.threads)
        self.stream = torch.cuda.current_stream()
        self.sp = ctypes.c_void_p(self.stream.cuda_stream)
        self.hbm_peak, self.peak_src = peaks()
        self.launches = 0                   # our kernels launched inside timed regions (b2d_kernel_launches deltas)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _reduce(self, x, op):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(self, x):
        return self._reduce(x, self.dist.ReduceOp.MAX) if self.world > 1 else x

    def sum_over_ranks(self, x):
        return self._reduce(x, self.dist.ReduceOp.SUM) if self.world > 1 else x

    def timed(self, fn, steps):
        """K steps bracketed by barrier + synchronize on both sides, CUDA events on the launching stream, max over
        ranks.  -> seconds for all K steps."""
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        l0 = self.L.b2d_kernel_launches()
        e0.record(self.stream)
        for _ in range(steps):
            fn()
        e1.record(self.stream)
        self.barrier()
        self.launches += self.L.b2d_kernel_launches() - l0
        return self.max_over_ranks(e0.elapsed_time(e1) / 1e3)

    def timed_host(self, fn, steps):
        """The same bracket with the host clock, for blocking host-pointer calls."""
        self.barrier()
        l0 = self.L.b2d_kernel_launches()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        self.torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        self.launches += self.L.b2d_kernel_launches() - l0
        return self.max_over_ranks(dt)


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

    if args.impl == "reference":
        if rank != 0:
            return 0
        return run_reference(args, cores, n_members)

    import b2d_loader
    b2d = b2d_loader.load()
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b2d arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    b2d.init(local_rank)
    env = Env(args, b2d, torch, dist, rank, world, local_rank, cores)

    line, members = inflate_leg(env, n_members)
    if not args.no_deflate:
        line["deflate"] = deflate_leg(env, args.size_mib)
    torch.cuda.empty_cache()
    if not args.no_config4 and args.config4_gib > 0:
        line["config4"] = config4_leg(env, int(args.config4_gib * 1024))
        torch.cuda.empty_cache()
    if not args.no_edge and args.edge_gib > 0:
        line["config5"] = config5_leg(env, int(args.edge_gib * 1024))
        torch.cuda.empty_cache()
    if not args.no_single_stream and args.single_stream_mib > 0:
        if rank == 0:
            line["single_stream"] = single_stream_leg(env, args.single_stream_mib)
        env.barrier()
        torch.cuda.empty_cache()

    # ---------------- CPU baseline (rank 0, N = 1) ----------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        O.build()
        n_st = min(n_members, 512)
        st_s = cpu_inflate(O, members[:n_st], 1)
        mt_s = cpu_inflate(O, members, cores)
        z_mt = cpu_inflate(O, members, cores, use_zlib=True)
        z_st = cpu_inflate(O, members[:n_st], 1, use_zlib=True)
        out_total = n_members * MEMBER_BYTES
        line["cpu_baseline"] = {
            "value": round(out_total / mt_s / 1e9, 4), "unit": "GB/s", "cores": cores, "kind": "port",
            "single_thread": round(n_st * MEMBER_BYTES / st_s / 1e9, 4),
            "zlib_context": {"value": round(out_total / z_mt / 1e9, 4), "single_thread": round(n_st * MEMBER_BYTES / z_st / 1e9, 4),
                             "note": "system zlib inflate + crc32 on the same members: what java.util.zip.Inflater (JDK-bundled zlib) would do"},
            "jvm": "absent (command -v java: not found) -- baseline/RefBench.java is the harness a JDK >= 22 would run",
            "sample": f"oracle_inflate (C restatement of decomp/Open.java; no JVM in this image) + zlib crc32 vs the gzip trailer over all {n_members} members "
                      f"with {cores} threads; single_thread over the first {n_st} members"}
        if "deflate" in line:
            cpu = line["deflate"].pop("_cpu")(O, cores)
            line["deflate"]["cpu_baseline"] = cpu
            line["deflate"]["vs_cpu"] = {
                "rle_dynamic": round(line["deflate"]["e2e"]["value"] / cpu["value"], 2),
                "full_dynamic": round(line["deflate"]["e2e"]["value"] / cpu["full_dynamic_GBps"], 1),
                "device_rle_dynamic": round(line["deflate"]["value"] / cpu["value"], 2),
                "note": "e2e (host buffers) GB/s of this arm / the all-thread oracle port's GB/s: RLE_DYNAMIC is the "
                        "DeflaterOutputStream default, FULL_DYNAMIC the strategy whose ratio this encoder matches"}
    elif "deflate" in line:
        line["deflate"].pop("_cpu", None)
    line["gpu_launches"] = int(env.launches)
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    b2d.shutdown()
    return 0


# ------------------------------------------------------------------ headline: configs[1]
def pcie_ceiling(env, h2d_bytes, d2h_bytes):
    """What this box's PCIe does for one step's bytes: an H2D copy of the step's input and a D2H copy of its output,
    pinned memory, copy engines, both directions at once, every rank at the same time (they share the host's root
    complex and memory).  -> seconds (max over ranks, best of 3)."""
    torch = env.torch
    h_a = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    h_b = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(h2d_bytes, dtype=torch.uint8, device=env.dev)
    d_b = torch.empty(d2h_bytes, dtype=torch.uint8, device=env.dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    best = {"both": 1e9, "h2d": 1e9, "d2h": 1e9}
    for which in ("both", "h2d", "d2h"):
        for _ in range(3):
            env.barrier()
            t0 = time.perf_counter()
            if which != "d2h":
                with torch.cuda.stream(s1):
                    d_a.copy_(h_a, non_blocking=True)
            if which != "h2d":
                with torch.cuda.stream(s2):
                    h_b.copy_(d_b, non_blocking=True)
            torch.cuda.synchronize()
            best[which] = min(best[which], env.max_over_ranks(time.perf_counter() - t0))
    return best


def inflate_leg(env, n_members):
    args, b2d, torch, L, dev, rank, world = env.args, env.b2d, env.torch, env.L, env.dev, env.rank, env.world
    # ---------------- inflate inputs (this rank's shard: members [rank*n, (rank+1)*n)) ----------------
    t_prep = time.perf_counter()
    if args.members == "zlib":
        members, raw = make_members(L, n_members, SEED + rank * n_members, env.pool)
    else:
        raw = np.empty(n_members * MEMBER_BYTES, dtype=np.uint8)
        list(env.pool.map(lambda i: L.b2d_corpus_text(SEED + rank * n_members + i, raw[i * MEMBER_BYTES:].ctypes.data,
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
    sp = env.sp

    def inflate_dev(flags):
        r = L.b2d_inflate_batch_dev(d_blob.data_ptr(), d_in_off.data_ptr(), n_members, d_out.data_ptr(),
                                    d_out_off.data_ptr(), d_out_len.data_ptr(), d_cons.data_ptr(), d_crc.data_ptr(),
                                    d_status.data_ptr(), flags, sp)
        if r != 0:
            raise RuntimeError(f"b2d_inflate_batch_dev: {b2d.status_name(r)}")

    # warm-up + correctness at full size (every byte, every member)
    warm = max(3, args.warmup)
    for _ in range(warm):
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

    sampler = ClockSampler(env.local_rank)
    sampler.start()
    time.sleep(0.3)
    tc0 = time.perf_counter()
    # the headline step: inflate kernel + CRC-32 of every member's output
    step_s = env.timed(lambda: inflate_dev(b2d.INFLATE_CRC32), args.steps) / args.steps
    # the dominant kernel alone (roofline)
    kern_s = env.timed(lambda: inflate_dev(0), args.steps) / args.steps
    tc1 = time.perf_counter()
    clocks = sampler.stop(tc0, tc1)
    total_uncomp = env.sum_over_ranks(float(out_total))
    value = total_uncomp / step_s / 1e9
    algo_bytes = comp_total + out_total
    achieved = algo_bytes / kern_s / 1e9

    # e2e: host pointers through b2d_gunzip_batch (H2D of the compressed blob, kernels, D2H of the output)
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
    assert np.array_equal(h_out.numpy()[-(1 << 24):], raw[-(1 << 24):])
    assert np.array_equal(h_cons.astype(np.int64), mem_off[1:] - mem_off[:-1])
    e2e_steps = max(3, args.steps // 2)
    e2e_s = env.timed_host(inflate_host, e2e_steps) / e2e_steps
    e2e_val = total_uncomp / e2e_s / 1e9
    h2d = comp_total + 3 * (n_members + 1) * 8
    d2h = out_total + n_members * (8 + 8 + 4 + 4)
    ceil = pcie_ceiling(env, comp_total, out_total)
    ceiling_gbs = total_uncomp / ceil["both"] / 1e9

    line = {
        "metric": METRIC,
        "value": round(value, 3), "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": round(step_s * 1e3, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(n_members, args.size_mib),
                   "baseline_metric": "deflate/inflate GB/s uncompressed at 1/2/4/8 B200; ratio vs ref; % HBM peak "
                                      "(headline = the inflate leg; the deflate leg, config 4 and config 5 ride in \"deflate\", \"config4\", \"config5\")",
                   "members_per_gpu": n_members, "member_bytes": MEMBER_BYTES,
                   "members_encoded_by": "zlib level 6" if args.members == "zlib" else "the GPU encoder",
                   "compressed_bytes_per_gpu": comp_total, "ratio": round(out_total / comp_total, 4),
                   "sharding": f"members by rank, no collective ({world} rank(s))",
                   "l2": "inputs larger than L2 (compressed blob + 1 GiB output > 126 MB), no flush needed",
                   "e2e_call": "b2d_gunzip_batch (host pointers, pinned): header checks, H2D, inflate + CRC-32 kernels, D2H, trailer checks",
                   "e2e_timer": "host clock around the blocking C-ABI call (ends with a stream sync), max over ranks",
                   "prep_s": round(prep_s, 1)},
        "e2e": {"value": round(e2e_val, 3), "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": round(e2e_s * 1e3, 3),
                "ceiling_gbs": round(ceiling_gbs, 3), "frac": round(e2e_val / ceiling_gbs, 4),
                "ceiling": {"both_ms": round(ceil["both"] * 1e3, 3), "h2d_ms": round(ceil["h2d"] * 1e3, 3),
                            "d2h_ms": round(ceil["d2h"] * 1e3, 3),
                            "h2d_gbs_per_gpu": round(comp_total / ceil["h2d"] / 1e9, 2),
                            "d2h_gbs_per_gpu": round(out_total / ceil["d2h"] / 1e9, 2),
                            "note": "copy-engine H2D of the step's compressed bytes and D2H of its output, pinned memory, both "
                                    f"at once, all {world} rank(s) at the same time (max over ranks, best of 3): the PCIe time "
                                    "a perfect pipeline could not go below"}},
        "roofline": {"bound": "hbm", "kernel": "b2d::inflate_kernel", "achieved": round(achieved, 2), "peak": env.hbm_peak,
                     "unit": "GB/s", "frac": round(achieved / env.hbm_peak, 5),
                     "traffic": measured_traffic("inflate_kernel", n_members), "peak_source": env.peak_src,
                     "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": round(kern_s * 1e3, 4),
                     "note": "latency/issue-bound by construction (serial Huffman decode per member); see DESIGN.md"},
        "clocks": clocks,
    }
    return line, members


# ------------------------------------------------------------------ configs[2]
def deflate_leg(env, n_chunks):
    args, b2d, torch, L, dev, rank, world, sp = env.args, env.b2d, env.torch, env.L, env.dev, env.rank, env.world, env.sp
    n_bytes = n_chunks * CHUNK_BYTES
    data = make_mixed(L, n_bytes, SEED + 1000 * rank, env.pool)
    h_in = torch.from_numpy(data).pin_memory()
    d_in = h_in.to(dev)
    bound = b2d.deflate_bound(n_bytes, CHUNK_BYTES)
    d_out = torch.empty(bound, dtype=torch.uint8, device=dev)
    d_total = torch.zeros(1, dtype=torch.int64, device=dev)
    d_clen = torch.zeros(n_chunks, dtype=torch.int64, device=dev)
    d_ccrc = torch.zeros(n_chunks, dtype=torch.int32, device=dev)
    opts = b2d.make_opts(chunk_bytes=CHUNK_BYTES, block_bytes=BLOCK_BYTES, mode=b2d.MODE_AUTO, is_last=1)
    d_bits = torch.zeros(n_bytes // BLOCK_BYTES, dtype=torch.int32, device=dev)     # restart index: bit offset of every block

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
    dec_steps = max(3, args.steps // 2)
    for _ in range(3):
        inflate_chunks()
    torch.cuda.synchronize()
    # configs[3]'s decompress direction: the chunk-indexed stream decoded with one warp per 1 MiB chunk
    unchunk_s = env.timed(inflate_chunks, dec_steps) / dec_steps
    assert int(d_st.abs().sum().item()) == 0 and torch.equal(d_dec, d_in), "deflate: GPU round trip differs"
    assert torch.equal(d_c2, d_ccrc), "deflate: chunk CRCs differ from the CRCs of the decoded chunks"
    # ... and with the block index: one warp per 64 KiB block, references replayed per chunk afterwards
    d_dec.zero_()
    d_cst = torch.zeros(n_chunks, dtype=torch.int32, device=dev)

    def inflate_blocks():
        r = L.b2d_inflate_chunks_dev(d_out.data_ptr(), d_coff.data_ptr(), n_chunks, d_bits.data_ptr(), CHUNK_BYTES, BLOCK_BYTES,
                                     n_bytes, d_dec.data_ptr(), d_c2.data_ptr(), d_cst.data_ptr(), b2d.INFLATE_CRC32, sp)
        assert r == 0
    for _ in range(3):
        inflate_blocks()
    torch.cuda.synchronize()
    assert int(d_cst.abs().sum().item()) == 0 and torch.equal(d_dec, d_in) and torch.equal(d_c2, d_ccrc), "block-indexed decode differs"
    unblock_s = env.timed(inflate_blocks, dec_steps) / dec_steps
    h_comp = d_out[:comp_len].cpu().numpy()
    for c in range(0, n_chunks, max(1, n_chunks // 16)):
        d = zlib.decompressobj(-15)
        got = d.decompress(h_comp[coff[c]:coff[c + 1]].tobytes())
        assert got == data[c * CHUNK_BYTES:(c + 1) * CHUNK_BYTES].tobytes(), f"deflate: zlib decode of chunk {c} differs"
    del d_dec

    step_s = env.timed(deflate_dev, args.steps) / args.steps
    total_in = env.sum_over_ranks(float(n_bytes))
    value = total_in / step_s / 1e9

    # the same input with adaptive block splitting (pieces of 16 KiB, the GPU counterpart of BinarySplit)
    opts_split = b2d.make_opts(chunk_bytes=CHUNK_BYTES, block_bytes=BLOCK_BYTES, mode=b2d.MODE_AUTO, is_last=1, split_min_bytes=16384)
    d_total_s = torch.zeros(1, dtype=torch.int64, device=dev)

    def deflate_split_dev():
        r = L.b2d_deflate_chunks_dev(d_in.data_ptr(), n_bytes, ctypes.byref(opts_split), d_out.data_ptr(), bound,
                                     d_total_s.data_ptr(), d_clen.data_ptr(), d_ccrc.data_ptr(), sp)
        if r != 0:
            raise RuntimeError(f"b2d_deflate_chunks_dev (split): {b2d.status_name(r)}")
    split_steps = max(3, args.steps // 2)
    for _ in range(2):
        deflate_split_dev()
    split_s = env.timed(deflate_split_dev, split_steps) / split_steps
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
    e2e_s = env.timed_host(deflate_host, e2e_steps) / e2e_steps

    res = {
        "workload": f"BASELINE configs[2]: chunked dynamic-Huffman deflate of {n_chunks} MiB G_MIXED per GPU, 1 MiB chunks "
                    "+ sync-flush markers, 64 KiB blocks, mode auto, CRC-32 per chunk (weak scaling, no exchange: config4 has the gather)",
        "value": round(value, 3), "unit": "GB/s", "ms_per_step": round(step_s * 1e3, 3),
        "compressed_bytes_per_gpu": comp_len, "ratio": round(n_bytes / comp_len, 4),
        "e2e": {"value": round(total_in / e2e_s / 1e9, 3), "unit": "GB/s", "h2d_bytes_per_step": n_bytes,
                "d2h_bytes_per_step": comp_len + n_chunks * 12 + 8, "ms_per_step": round(e2e_s * 1e3, 3)},
        "roofline": {"bound": "hbm", "achieved": round((n_bytes + comp_len) / step_s / 1e9, 2), "peak": env.hbm_peak,
                     "unit": "GB/s", "frac": round((n_bytes + comp_len) / step_s / 1e9 / env.hbm_peak, 5),
                     "note": "whole pipeline (all kernels of one call + memset); algorithmic bytes = input read + compressed written"},
        "inflate_chunk_indexed": {"value": round(total_in / unchunk_s / 1e9, 3), "unit": "GB/s",
                                  "ms_per_step": round(unchunk_s * 1e3, 3),
                                  "note": f"decode of this stream, one warp per 1 MiB chunk ({n_chunks} units per GPU: latency-bound below ~4096 units)"},
        "inflate_block_indexed": {"value": round(total_in / unblock_s / 1e9, 3), "unit": "GB/s",
                                  "ms_per_step": round(unblock_s * 1e3, 3),
                                  "note": "b2d_inflate_chunks_dev: one warp per 64 KiB block (Huffman decode), then one warp per chunk replays the back-references"},
        "adaptive_split": {"value": round(total_in / split_s / 1e9, 3), "unit": "GB/s",
                           "ms_per_step": round(split_s * 1e3, 3), "compressed_bytes_per_gpu": split_len,
                           "ratio": round(n_bytes / split_len, 4), "bytes_vs_unsplit": round(split_len / comp_len, 5),
                           "note": "split_min_bytes = 16 KiB: every 64 KiB span becomes the cheapest partition of its tree of "
                                   "pieces (SURVEY 8f N3, comp/BinarySplit.java)"},
    }

    def cpu(O, cores):
        n_s = min(n_chunks, 8 * cores)
        st_s, _ = cpu_deflate(O, data, min(n_chunks, 16), 1, O.RLE_DYNAMIC)
        mt_s, rle_bytes = cpu_deflate(O, data, n_s, cores, O.RLE_DYNAMIC)
        n_f = min(n_chunks, 2 * cores)
        f_s, full_bytes = cpu_deflate(O, data, n_f, cores, O.FULL_DYNAMIC)
        z1_s, z1_b = cpu_zlib_deflate(data, n_s, cores, 1)
        z6_s, z6_b = cpu_zlib_deflate(data, n_s, cores, 6)
        gpu_same = int(clen[:n_f].sum())
        gpu_ns = int(clen[:n_s].sum())
        return {"value": round(n_s * CHUNK_BYTES / mt_s / 1e9, 4), "unit": "GB/s", "cores": cores, "kind": "port",
                "single_thread": round(min(n_chunks, 16) * CHUNK_BYTES / st_s / 1e9, 4),
                "sample": f"oracle_deflate RLE_DYNAMIC (the DeflaterOutputStream default, 64 KiB blocks) over the first {n_s} "
                          f"chunks with {cores} threads; FULL_DYNAMIC over the first {n_f} chunks for the ratio bar",
                "full_dynamic_GBps": round(n_f * CHUNK_BYTES / f_s / 1e9, 5),
                "ratio_rle_dynamic": round(n_s * CHUNK_BYTES / rle_bytes, 4),
                "ratio_full_dynamic": round(n_f * CHUNK_BYTES / (full_bytes + 5 * n_f), 4),
                "gpu_bytes_vs_full_dynamic_same_chunks": round(gpu_same / (full_bytes + 5 * n_f), 5),
                "zlib_context": {"level1_GBps": round(n_s * CHUNK_BYTES / z1_s / 1e9, 4), "level6_GBps": round(n_s * CHUNK_BYTES / z6_s / 1e9, 4),
                                 "gpu_bytes_vs_level1": round(gpu_ns / z1_b, 5), "gpu_bytes_vs_level6": round(gpu_ns / z6_b, 5),
                                 "note": f"system zlib (java.util.zip.Deflater's engine), sync-flushed 1 MiB chunks, first {n_s} chunks, {cores} threads"}}
    res["_cpu"] = cpu
    return res


# ------------------------------------------------------------------ configs[3]: one input shared by all ranks
def config4_leg(env, total_mib, n_slices=4):
    """8 GiB G_MIXED = 8192 chunks of 1 MiB shared by the N ranks, slice-major (sharding.slice_ranges): per slice k every
    rank compresses its own contiguous chunk range, the slice's sizes / CRCs / block index are all-gathered, and the
    payloads travel to GPU 0 straight to their final offsets while slice k + 1 is being compressed."""
    args, b2d, torch, dist, L, dev, rank, world, sp = (env.args, env.b2d, env.torch, env.dist, env.L, env.dev, env.rank,
                                                       env.world, env.sp)
    from importlib import import_module
    sharding = import_module("b2deflate.sharding")
    n_chunks = total_mib
    # slices of at least 512 MiB per rank (one full wave of the encoder's chains warps: smaller launches only add up
    # their latencies), at most four
    n_slices = max(1, min(n_slices, n_chunks // world // 512))
    bpc = CHUNK_BYTES // BLOCK_BYTES
    ranges = sharding.slice_ranges(n_chunks, rank, world, n_slices)
    my_chunks = sum(hi - lo for lo, hi in ranges)
    my_bytes = my_chunks * CHUNK_BYTES
    # this rank's chunks, slice after slice, generated piece by piece (16 MiB pieces seeded by their global index)
    t_prep = time.perf_counter()
    d_in = torch.empty(my_bytes, dtype=torch.uint8, device=dev)
    stage = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    cpp = PIECE // CHUNK_BYTES
    o = 0
    for lo, hi in ranges:
        c = lo
        while c < hi:
            e = min(hi, c + stage.numel() // CHUNK_BYTES)
            # chunks [c, e): the 16 MiB pieces that cover them (pieces are generated whole, then cut)
            p0, p1 = c // cpp, (e + cpp - 1) // cpp
            buf = np.empty((p1 - p0) * PIECE, dtype=np.uint8)
            fill_corpus(L, "mixed", buf, SEED + 4000, p0, env.pool)
            a = (c - p0 * cpp) * CHUNK_BYTES
            n = (e - c) * CHUNK_BYTES
            stage[:n] = torch.from_numpy(buf[a:a + n])
            d_in[o:o + n].copy_(stage[:n], non_blocking=True)
            torch.cuda.current_stream().synchronize()
            o += n
            c = e
    assert o == my_bytes
    del stage
    prep_s = time.perf_counter() - t_prep

    slice_chunks = [hi - lo for lo, hi in ranges]
    slice_off = np.concatenate([[0], np.cumsum(slice_chunks)]) * CHUNK_BYTES
    max_sc = max(1, int(env.max_over_ranks(float(max(slice_chunks)))))
    bound = b2d.deflate_bound(max_sc * CHUNK_BYTES, CHUNK_BYTES)
    d_out = [torch.empty(bound, dtype=torch.uint8, device=dev) for _ in range(n_slices)]
    d_total = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(n_slices)]
    d_clen = [torch.zeros(max_sc, dtype=torch.int64, device=dev) for _ in range(n_slices)]
    d_ccrc = [torch.zeros(max_sc, dtype=torch.int32, device=dev) for _ in range(n_slices)]
    d_bits = [torch.zeros(max_sc * bpc, dtype=torch.int32, device=dev) for _ in range(n_slices)]
    # the last chunk of the stream lives in the last slice of the last rank
    opts = [b2d.make_opts(chunk_bytes=CHUNK_BYTES, block_bytes=BLOCK_BYTES, mode=b2d.MODE_AUTO,
                          is_last=int(k == n_slices - 1 and rank == world - 1)) for k in range(n_slices)]
    gathered = torch.empty(b2d.deflate_bound(n_chunks * CHUNK_BYTES, CHUNK_BYTES), dtype=torch.uint8, device=dev) if rank == 0 else None
    meta_len = 2 + max_sc * (2 + bpc)
    d_allmeta = [torch.zeros(world * meta_len, dtype=torch.int64, device=dev) for _ in range(n_slices)]
    comm = torch.cuda.Stream()
    evs = [torch.cuda.Event() for _ in range(n_slices)]
    state = {}

    def compress_slice(k):
        n = slice_chunks[k] * CHUNK_BYTES
        r = L.b2d_deflate_chunks_indexed_dev(d_in.data_ptr() + int(slice_off[k]), n, ctypes.byref(opts[k]), d_out[k].data_ptr(), bound,
                                             d_total[k].data_ptr(), d_clen[k].data_ptr(), d_ccrc[k].data_ptr(), d_bits[k].data_ptr(), sp)
        if r != 0:
            raise RuntimeError(f"config4: b2d_deflate_chunks_indexed_dev: {b2d.status_name(r)}")

    def compress_only():
        for k in range(n_slices):
            compress_slice(k)

    def compress_and_gather():
        for k in range(n_slices):
            compress_slice(k)
            evs[k].record(env.stream)
        base = 0
        metas = []
        with torch.cuda.stream(comm):
            for k in range(n_slices):
                comm.wait_event(evs[k])
                sc = slice_chunks[k]
                meta = torch.zeros(meta_len, dtype=torch.int64, device=dev)
                meta[0] = d_total[k][0]
                meta[1] = sc
                meta[2:2 + sc] = d_clen[k][:sc]
                meta[2 + max_sc:2 + max_sc + sc] = d_ccrc[k][:sc].to(torch.int64) & 0xFFFFFFFF
                meta[2 + 2 * max_sc:2 + 2 * max_sc + sc * bpc] = d_bits[k][:sc * bpc].to(torch.int64)
                if world > 1:
                    dist.all_gather_into_tensor(d_allmeta[k], meta)
                else:
                    d_allmeta[k].copy_(meta)
                h = d_allmeta[k].view(world, meta_len).cpu()          # waits for slice k's kernels only: k + 1.. are queued behind
                totals = [int(h[r, 0]) for r in range(world)]
                if world > 1:
                    base += sharding.gather_slice(d_out[k], totals[rank], totals, gathered, base)
                else:
                    gathered[base:base + totals[0]].copy_(d_out[k][:totals[0]], non_blocking=True)
                    base += totals[0]
                metas.append(h)
        env.stream.wait_stream(comm)
        state["metas"], state["total"] = metas, base

    for _ in range(2):
        compress_and_gather()
    torch.cuda.synchronize()
    steps = max(3, args.steps // 2)
    only_s = env.timed(compress_only, steps) / steps
    step_s = env.timed(compress_and_gather, steps) / steps
    total_bytes = n_chunks * CHUNK_BYTES
    metas, comp_total = state["metas"], state["total"]

    # ---- the stream's index in global order (slice-major, rank-minor), from the all-gathered metas
    sizes, crcs, bits = [], [], []
    for k in range(n_slices):
        h = metas[k].numpy()
        for r in range(world):
            sc = int(h[r, 1])
            sizes.append(h[r, 2:2 + sc]); crcs.append(h[r, 2 + max_sc:2 + max_sc + sc])
            bits.append(h[r, 2 + 2 * max_sc:2 + 2 * max_sc + sc * bpc])
    sizes = np.concatenate(sizes); crcs = np.concatenate(crcs).astype(np.uint32); bits = np.concatenate(bits).astype(np.uint32)
    assert sizes.size == n_chunks and int(sizes.sum()) == comp_total

    # ---- decompress direction, same sharding: every rank decodes its own slices block-parallel (+ CRC-32)
    d_dec = torch.zeros(my_bytes, dtype=torch.uint8, device=dev)
    d_coff = []
    for k in range(n_slices):
        off = torch.zeros(slice_chunks[k] + 1, dtype=torch.int64, device=dev)
        off[1:] = torch.cumsum(d_clen[k][:slice_chunks[k]], 0)
        d_coff.append(off)
    d_dcrc = [torch.zeros(max_sc, dtype=torch.int32, device=dev) for _ in range(n_slices)]
    d_dst = [torch.zeros(max_sc, dtype=torch.int32, device=dev) for _ in range(n_slices)]

    def decompress():
        for k in range(n_slices):
            if slice_chunks[k] == 0:
                continue
            r = L.b2d_inflate_chunks_dev(d_out[k].data_ptr(), d_coff[k].data_ptr(), slice_chunks[k], d_bits[k].data_ptr(), CHUNK_BYTES,
                                         BLOCK_BYTES, slice_chunks[k] * CHUNK_BYTES, d_dec.data_ptr() + int(slice_off[k]),
                                         d_dcrc[k].data_ptr(), d_dst[k].data_ptr(), b2d.INFLATE_CRC32, sp)
            assert r == 0
    for _ in range(2):
        decompress()
    torch.cuda.synchronize()
    assert torch.equal(d_dec, d_in), "config4: this rank's round trip differs"
    for k in range(n_slices):
        sc = slice_chunks[k]
        assert int(d_dst[k][:sc].abs().sum().item()) == 0 and torch.equal(d_dcrc[k][:sc], d_ccrc[k][:sc])
    dec_s = env.timed(decompress, steps) / steps
    del d_dec

    # ---- rank 0 verifies the GATHERED stream (every rank's payload): GPU decode of the whole stream against the input
    # CRCs that came with the all-gather, and a strided sample of chunks through zlib
    verify = None
    if rank == 0:
        coff = np.zeros(n_chunks + 1, dtype=np.int64); coff[1:] = np.cumsum(sizes)
        d_goff = torch.from_numpy(coff).to(dev)
        d_gbits = torch.from_numpy(bits.astype(np.int32)).to(dev)
        d_whole = torch.empty(total_bytes, dtype=torch.uint8, device=dev)
        d_gcrc = torch.zeros(n_chunks, dtype=torch.int32, device=dev); d_gst = torch.zeros(n_chunks, dtype=torch.int32, device=dev)
        r = L.b2d_inflate_chunks_dev(gathered.data_ptr(), d_goff.data_ptr(), n_chunks, d_gbits.data_ptr(), CHUNK_BYTES, BLOCK_BYTES,
                                     total_bytes, d_whole.data_ptr(), d_gcrc.data_ptr(), d_gst.data_ptr(), b2d.INFLATE_CRC32, sp)
        assert r == 0
        torch.cuda.synchronize()
        assert int(d_gst.abs().sum().item()) == 0, "config4: a chunk of the gathered stream failed to decode"
        got = d_gcrc.cpu().numpy().view(np.uint32)
        assert np.array_equal(got, crcs), "config4: gathered stream decodes to different bytes than the ranks compressed"
        # rank 0's own chunks byte for byte (its slices sit at known places of the global order)
        g = 0
        for k in range(n_slices):
            lo, hi = sharding.unit_range(n_chunks, k, n_slices)
            a, b = sharding.unit_range(hi - lo, 0, world)
            assert torch.equal(d_whole[(lo + a) * CHUNK_BYTES:(lo + b) * CHUNK_BYTES],
                               d_in[int(slice_off[k]):int(slice_off[k]) + (b - a) * CHUNK_BYTES])
        n_z = 0
        for c in range(0, n_chunks, max(1, n_chunks // 64)):
            body = gathered[int(coff[c]):int(coff[c + 1])].cpu().numpy().tobytes()
            z = zlib.decompressobj(-15)
            out = z.decompress(body)
            assert len(out) == CHUNK_BYTES and zlib.crc32(out) == int(crcs[c]), f"config4: zlib decode of gathered chunk {c} differs"
            n_z += 1
        crc_all = 0
        for c in range(n_chunks):
            crc_all = L.b2d_crc32_combine(crc_all, int(crcs[c]), CHUNK_BYTES)
        zl = zlib.decompressobj(-15)                       # the stream's last chunk carries BFINAL: zlib sees the end there
        zl.decompress(gathered[int(coff[n_chunks - 1]):comp_total].cpu().numpy().tobytes())
        assert zl.eof and gathered[comp_total - 4:comp_total].cpu().numpy().tobytes() == b"\x00\x00\xff\xff", \
            "config4: the gathered stream does not end with the final empty stored block"
        verify = {"gathered_stream_bytes": comp_total, "chunks_gpu_decoded": n_chunks, "chunks_zlib_decoded": n_z,
                  "crc32_of_input_from_chunk_crcs": f"{crc_all:08x}",
                  "how": "rank 0: b2d_inflate_chunks_dev over the whole gathered stream, per-chunk CRC-32 == the input CRC-32 "
                         "every rank computed before compressing (all-gathered with the sizes); own chunks compared byte for "
                         "byte; strided zlib sample; stream ends with BFINAL marker"}
        del d_whole
    total = float(total_bytes)
    return {
        "workload": f"BASELINE configs[3]: {total_mib / 1024:g} GiB G_MIXED = {n_chunks} chunks of 1 MiB shared by {world} rank(s) "
                    f"(strong scaling), {n_slices} slices per rank; compress -> size all-gather -> payloads to GPU 0 inside the timed step",
        "scaling": "strong",
        "deflate": {"value": round(total / step_s / 1e9, 3), "unit": "GB/s", "ms_per_step": round(step_s * 1e3, 3),
                    "compress_only_ms": round(only_s * 1e3, 3), "gather_exposed_ms": round((step_s - only_s) * 1e3, 3),
                    "note": "value includes the all-gather of sizes / CRCs / block index and the payload gather to GPU 0 (NCCL batched "
                            "send/recv per slice on a side stream, under the next slice's kernels); gather_exposed_ms = step - the same "
                            "kernels without any exchange"},
        "inflate": {"value": round(total / dec_s / 1e9, 3), "unit": "GB/s", "ms_per_step": round(dec_s * 1e3, 3),
                    "note": "b2d_inflate_chunks_dev (block-parallel, CRC-32 per chunk) over each rank's own chunks; no collective"},
        "compressed_bytes": comp_total, "ratio": round(total_bytes / comp_total, 4),
        "chunks_per_rank": my_chunks, "verify": verify, "prep_s": round(prep_s, 1),
    }


# ------------------------------------------------------------------ configs[4]: edge cases
def config5_leg(env, total_mib):
    args, b2d, torch, L, dev, rank, world, sp = env.args, env.b2d, env.torch, env.L, env.dev, env.rank, env.world, env.sp
    from importlib import import_module
    sharding = import_module("b2deflate.sharding")
    n_chunks_all = total_mib
    lo, hi = sharding.unit_range(n_chunks_all // (PIECE // CHUNK_BYTES), rank, world)       # whole 16 MiB pieces per rank
    n_bytes = (hi - lo) * PIECE
    n_chunks = n_bytes // CHUNK_BYTES
    out = {"workload": f"BASELINE configs[4]: {total_mib / 1024:g} GiB per case shared by {world} rank(s) (strong scaling): random bytes, "
                       "zeros, fixed-Huffman-only text; compress, decode block-parallel and one warp per chunk, byte-exact",
           "scaling": "strong"}
    if n_bytes == 0:
        return out
    bound = b2d.deflate_bound(n_bytes, CHUNK_BYTES)
    d_in = torch.empty(n_bytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(bound, dtype=torch.uint8, device=dev)
    d_dec = torch.empty(n_bytes, dtype=torch.uint8, device=dev)
    d_total = torch.zeros(1, dtype=torch.int64, device=dev)
    d_clen = torch.zeros(n_chunks, dtype=torch.int64, device=dev)
    d_ccrc = torch.zeros(n_chunks, dtype=torch.int32, device=dev)
    d_bits = torch.zeros(n_chunks * (CHUNK_BYTES // BLOCK_BYTES), dtype=torch.int32, device=dev)
    d_c2 = torch.zeros(n_chunks, dtype=torch.int32, device=dev); d_st = torch.zeros(n_chunks, dtype=torch.int32, device=dev)
    d_ol = torch.zeros(n_chunks, dtype=torch.int64, device=dev); d_ic = torch.zeros(n_chunks, dtype=torch.int64, device=dev)
    d_ooff = torch.arange(n_chunks + 1, dtype=torch.int64, device=dev) * CHUNK_BYTES
    steps = max(3, args.steps // 4)
    host = np.empty(min(n_bytes, 256 << 20), dtype=np.uint8)
    for kind in ("random", "zeros", "fixed"):
        keep = {}                                   # host copies of a few chunks for the zlib check
        sample = list(range(0, n_chunks, max(1, n_chunks // 8)))
        if kind == "zeros":
            d_in.zero_()
            for c in sample:
                keep[c] = bytes(CHUNK_BYTES)
        else:
            for a in range(0, n_bytes, host.size):
                n = min(host.size, n_bytes - a)
                fill_corpus(L, "random" if kind == "random" else "text", host[:n], SEED + (7000 if kind == "random" else 9000),
                            lo + a // PIECE, env.pool)
                d_in[a:a + n].copy_(torch.from_numpy(host[:n]))
                for c in sample:
                    if a <= c * CHUNK_BYTES < a + n:
                        keep[c] = host[c * CHUNK_BYTES - a:(c + 1) * CHUNK_BYTES - a].tobytes()
        opts = b2d.make_opts(chunk_bytes=CHUNK_BYTES, block_bytes=BLOCK_BYTES, mode=b2d.MODE_FIXED if kind == "fixed" else b2d.MODE_AUTO,
                             is_last=int(rank == world - 1))

        def compress():
            r = L.b2d_deflate_chunks_indexed_dev(d_in.data_ptr(), n_bytes, ctypes.byref(opts), d_out.data_ptr(), bound, d_total.data_ptr(),
                                                 d_clen.data_ptr(), d_ccrc.data_ptr(), d_bits.data_ptr(), sp)
            assert r == 0, b2d.status_name(r)
        for _ in range(2):
            compress()
        comp_s = env.timed(compress, steps) / steps
        comp_len = int(d_total.item())
        clen = d_clen.cpu().numpy()
        coff = np.zeros(n_chunks + 1, dtype=np.int64); coff[1:] = np.cumsum(clen)
        assert int(coff[-1]) == comp_len
        d_coff = torch.from_numpy(coff).to(dev)

        def dec_blocks():
            r = L.b2d_inflate_chunks_dev(d_out.data_ptr(), d_coff.data_ptr(), n_chunks, d_bits.data_ptr(), CHUNK_BYTES, BLOCK_BYTES, n_bytes,
                                         d_dec.data_ptr(), d_c2.data_ptr(), d_st.data_ptr(), b2d.INFLATE_CRC32, sp)
            assert r == 0

        def dec_chunks():
            r = L.b2d_inflate_batch_dev(d_out.data_ptr(), d_coff.data_ptr(), n_chunks, d_dec.data_ptr(), d_ooff.data_ptr(), d_ol.data_ptr(),
                                        d_ic.data_ptr(), d_c2.data_ptr(), d_st.data_ptr(), b2d.INFLATE_CHUNK_INDEXED | b2d.INFLATE_CRC32, sp)
            assert r == 0
        res = {}
        for name, fn in (("inflate_block_indexed", dec_blocks), ("inflate_chunk_indexed", dec_chunks)):
            d_dec.zero_()
            fn()
            torch.cuda.synchronize()
            assert int(d_st.abs().sum().item()) == 0 and torch.equal(d_dec, d_in) and torch.equal(d_c2, d_ccrc), f"config5 {kind}: {name} differs"
            s = env.timed(fn, steps) / steps
            res[name] = round(env.sum_over_ranks(float(n_bytes)) / s / 1e9, 3)
        for c in sample:                                 # zlib (java.util.zip.Inflater's engine) reads the chunks alone
            body = d_out[int(coff[c]):int(coff[c + 1])].cpu().numpy().tobytes()
            assert zlib.decompressobj(-15).decompress(body) == keep[c], f"config5 {kind}: zlib decode of chunk {c} differs"
            first = body[0] & 7
            if kind == "fixed":
                assert first == 0b010, "config5 fixed: a chunk does not start with a fixed-Huffman block"
            if kind == "random":
                assert first == 0b000, "config5 random: a chunk does not start with a stored block"
        if kind == "random":                             # Uncompressed.java:23-25: two stored pieces per 64 KiB block + the chunk marker
            assert comp_len == n_bytes + 10 * (n_bytes >> 16) + 5 * (n_bytes >> 20)
        tot = env.sum_over_ranks(float(n_bytes))
        comp_all = env.sum_over_ranks(float(comp_len))
        out[kind] = {"deflate": round(tot / comp_s / 1e9, 3), **res, "unit": "GB/s", "ratio": round(tot / comp_all, 3),
                     "zlib_checked_chunks_per_rank": len(sample)}
    out["bytes_per_rank"] = n_bytes
    out["note"] = ("per case: device-resident, CUDA events, max over ranks; zeros = literal 0 + (258, distance 1) matches, the decoder's "
                   "pattern-replication path (Open.java:596-603; InflaterInputStreamTest.testFixedHuffmanOverlappingRun1)")
    return out


# ------------------------------------------------------------------ ONE foreign stream (InflaterInputStream on a file from system gzip)
def single_stream_leg(env, mib):
    """One raw-DEFLATE stream made by zlib level 6 with history carried across blocks -- what `gunzip` of a system-gzip
    file hands to InflaterInputStream (InflaterInputStream.java:147-164 -> Open.java:83-110), no index, no chunk
    boundaries -- through b2d_inflate_stream (speculative parallel decode).  Rank 0 only: one stream is one unit."""
    args, b2d, torch, L, dev = env.args, env.b2d, env.torch, env.L, env.dev
    n = mib << 20
    data = fill_corpus(L, "text", np.empty(n, dtype=np.uint8), SEED + 31000, 0, env.pool)
    t = time.perf_counter()
    z = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = z.compress(data.data) + z.flush()
    t_comp = time.perf_counter() - t
    t = time.perf_counter()
    assert zlib.crc32(zlib.decompress(comp, -15)) == zlib.crc32(data.data)
    t_zlib = time.perf_counter() - t
    m = len(comp)
    d_in = torch.zeros(m + 64, dtype=torch.uint8, device=dev)
    d_in[:m] = torch.from_numpy(np.frombuffer(comp, dtype=np.uint8).copy()).to(dev)
    d_out = torch.zeros(n + 256, dtype=torch.uint8, device=dev)
    d_res = torch.zeros(8, dtype=torch.int64, device=dev)

    def dev_call():
        r = L.b2d_inflate_stream_dev(d_in.data_ptr(), m, d_out.data_ptr(), n, d_res.data_ptr(), env.sp)
        assert r == 0, b2d.status_name(r)
    for _ in range(2):
        dev_call()
    torch.cuda.synchronize()
    res = d_res.cpu().numpy()
    assert int(res[2]) == 0 and int(res[0]) == n and int(res[1]) == m, "single stream: the parallel decode did not vouch for the result"
    assert torch.equal(d_out[:n], torch.from_numpy(data).to(dev)), "single stream: output differs"
    steps = max(3, args.steps // 2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = L.b2d_kernel_launches()
    e0.record(env.stream)
    for _ in range(steps):
        dev_call()
    e1.record(env.stream)
    torch.cuda.synchronize()
    env.launches += L.b2d_kernel_launches() - l0
    dev_s = e0.elapsed_time(e1) / 1e3 / steps
    del d_out, d_in
    # host pointers (pinned): H2D, decode, CRC-32, D2H inside the call
    h_in = torch.from_numpy(np.frombuffer(comp, dtype=np.uint8).copy()).pin_memory()
    h_out = torch.empty(n + 64, dtype=torch.uint8).pin_memory()
    ol, ic = ctypes.c_uint64(0), ctypes.c_uint64(0)
    crc, st, par = ctypes.c_uint32(0), ctypes.c_int32(0), ctypes.c_int32(0)

    def host_call():
        r = L.b2d_inflate_stream(h_in.data_ptr(), m, h_out.data_ptr(), n, ctypes.byref(ol), ctypes.byref(ic), ctypes.byref(crc),
                                 ctypes.byref(st), b2d.INFLATE_CRC32, ctypes.byref(par))
        assert r == 0 and st.value == 0 and ol.value == n and ic.value == m
    for _ in range(max(3, args.warmup)):                 # (the library's scratch for the parallel decode settles in the first calls)
        host_call()
    assert crc.value == zlib.crc32(data.data) and par.value == 1 and np.array_equal(h_out.numpy()[:n], data)
    l0 = L.b2d_kernel_launches()
    t = time.perf_counter()
    for _ in range(steps):
        host_call()
        assert par.value == 1
    host_s = (time.perf_counter() - t) / steps
    env.launches += L.b2d_kernel_launches() - l0
    # the one-warp sequential decoder on a 4 MiB prefix-sized stream, for scale (what every foreign stream got before)
    small = zlib.compressobj(6, zlib.DEFLATED, -15)
    sm_comp = small.compress(data[:4 << 20].data) + small.flush()
    os.environ["B2D_STREAM_PARALLEL"] = "0"
    hs = torch.from_numpy(np.frombuffer(sm_comp, dtype=np.uint8).copy()).pin_memory()
    t = time.perf_counter()
    r = L.b2d_inflate_stream(hs.data_ptr(), len(sm_comp), h_out.data_ptr(), 4 << 20, ctypes.byref(ol), ctypes.byref(ic), ctypes.byref(crc),
                             ctypes.byref(st), b2d.INFLATE_CRC32, ctypes.byref(par))
    seq_s = time.perf_counter() - t
    os.environ.pop("B2D_STREAM_PARALLEL", None)
    assert r == 0 and st.value == 0 and par.value == 0 and ol.value == 4 << 20
    return {
        "workload": f"one raw-DEFLATE stream of {mib} MiB G_TEXT made by zlib level 6 (history carried across blocks, no index): "
                    "InflaterInputStream on a file from system gzip",
        "value": round(n / dev_s / 1e9, 3), "unit": "GB/s", "ms_per_step": round(dev_s * 1e3, 3),
        "e2e": {"value": round(n / host_s / 1e9, 3), "unit": "GB/s", "ms_per_step": round(host_s * 1e3, 3),
                "h2d_bytes_per_step": m, "d2h_bytes_per_step": n + 40 + 4 * (mib + 1)},
        "compressed_bytes": m, "units_decoded_in_parallel": int(res[4]) & 0xFFFFFFFF,
        "sequential_one_warp_GBps": round((4 << 20) / seq_s / 1e9, 4),
        "cpu_zlib_single_thread_GBps": round(n / t_zlib / 1e9, 4),
        "note": "b2d_inflate_stream[_dev]: block starts found by a header-plausibility scan, one warp per found start, "
                "back-references into the unknown window as markers, windows by a parallel prefix scan; the sequential decoder "
                "(one warp, the number next to it) takes over for anything unusual",
    }


def run_reference(args, cores, n_members):
    """The reference's own CPU path on the host cores.  No JVM exists in this image, so the
    reference's Java cannot run; the arm is the oracle port (C restatement of decomp/Open.java), all host threads, on
    the b2d arm's workload: every step decodes all members of one GPU's batch.  Only oracle/ is loaded here."""
    from oracle import oracle as O
    O.build()
    OL = O.lib()
    pool = cf.ThreadPoolExecutor(cores)
    members, raw = make_members(OL, n_members, SEED, pool)
    for _ in range(min(args.warmup, 1)):
        cpu_inflate(O, members, cores)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_inflate(O, members, cores)
    step_s = t / args.steps
    val = n_members * MEMBER_BYTES / step_s / 1e9
    sample = (f"oracle_inflate (C restatement of decomp/Open.java; JVM absent) + zlib crc32 vs the gzip trailer over all {n_members} "
              f"members of one GPU's batch per step, {cores} threads"
              + (f"; the b2d arm at {args.gpus} GPUs decodes {args.gpus} such batches per step (throughput is what is compared)" if args.gpus > 1 else ""))
    emit({
        "impl": "reference", "metric": METRIC, "value": round(val, 4), "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(step_s * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(n_members, args.size_mib)},
        "cpu_baseline": {"value": round(val, 4), "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(val, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0})
    return 0


if __name__ == "__main__":
    sys.exit(main())
