#!/usr/bin/env python3
"""ONE zlib-made DEFLATE stream (history carried across blocks, like a file from system gzip) through the speculative
parallel decoder: device-resident time per call, host-pointer time, next to zlib and the oracle on one host thread.
usage: tools/stream_probe.py [--mib 256] [--kind text|mixed] [--level 6] [--steps 5]"""
import argparse
import ctypes
import os
import sys
import time
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import b2d_loader

ap = argparse.ArgumentParser()
ap.add_argument("--mib", type=int, default=256)
ap.add_argument("--kind", default="text")
ap.add_argument("--level", type=int, default=6)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
b2d = b2d_loader.load()
b2d.init(0)
L = b2d.lib()
n = a.mib << 20
data = np.concatenate([b2d.corpus(a.kind, 0xDEF1A7E + k, min(16 << 20, n - (k << 24))) for k in range((n + (16 << 20) - 1) >> 24)])
t = time.perf_counter()
c = zlib.compressobj(a.level, zlib.DEFLATED, -15)
comp = c.compress(data.data) + c.flush()
t_comp = time.perf_counter() - t
t = time.perf_counter()
assert zlib.decompress(comp, -15) == data.tobytes()
t_zlib = time.perf_counter() - t
print(f"{a.kind} {a.mib} MiB, zlib level {a.level}: {len(comp)} bytes (ratio {n / len(comp):.3f}); zlib inflate on one thread {n / t_zlib / 1e9:.3f} GB/s")
h_in = torch.from_numpy(np.frombuffer(comp, dtype=np.uint8).copy())
d_in = torch.zeros(len(comp) + 64, dtype=torch.uint8, device="cuda")
d_in[:len(comp)] = h_in.cuda()
d_out = torch.zeros(n + 256, dtype=torch.uint8, device="cuda")
d_res = torch.zeros(8, dtype=torch.int64, device="cuda")


def run():
    r = L.b2d_inflate_stream_dev(d_in.data_ptr(), len(comp), d_out.data_ptr(), n, d_res.data_ptr(), None)
    assert r == 0, r


for _ in range(2):
    run()
torch.cuda.synchronize()
res = d_res.cpu().numpy()
print("result: out_len", int(res[0]), "consumed", int(res[1]), "status", int(res[2]), "units", int(res[4]) & 0xFFFFFFFF)
if int(res[2]) == 0:
    assert int(res[0]) == n and int(res[1]) == len(comp)
    assert torch.equal(d_out[:n], torch.from_numpy(data).cuda())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(f"b2d_inflate_stream_dev: {ms:.3f} ms/call = {n / ms / 1e6:.2f} GB/s (device-resident)")
else:
    print("the default unit buffers are too small for this stream (the host entry point retries with larger ones)")
del d_out, d_in
torch.cuda.empty_cache()
# host pointers, pinned
pin_in = b2d.PinnedBuffer(len(comp) + 64)
pin_in.array[:len(comp)] = np.frombuffer(comp, dtype=np.uint8)
pin_out = b2d.PinnedBuffer(n + 64)
ol, ic = ctypes.c_uint64(0), ctypes.c_uint64(0)
crc, st, par = ctypes.c_uint32(0), ctypes.c_int32(0), ctypes.c_int32(0)


def host(flags):
    r = L.b2d_inflate_stream(pin_in.array.ctypes.data, len(comp), pin_out.array.ctypes.data, n, ctypes.byref(ol), ctypes.byref(ic),
                             ctypes.byref(crc), ctypes.byref(st), flags, ctypes.byref(par))
    assert r == 0 and st.value == 0 and ol.value == n, (r, st.value, ol.value)


host(1)
assert crc.value == zlib.crc32(data.data)
t = time.perf_counter()
for _ in range(a.steps):
    host(1)
dt = (time.perf_counter() - t) / a.steps
print(f"b2d_inflate_stream (host pointers, pinned, + CRC-32): {dt * 1e3:.3f} ms/call = {n / dt / 1e9:.2f} GB/s (parallel = {par.value})")
assert np.array_equal(pin_out.array[:n], data)
os.environ["B2D_STREAM_PARALLEL"] = "0"
if a.mib <= 64:
    t = time.perf_counter()
    r = L.b2d_inflate_stream(pin_in.array.ctypes.data, len(comp), pin_out.array.ctypes.data, n, ctypes.byref(ol), ctypes.byref(ic),
                             ctypes.byref(crc), ctypes.byref(st), 1, ctypes.byref(par))
    dt = time.perf_counter() - t
    print(f"sequential one-warp decoder for comparison: {dt * 1e3:.1f} ms = {n / dt / 1e9:.4f} GB/s (parallel = {par.value})")
