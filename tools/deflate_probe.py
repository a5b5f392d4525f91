#!/usr/bin/env python3
"""The deflate pipeline alone on cuda:0: 1 GiB (or --mib) of G_MIXED / G_TEXT resident in HBM, N timed calls of
b2d_deflate_chunks_dev.  For A/B runs of kernel changes and as the command ncu captures.
usage: tools/deflate_probe.py [--mib 1024] [--kind mixed|text] [--steps 5] [--split 0] [--depth 0]"""
import argparse
import concurrent.futures as cf
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import b2d_loader

ap = argparse.ArgumentParser()
ap.add_argument("--mib", type=int, default=1024)
ap.add_argument("--kind", default="mixed")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--split", type=int, default=0)
ap.add_argument("--crc", type=int, default=0, help="1: CRC-32 per chunk as well (config 3)")
ap.add_argument("--depth", type=int, default=0, help="chain depth (0 = the default)")
a = ap.parse_args()
b2d = b2d_loader.load()
b2d.init(0)
L = b2d.lib()
n = a.mib << 20
data = np.empty(n, np.uint8)
piece = 16 << 20
with cf.ThreadPoolExecutor(os.cpu_count()) as ex:
    list(ex.map(lambda k: getattr(L, "b2d_corpus_" + a.kind)(0xDEF1A7E + k, data[k * piece:].ctypes.data, min(piece, n - k * piece)),
                range((n + piece - 1) // piece)))
d_in = torch.from_numpy(data).cuda()
bound = b2d.deflate_bound(n, 1 << 20)
d_out = torch.empty(bound, dtype=torch.uint8, device="cuda")
d_tot = torch.zeros(1, dtype=torch.int64, device="cuda")
d_cl = torch.zeros(n >> 20, dtype=torch.int64, device="cuda")
d_crc = torch.zeros(n >> 20, dtype=torch.int32, device="cuda")
opts = b2d.make_opts(chunk_bytes=1 << 20, block_bytes=1 << 16, split_min_bytes=a.split, chain_depth=a.depth)


def run():
    r = L.b2d_deflate_chunks_dev(d_in.data_ptr(), n, ctypes.byref(opts), d_out.data_ptr(), bound, d_tot.data_ptr(), d_cl.data_ptr(), ctypes.c_void_p(d_crc.data_ptr()) if a.crc else None, None)
    assert r == 0, r


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
print(f"deflate {a.kind} {a.mib} MiB depth {a.depth}: {ms:.3f} ms/call = {n / ms / 1e6:.2f} GB/s, out {int(d_tot.item())} bytes (ratio {n / int(d_tot.item()):.4f})")
