"""BASELINE configs[3] on one GPU: 8 GiB chunked gzip-style compress + decompress through the host-pointer C ABI
(exercises the > 4 GiB paths: 64-bit offsets, sliced pipeline, 8192-chunk index)."""
import concurrent.futures as cf, ctypes, sys, time, zlib
sys.path.insert(0, '.')
import numpy as np
import b2d_loader
b2d = b2d_loader.load(); b2d.init(0); L = b2d.lib()
gib = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = gib << 30
CH = 1 << 20
data = np.empty(n, np.uint8)
def gen(k): L.b2d_corpus_mixed(0xDEF1A7E + k, data[k << 26:].ctypes.data, 64 << 20)
with cf.ThreadPoolExecutor(16) as ex: list(ex.map(gen, range(n >> 26)))
bound = b2d.deflate_bound(n, CH)
p_in = L.b2d_alloc_pinned(n); p_out = L.b2d_alloc_pinned(bound); p_back = L.b2d_alloc_pinned(n)
assert p_in and p_out and p_back
ctypes.memmove(p_in, data.ctypes.data, n)
nc = n // CH
idx = np.zeros(nc, np.uint64)
opts = b2d.make_opts()
for it in range(2):
    crc = ctypes.c_uint32(0)
    t = time.perf_counter()
    r = L.b2d_deflate_chunks(p_in, n, ctypes.byref(opts), p_out, bound, ctypes.byref(crc), idx.ctypes.data)
    td = time.perf_counter() - t
    assert r > 0, r
print(f"deflate {gib} GiB: {r} bytes (ratio {n / r:.3f}) in {td * 1e3:.1f} ms = {n / td / 1e9:.2f} GB/s e2e; chunks {nc}; index sum ok {int(idx.sum()) == r}")
t = time.perf_counter(); c = 0
for k in range(0, n, 1 << 30): c = zlib.crc32(data[k:k + (1 << 30)].data, c)
print(f"crc ok {c == crc.value} (zlib crc32 took {time.perf_counter() - t:.1f} s)")
in_off = np.zeros(nc + 1, np.uint64); in_off[1:] = np.cumsum(idx)
out_off = np.arange(nc + 1, dtype=np.uint64) * np.uint64(CH)
ol = np.zeros(nc, np.uint64); ic = np.zeros(nc, np.uint64); cr = np.zeros(nc, np.uint32); st = np.zeros(nc, np.int32)
for it in range(2):
    t = time.perf_counter()
    rr = L.b2d_inflate_batch(p_out, in_off.ctypes.data, nc, p_back, out_off.ctypes.data, ol.ctypes.data, ic.ctypes.data, cr.ctypes.data, st.ctypes.data, 3)
    ti = time.perf_counter() - t
    assert rr == 0 and not st.any()
back = np.ctypeslib.as_array((ctypes.c_uint8 * n).from_address(p_back))
print(f"inflate: {ti * 1e3:.1f} ms = {n / ti / 1e9:.2f} GB/s e2e; identical {np.array_equal(back, data)}")
whole = np.ctypeslib.as_array((ctypes.c_uint8 * r).from_address(p_out))
d = zlib.decompressobj(-15); got = d.decompress(whole[:int(in_off[3])].tobytes())
print("zlib reads the first 3 chunks:", got == data[:3 << 20].tobytes())
