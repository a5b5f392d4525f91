"""Device-side time of the member decoder on BASELINE config 2 (4096 x 256 KiB zlib-6 members of G_TEXT), nothing else:
the quick A/B measurement for kernel variants.  B2D_SO=<path> selects the library."""
import ctypes, os, sys, time, zlib
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import b2d_loader
b2d = b2d_loader.load()
if os.environ.get("B2D_SO"):                      # another build of the library (any age: only the calls below are bound)
    L = ctypes.CDLL(os.environ["B2D_SO"])
    assert L.b2d_init(0) == 0
    vp = ctypes.c_void_p
    L.b2d_inflate_batch_dev.argtypes = [vp, vp, ctypes.c_uint32, vp, vp, vp, vp, vp, vp, ctypes.c_uint32, vp]
else:
    b2d.init(0); L = b2d.lib()
N, SZ = 4096, 256 << 10
raw = np.concatenate([b2d.corpus('text', 0xDEF1A7E + k, 64 << 20) for k in range(N * SZ >> 26)])
def comp(i):
    c = zlib.compressobj(6, zlib.DEFLATED, -15); return c.compress(raw[i * SZ:(i + 1) * SZ].tobytes()) + c.flush()
with ThreadPoolExecutor(16) as ex: members = list(ex.map(comp, range(N)))
in_off = np.zeros(N + 1, np.uint64); in_off[1:] = np.cumsum([len(m) for m in members])
blob = np.frombuffer(b"".join(members), np.uint8)
dev = torch.device('cuda')
d_in = torch.from_numpy(np.concatenate([blob, np.zeros(64, np.uint8)])).to(dev)
d_ioff = torch.from_numpy(in_off.astype(np.int64)).to(dev)
d_ooff = (torch.arange(N + 1, dtype=torch.int64, device=dev) * SZ)
d_out = torch.zeros(N * SZ, dtype=torch.uint8, device=dev)
ol = torch.zeros(N, dtype=torch.int64, device=dev); ic = torch.zeros_like(ol)
st = torch.zeros(N, dtype=torch.int32, device=dev); crc = torch.zeros_like(st)
def run(n=N): assert L.b2d_inflate_batch_dev(d_in.data_ptr(), d_ioff.data_ptr(), n, d_out.data_ptr(), d_ooff.data_ptr(), ol.data_ptr(), ic.data_ptr(), crc.data_ptr(), st.data_ptr(), 0, ctypes.c_void_p(0)) == 0
for _ in range(3): run()
torch.cuda.synchronize()
assert int(st.abs().sum()) == 0 and torch.equal(d_out.cpu(), torch.from_numpy(raw))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
res = []
for n in (N, 1024, 148):          # one wave of 28 warps per SM; 7 warps per SM; one warp per SM (pure latency)
    e0.record(); [run(n) for _ in range(10)]; e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 10
    res.append(f"{n} members {t:7.3f} ms {n * SZ / t / 1e6:6.2f} GB/s")
print(f"{os.path.basename(os.environ.get('B2D_SO', 'default')):28s} " + " | ".join(res), flush=True)
