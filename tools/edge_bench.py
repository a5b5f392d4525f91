"""Device-resident throughput of deflate + inflate on the BASELINE configs[4] edge corpora (and text / mixed)."""
import ctypes, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import b2d_loader
b2d = b2d_loader.load(); b2d.init(0); L = b2d.lib()
n = int(sys.argv[1]) << 20 if len(sys.argv) > 1 else 1 << 30
CH = (int(sys.argv[2]) << 10) if len(sys.argv) > 2 else (1 << 20)
dev = torch.device('cuda')
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def run(kind, mode=0):
    if kind == 'zeros': data = np.zeros(n, np.uint8)
    else: data = np.concatenate([b2d.corpus(kind, 7 + k, 64 << 20) for k in range(n >> 26)])
    d_in = torch.from_numpy(data).to(dev)
    bound = b2d.deflate_bound(n, CH)
    d_out = torch.empty(bound, dtype=torch.uint8, device=dev)
    d_total = torch.zeros(1, dtype=torch.int64, device=dev)
    nc = n // CH
    d_clen = torch.zeros(nc, dtype=torch.int64, device=dev); d_crc = torch.zeros(nc, dtype=torch.int32, device=dev)
    opts = b2d.make_opts(mode=mode, chunk_bytes=CH)
    nb = n // 65536
    d_bits = torch.zeros(nb, dtype=torch.int32, device=dev)
    def deflate():
        assert L.b2d_deflate_chunks_indexed_dev(d_in.data_ptr(), n, ctypes.byref(opts), d_out.data_ptr(), bound, d_total.data_ptr(), d_clen.data_ptr(), d_crc.data_ptr(), d_bits.data_ptr(), sp) == 0
    for _ in range(2): deflate()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [deflate() for _ in range(3)]; e1.record(); torch.cuda.synchronize()
    td = e0.elapsed_time(e1) / 3
    comp = int(d_total.item())
    coff = torch.zeros(nc + 1, dtype=torch.int64, device=dev); coff[1:] = torch.cumsum(d_clen, 0)
    ooff = torch.arange(nc + 1, dtype=torch.int64, device=dev) * CH
    d_dec = torch.zeros(n, dtype=torch.uint8, device=dev)
    ol = torch.zeros(nc, dtype=torch.int64, device=dev); ic = torch.zeros_like(ol); st = torch.zeros(nc, dtype=torch.int32, device=dev); c2 = torch.zeros_like(st)
    def inflate():
        assert L.b2d_inflate_batch_dev(d_out.data_ptr(), coff.data_ptr(), nc, d_dec.data_ptr(), ooff.data_ptr(), ol.data_ptr(), ic.data_ptr(), c2.data_ptr(), st.data_ptr(), 3, sp) == 0
    for _ in range(2): inflate()
    torch.cuda.synchronize()
    assert int(st.abs().sum()) == 0 and torch.equal(d_dec, d_in)
    e0.record(); [inflate() for _ in range(3)]; e1.record(); torch.cuda.synchronize()
    ti = e0.elapsed_time(e1) / 3
    d_dec.zero_(); cst = torch.zeros(nc, dtype=torch.int32, device=dev)
    def inflate_blocks():
        assert L.b2d_inflate_chunks_dev(d_out.data_ptr(), coff.data_ptr(), nc, d_bits.data_ptr(), CH, 65536, n, d_dec.data_ptr(), c2.data_ptr(), cst.data_ptr(), 1, sp) == 0
    for _ in range(2): inflate_blocks()
    torch.cuda.synchronize()
    assert int(cst.abs().sum()) == 0 and torch.equal(d_dec, d_in) and torch.equal(c2, d_crc)
    e0.record(); [inflate_blocks() for _ in range(3)]; e1.record(); torch.cuda.synchronize()
    tb = e0.elapsed_time(e1) / 3
    print(f"{kind:7s} mode={mode} ratio {n / comp:9.2f}  deflate {td:8.2f} ms = {n / td / 1e6:7.2f} GB/s   inflate ({nc} x 1 MiB chunks) {ti:8.2f} ms = {n / ti / 1e6:7.2f} GB/s   block-indexed {tb:8.2f} ms = {n / tb / 1e6:7.2f} GB/s", flush=True)
for kind, mode in ((('mixed', 0), ('text', 0)) if len(sys.argv) > 2 else (('mixed', 0), ('text', 0), ('random', 0), ('zeros', 0), ('text', 2))):
    run(kind, mode)
