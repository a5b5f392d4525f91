"""Per-symbol latency of the inflate loop: literal-only streams (search = LITERAL) at low occupancy."""
import ctypes, sys
sys.path.insert(0, '.')
import numpy as np, torch
import b2d_loader
b2d = b2d_loader.load(); b2d.init(0); L = b2d.lib()
dev = torch.device('cuda'); sp = ctypes.c_void_p(0)
CH = 1 << 20
for nc, search in ((148, 1), (1024, 1), (4096, 1), (148, 0), (1024, 0), (4096, 0)):
    n = nc * CH
    data = np.concatenate([b2d.corpus('text', 7 + k, 64 << 20) for k in range((n + (64 << 20) - 1) >> 26)])[:n].copy()
    d_in = torch.from_numpy(data).to(dev)
    bound = b2d.deflate_bound(n, CH)
    d_out = torch.empty(bound, dtype=torch.uint8, device=dev)
    d_total = torch.zeros(1, dtype=torch.int64, device=dev)
    d_clen = torch.zeros(nc, dtype=torch.int64, device=dev); d_crc = torch.zeros(nc, dtype=torch.int32, device=dev)
    opts = b2d.make_opts(search=search, mode=3)
    r = L.b2d_deflate_chunks_dev(d_in.data_ptr(), n, ctypes.byref(opts), d_out.data_ptr(), bound, d_total.data_ptr(), d_clen.data_ptr(), d_crc.data_ptr(), sp)
    assert r == 0, (r, b2d.lib().b2d_last_error())
    torch.cuda.synchronize()
    coff = torch.zeros(nc + 1, dtype=torch.int64, device=dev); coff[1:] = torch.cumsum(d_clen, 0)
    ooff = torch.arange(nc + 1, dtype=torch.int64, device=dev) * CH
    d_dec = torch.zeros(n, dtype=torch.uint8, device=dev)
    ol = torch.zeros(nc, dtype=torch.int64, device=dev); ic = torch.zeros_like(ol); st = torch.zeros(nc, dtype=torch.int32, device=dev); c2 = torch.zeros_like(st)
    def inflate():
        assert L.b2d_inflate_batch_dev(d_out.data_ptr(), coff.data_ptr(), nc, d_dec.data_ptr(), ooff.data_ptr(), ol.data_ptr(), ic.data_ptr(), c2.data_ptr(), st.data_ptr(), 2, sp) == 0
    for _ in range(2): inflate()
    torch.cuda.synchronize()
    assert int(st.abs().sum()) == 0 and torch.equal(d_dec, d_in)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [inflate() for _ in range(3)]; e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 3
    print(f"units={nc:5d} search={search} ratio={n/int(d_total.item()):.2f} inflate {t:7.2f} ms  {n/t/1e6:7.2f} GB/s  ns/byte/warp={t*1e6/CH:.1f}", flush=True)
