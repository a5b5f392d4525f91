"""Ratio and device throughput of the deflate pipeline against the DEFLATE block size (b2d_deflate_opts.block_bytes),
on 1 GiB of G_MIXED and G_TEXT in 1 MiB chunks; the stream is decoded back (one warp per chunk, and block-parallel
through the block index) and compared.  The SURVEY 8f N3 question in numbers: how much of BinarySplit's gain does a
smaller fixed block already give, and what does it cost."""
import ctypes, sys
sys.path.insert(0, '.')
import numpy as np, torch
import b2d_loader
b2d = b2d_loader.load(); b2d.init(0); L = b2d.lib()
n = int(sys.argv[1]) << 20 if len(sys.argv) > 1 else 1 << 30
CH = 1 << 20
dev = torch.device('cuda')
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for kind in ('mixed', 'text'):
    seed0 = 0xDEF1A7E if kind == 'mixed' else 7
    data = np.concatenate([b2d.corpus(kind, seed0 + k, 64 << 20) for k in range(n >> 26)])
    d_in = torch.from_numpy(data).to(dev)
    bound = b2d.deflate_bound(n, CH) + (n >> 5)
    d_out = torch.empty(bound, dtype=torch.uint8, device=dev)
    d_total = torch.zeros(1, dtype=torch.int64, device=dev)
    nc = n // CH
    d_clen = torch.zeros(nc, dtype=torch.int64, device=dev); d_crc = torch.zeros(nc, dtype=torch.int32, device=dev)
    for bb, leaf in ((65536, 0), (32768, 0), (16384, 0), (8192, 0), (65536, 32768), (65536, 16384), (65536, 8192), (65536, 4096)):
        opts = b2d.make_opts(chunk_bytes=CH, block_bytes=bb, split_min_bytes=leaf)
        d_bits = torch.zeros(n // bb, dtype=torch.int32, device=dev)
        def deflate():
            r = L.b2d_deflate_chunks_indexed_dev(d_in.data_ptr(), n, ctypes.byref(opts), d_out.data_ptr(), bound, d_total.data_ptr(), d_clen.data_ptr(), d_crc.data_ptr(), d_bits.data_ptr(), sp)
            assert r == 0, (r, L.b2d_last_error())
        for _ in range(2): deflate()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); [deflate() for _ in range(3)]; e1.record(); torch.cuda.synchronize()
        td = e0.elapsed_time(e1) / 3
        comp = int(d_total.item())
        coff = torch.zeros(nc + 1, dtype=torch.int64, device=dev); coff[1:] = torch.cumsum(d_clen, 0)
        d_dec = torch.zeros(n, dtype=torch.uint8, device=dev)
        c2 = torch.zeros(nc, dtype=torch.int32, device=dev); cst = torch.zeros(nc, dtype=torch.int32, device=dev)
        def inflate_blocks():
            assert L.b2d_inflate_chunks_dev(d_out.data_ptr(), coff.data_ptr(), nc, d_bits.data_ptr(), CH, bb, n, d_dec.data_ptr(), c2.data_ptr(), cst.data_ptr(), 1, sp) == 0
        for _ in range(2): inflate_blocks()
        torch.cuda.synchronize()
        assert int(cst.abs().sum()) == 0 and torch.equal(d_dec, d_in) and torch.equal(c2, d_crc)
        e0.record(); [inflate_blocks() for _ in range(3)]; e1.record(); torch.cuda.synchronize()
        tb = e0.elapsed_time(e1) / 3
        print(f"{kind:6s} block {bb >> 10:3d} KiB{f' split to {leaf >> 10:2d} KiB' if leaf else '                ':s}: {comp:11d} bytes, ratio {n / comp:7.4f}  deflate {td:7.2f} ms = {n / td / 1e6:6.2f} GB/s   block-parallel inflate {tb:7.2f} ms = {n / tb / 1e6:6.2f} GB/s", flush=True)
