"""Small end-to-end calls of every host entry point, meant to run under compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b2d_loader

b2d = b2d_loader.load()
b2d.init(0)
text = b2d.corpus("text", 1, 3 << 20).tobytes()
mixed = b2d.corpus("mixed", 2, (2 << 20) + 777).tobytes()
# deflate (chunked, indexed, split, reference framing) + both decoders of our own streams
for data in (text[:300000], mixed):
    comp, crc, idx, bits = b2d.deflate_chunks_indexed(data, b2d.make_opts(), crc=0)
    assert zlib.decompress(bytes(comp), -15) == data and crc == zlib.crc32(data)
    out, crcs, st = b2d.inflate_chunks(comp, idx, bits, len(data))
    assert not st.any() and out.tobytes() == data
    comp2 = b2d.deflate_chunks(data, b2d.make_opts(split_min_bytes=8192))
    assert zlib.decompress(bytes(comp2), -15) == data
    comp3 = b2d.deflate_chunks(data[:200000], b2d.make_opts(framing=b2d.FRAMING_REFERENCE, search=b2d.SEARCH_RLE, lazy=0, mode=b2d.MODE_DYNAMIC))
    assert zlib.decompress(bytes(comp3), -15) == data[:200000]
# batch inflate: pageable, pinned (progress + streaming input), ragged (mirror)
members = []
for i in range(80):
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    members.append(c.compress(text[i * 30000:(i + 1) * 30000 + 5000]) + c.flush())
ref = b2d.inflate_batch(members, 40000, b2d.INFLATE_CRC32)
buf = b2d.PinnedBuffer(40000 * len(members))
got = b2d.inflate_batch(members, 40000, b2d.INFLATE_CRC32, out=buf.array, pinned_in=True)
assert got[0] == ref[0] and not got[4].any()
caps = [35001 + 13 * i for i in range(len(members))]
buf2 = b2d.PinnedBuffer(sum(caps))
got = b2d.inflate_batch(members, caps, b2d.INFLATE_CRC32, out=buf2.array, pinned_in=True)
assert got[0] == ref[0]
# one foreign stream, parallel
c = zlib.compressobj(6, zlib.DEFLATED, -15)
stream = c.compress(text) + c.flush()
out, consumed, crc, st, par = b2d.inflate_stream(stream, len(text) + 10)
assert st == 0 and par == 1 and out.tobytes() == text and consumed == len(stream)
out, consumed, crc, st, par = b2d.inflate_stream(stream[:len(stream) // 2], len(text) + 10)
assert st == 1
print("sanitize_small ok")
