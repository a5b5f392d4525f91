"""GPU parity: b2d_inflate_stream -- ONE raw-DEFLATE stream of any origin, decoded speculatively in parallel -- against
zlib and the oracle (decomp/Open.java restated).  Whatever path produced the result (parallel, or the sequential
fallback), bytes, consumed input and status must be the sequential decoder's."""
import os
import random
import zlib

import numpy as np
import pytest

from util import zlib_raw

pytestmark = pytest.mark.gpu


def _text(b2d, seed, n):
    return b2d.corpus("text", seed, n).tobytes()


@pytest.mark.parametrize("level", [1, 6, 9])
def test_zlib_made_stream_decodes_in_parallel(b2d, oracle, level):
    n = 24 << 20
    data = _text(b2d, 4711 + level, n)
    comp = zlib_raw(data, level)
    out, consumed, crc, st, par = b2d.inflate_stream(comp, n + 1000)
    assert st == 0 and par == 1, (st, par)
    assert consumed == len(comp) and crc == zlib.crc32(data)
    assert out.tobytes() == data
    # exact capacity, and one byte too few (the sequential decoder reports the overflow after delivering what fits)
    out, consumed, crc, st, par = b2d.inflate_stream(comp, n)
    assert st == 0 and out.tobytes() == data
    out, consumed, crc, st, par = b2d.inflate_stream(comp, n - 1)
    assert st == b2d.ERR_OUTPUT_OVERFLOW and par == 0 and out.tobytes() == data[:n - 1]


def test_mixed_content_and_trailing_bytes(b2d, oracle):
    """Text, zeros, random bytes (stored blocks) and long-range repeats in one stream; bytes after the final block are
    not consumed (Open.finish, Open.java:113-124)."""
    data = b2d.corpus("mixed", 99, 40 << 20).tobytes()
    comp = zlib_raw(data, 6)
    out, consumed, crc, st, par = b2d.inflate_stream(comp + b"trailing bytes", len(data) + 64)
    assert st == 0 and consumed == len(comp) and out.tobytes() == data and crc == zlib.crc32(data)
    # the oracle agrees on a prefix-sized sample of the same stream (status, consumed)
    small = zlib_raw(data[:3 << 20], 6)
    o_st, o_out, o_cons = oracle.inflate(small, out_cap=(3 << 20) + 8)
    out, consumed, crc, st, par = b2d.inflate_stream(small, (3 << 20) + 8)
    assert (st, consumed, out.tobytes()) == (o_st, o_cons, o_out)


def test_streams_of_other_encoders(b2d, oracle):
    """The reference's own encoder (oracle restatement, RLE_DYNAMIC and FULL_DYNAMIC: history carried across blocks), our
    GPU encoder's chunked stream, fixed-Huffman-only and stored-only streams (nothing for the block finder: sequential)."""
    rng = random.Random(8)
    data = _text(b2d, 5, 6 << 20)
    for comp in (oracle.deflate(data, (oracle.RLE_DYNAMIC,)), bytes(b2d.deflate_chunks(data, b2d.make_opts())),
                 zlib_raw(data, 6, zlib.Z_FIXED), zlib_raw(rng.randbytes(3 << 20), 0)):
        want, used = zlib.decompressobj(-15), None
        ref = want.decompress(comp)
        out, consumed, crc, st, par = b2d.inflate_stream(comp, len(ref) + 10)
        assert st == 0 and out.tobytes() == ref and consumed == len(comp) - len(want.unused_data)


def test_corrupt_and_truncated_streams_report_like_the_sequential_decoder(b2d, oracle):
    data = _text(b2d, 77, 8 << 20)
    comp = zlib_raw(data, 6)
    rng = random.Random(3)
    cases = [comp[:len(comp) // 2], comp[:-1], comp[:1 << 20] + rng.randbytes(1 << 20) + comp[2 << 20:]]
    for _ in range(3):
        b = bytearray(comp)
        b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
        cases.append(bytes(b))
    for c in cases:
        o_st, o_out, o_cons = oracle.inflate(c, out_cap=len(data) + 100)
        out, consumed, crc, st, par = b2d.inflate_stream(c, len(data) + 100)
        assert st == o_st, (b2d.status_name(st), oracle.status_name(o_st))
        assert out.tobytes() == o_out
        if o_st == 0:
            assert consumed == o_cons


def test_segment_and_buffer_knobs(b2d):
    """Small segments / tight unit buffers change how much is decoded in parallel, never the result."""
    data = _text(b2d, 123, 12 << 20)
    comp = zlib_raw(data, 6)
    ref = None
    for seg, cap in ((16384, 0), (32768, 1 << 18), (262144, 0)):
        os.environ["B2D_STREAM_SEGMENT"] = str(seg)
        if cap:
            os.environ["B2D_STREAM_UNIT_CAP"] = str(cap)
        # (the knobs are read once per process: this mostly checks that whatever is in force gives the same bytes)
        out, consumed, crc, st, par = b2d.inflate_stream(comp, len(data) + 5)
        assert st == 0 and out.tobytes() == data and consumed == len(comp)
    os.environ.pop("B2D_STREAM_SEGMENT", None)
    os.environ.pop("B2D_STREAM_UNIT_CAP", None)


def test_long_stretch_without_a_dynamic_block(b2d, oracle):
    """Three MiB of random bytes in the middle of a text: zlib stores them (48 stored blocks, nothing for the block finder
    to restart at).  The unit that runs into the stretch moves on through spare units instead of outgrowing its buffer,
    so the stream is still decoded by the parallel path."""
    rng = random.Random(12)
    data = _text(b2d, 31, 2 << 20) + rng.randbytes(3 << 20) + _text(b2d, 32, 2 << 20)
    comp = zlib_raw(data, 6)
    out, consumed, crc, st, par = b2d.inflate_stream(comp, len(data) + 16)
    assert st == 0 and par == 1 and consumed == len(comp)
    assert out.tobytes() == data and crc == zlib.crc32(data)
