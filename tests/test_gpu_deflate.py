"""GPU parity: b2d_deflate_chunks (through the C ABI, host pointers).  Every stream must decode through the oracle
(restating the reference decoder, decomp/Open.java) AND zlib back to the exact input; with reference framing and the
reference's own greedy searches the bytes must equal the oracle encoder's (Lz77Huffman.java restated); the default
search must land within 1 % of the reference's FULL_DYNAMIC at the same chunking."""
import random
import zlib

import numpy as np
import pytest

from util import zlib_inflate_raw

pytestmark = pytest.mark.gpu


def _decode_both(oracle, comp, n):
    comp = bytes(comp)
    st, out, consumed = oracle.inflate(comp, out_cap=n + 8)
    assert st == 0, oracle.status_name(st)
    assert consumed == len(comp)
    zout, zused = zlib_inflate_raw(comp)
    assert zused == len(comp)
    assert zout == out
    return out


def _text(rng, n):
    words = [bytes(rng.choices(b"etaoinshrdlucmfw", k=rng.randrange(1, 10))) for _ in range(700)]
    out = bytearray()
    while len(out) < n:
        out += rng.choice(words) + b" "
    return bytes(out[:n])


SIZES = [0, 1, 2, 3, 4, 5, 100, 4095, 4096, 65535, 65536, 65537, 200000, (1 << 20) - 1, 1 << 20, (1 << 20) + 1,
         3 * (1 << 20) + 12345]


@pytest.mark.parametrize("n", SIZES)
def test_roundtrip_default(b2d, oracle, n):
    rng = random.Random(n)
    data = _text(rng, n)
    comp, crc, idx = b2d.deflate_chunks(data, b2d.make_opts(), crc=0)
    assert _decode_both(oracle, comp, n) == data
    assert crc == zlib.crc32(data)
    if n:
        assert int(idx.sum()) == len(comp)
    if n >= 4096:
        assert len(comp) < n * 0.6


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
@pytest.mark.parametrize("search", [0, 1, 2, 3])
def test_modes_and_searches(b2d, oracle, mode, search):
    rng = random.Random(mode * 10 + search)
    data = (_text(rng, 150000) + rng.randbytes(70000) + bytes(80000) + _text(rng, 30000) * 3)
    opts = b2d.make_opts(mode=mode, search=search, chunk_bytes=1 << 17, block_bytes=1 << 15)
    comp = b2d.deflate_chunks(data, opts)
    assert _decode_both(oracle, comp, len(data)) == data


def test_edge_corpora(b2d, oracle):
    """Config-5 shapes at test size: incompressible bytes -> stored blocks; zeros -> 258/dist-1 runs; fixed only."""
    n = 3 << 20
    rnd = b2d.corpus("random", 0xDEF1A7E, n).tobytes()
    comp = b2d.deflate_chunks(rnd, b2d.make_opts())
    assert _decode_both(oracle, comp, n) == rnd
    # a 64 KiB block is two stored pieces (<= 65535 bytes each, Uncompressed.java:51) = 10 bytes; + 5 per chunk marker
    assert len(comp) == n + 10 * (n >> 16) + 5 * (n >> 20)
    zeros = bytes(n)
    comp = b2d.deflate_chunks(zeros, b2d.make_opts())
    assert _decode_both(oracle, comp, n) == zeros
    assert len(comp) < n // 500
    text = b2d.corpus("text", 0xDEF1A7E, n).tobytes()
    comp = b2d.deflate_chunks(text, b2d.make_opts(mode=b2d.MODE_FIXED))
    assert _decode_both(oracle, comp, n) == text
    # fixed-only: every block header is BTYPE=01 -> first 3 bits are 0,1,0 (BFINAL=0, type 1 LSB first)
    assert comp[0] & 7 == 0b010


@pytest.mark.parametrize("search,strat", [(1, "LITERAL"), (2, "RLE"), (3, "FULL")])
@pytest.mark.parametrize("dynamic", [False, True])
def test_reference_framing_is_byte_identical_to_oracle(b2d, oracle, search, strat, dynamic):
    """framing=REFERENCE + the reference's greedy searches = the bytes DeflaterOutputStream would write
    (oracle restatement of Lz77Huffman.java:42-288 framed per DeflaterOutputStream.java:119-137)."""
    rng = random.Random(search)
    datas = [b"", b"A", b"abc" * 1000, bytes(1000), bytes(range(256)), b"a" * 10 + b"b" * 10,
             _text(rng, 200000), rng.randbytes(3000) * 30,
             b"".join(bytes([rng.randrange(256)]) * rng.randrange(1, 600) for _ in range(400))]
    sid = getattr(oracle, f"{strat}_{'DYNAMIC' if dynamic else 'STATIC'}")
    for d in datas:
        opts = b2d.make_opts(mode=b2d.MODE_DYNAMIC if dynamic else b2d.MODE_FIXED, search=search, lazy=0,
                             framing=b2d.FRAMING_REFERENCE)
        comp = bytes(b2d.deflate_chunks(d, opts))
        want = oracle.deflate(d, (sid,))
        assert comp == want, (strat, dynamic, len(d), len(comp), len(want))


def test_reference_framing_multistrategy(b2d, oracle):
    rng = random.Random(77)
    d = _text(rng, 100000) + rng.randbytes(140000) + bytes(70000)
    opts = b2d.make_opts(mode=b2d.MODE_AUTO, search=b2d.SEARCH_FULL, lazy=0, framing=b2d.FRAMING_REFERENCE)
    comp = bytes(b2d.deflate_chunks(d, opts))
    want = oracle.deflate(d, (oracle.UNCOMPRESSED, oracle.FULL_STATIC, oracle.FULL_DYNAMIC))
    assert comp == want


def test_ratio_within_one_percent_of_full_dynamic(b2d, oracle):
    """north_star: compressed size within 1 % of the reference's dynamic-Huffman strategy at the same chunking
    (FULL_DYNAMIC, history reset per 1 MiB chunk, 64 KiB blocks)."""
    for kind in ("text", "mixed"):
        data = b2d.corpus(kind, 0xDEF1A7E, 4 << 20).tobytes()
        comp, crc, idx = b2d.deflate_chunks(data, b2d.make_opts(), crc=0)
        assert _decode_both(oracle, comp, len(data)) == data
        ref = 0
        for c in range(4):
            chunk = data[c << 20:(c + 1) << 20]
            ref += len(oracle.deflate(chunk, (oracle.FULL_DYNAMIC,))) + 5     # + the chunk marker we add
        assert len(comp) <= ref * 1.01, (kind, len(comp), ref)


def test_streaming_calls_concatenate(b2d, oracle):
    """is_last=0 calls followed by an is_last=1 call form ONE stream; the CRC runs across calls."""
    rng = random.Random(9)
    parts = [_text(rng, 1 << 20), rng.randbytes(300000), b"", _text(rng, 77777)]
    out, crc = b"", 0
    for i, p in enumerate(parts):
        comp, crc, _ = b2d.deflate_chunks(p, b2d.make_opts(is_last=int(i == len(parts) - 1)), crc=crc)
        out += bytes(comp)
    whole = b"".join(parts)
    assert _decode_both(oracle, out, len(whole)) == whole
    assert crc == zlib.crc32(whole)


def test_chunk_index_allows_independent_decode(b2d, oracle):
    """Each chunk of the stream decodes on its own (chunk-indexed inflate): the multi-GPU / random-access unit."""
    data = b2d.corpus("mixed", 5, (5 << 20) + 999).tobytes()
    comp, crc, idx = b2d.deflate_chunks(data, b2d.make_opts(), crc=0)
    comp = bytes(comp)
    offs = np.concatenate([[0], np.cumsum(idx)]).astype(np.uint64)
    members = [comp[int(offs[i]):int(offs[i + 1])] for i in range(len(idx))]
    outs, out_len, consumed, crcs, status = b2d.inflate_batch(members, 1 << 20, flags=b2d.INFLATE_CHUNK_INDEXED | b2d.INFLATE_CRC32)
    assert all(s == 0 for s in status), [b2d.status_name(int(s)) for s in status]
    assert b"".join(outs) == data
    c = 0
    for i, o in enumerate(outs):
        c = b2d.crc32_combine(c, int(crcs[i]), len(o))
    assert c == zlib.crc32(data) == crc


def test_crc32(b2d):
    rng = random.Random(3)
    assert b2d.crc32(b"123456789") == 0xCBF43926
    for n in (0, 1, 15, 16, 17, 4095, 65536, (1 << 20) + 3, 5_000_001):
        d = rng.randbytes(n)
        assert b2d.crc32(d) == zlib.crc32(d), n
        assert b2d.crc32(d, 0x12345678) == zlib.crc32(d, 0x12345678), n
    d = np.frombuffer(rng.randbytes(100003), np.uint8)
    assert b2d.crc32(d[3:]) == zlib.crc32(d[3:].tobytes())       # unaligned start


def _indexed_roundtrip(b2d, data, opts):
    comp, crc, sizes, bits = b2d.deflate_chunks_indexed(data, opts, crc=0)
    plain = b2d.deflate_chunks(data, opts)
    assert np.array_equal(comp, plain)                             # the same stream, just with the block index beside it
    cb, bb = opts.chunk_bytes or (1 << 20), opts.block_bytes or (1 << 16)
    out, crcs, st = b2d.inflate_chunks(comp, sizes, bits, len(data), cb, bb)
    assert not st.any(), [b2d.status_name(int(s)) for s in st[st != 0][:4]]
    assert out.tobytes() == bytes(data)
    c = 0
    for i in range(len(sizes)):
        c = b2d.crc32_combine(c, int(crcs[i]), min(cb, len(data) - i * cb))
    assert c == crc == zlib.crc32(bytes(data))
    return comp, sizes, bits


@pytest.mark.parametrize("n", [1, 100, 65535, 65536, 65537, 200000, (1 << 20) - 1, 1 << 20, (1 << 20) + 1, 5 * (1 << 20) + 4321])
def test_block_indexed_roundtrip_sizes(b2d, n):
    rng = random.Random(n)
    _indexed_roundtrip(b2d, _text(rng, n), b2d.make_opts())


@pytest.mark.parametrize("kind,mode", [("mixed", 0), ("random", 0), ("zeros", 0), ("text", 2), ("text", 1), ("mixed", 3)])
def test_block_indexed_roundtrip_corpora(b2d, kind, mode):
    n = (6 << 20) + 777
    data = bytes(n) if kind == "zeros" else b2d.corpus(kind, 0xDEF1A7E, n).tobytes()
    _indexed_roundtrip(b2d, data, b2d.make_opts(mode=mode))
    _indexed_roundtrip(b2d, data[:(1 << 20) + 99], b2d.make_opts(mode=mode, chunk_bytes=1 << 17, block_bytes=1 << 14))


def test_block_indexed_failures_fall_back_to_the_sequential_outcome(b2d, oracle):
    data = b2d.corpus("text", 11, 3 << 20).tobytes()
    comp, sizes, bits = _indexed_roundtrip(b2d, data, b2d.make_opts())
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    # a wrong block offset: the parallel decode of that chunk stumbles, the sequential re-decode does not need the index
    bad_bits = bits.copy()
    bad_bits[16 + 5] += 3
    out, crcs, st = b2d.inflate_chunks(comp, sizes, bad_bits, len(data))
    assert not st.any() and out.tobytes() == data
    # a corrupted chunk: its status and delivered bytes are what the sequential decoder reports; the others are intact
    bad = comp.copy()
    bad[int(offs[1]) + int(sizes[1]) // 2] ^= 0x55
    out, crcs, st = b2d.inflate_chunks(bad, sizes, bits, len(data))
    assert st[0] == 0 and st[2] == 0 and out[:1 << 20].tobytes() == data[:1 << 20] and out[2 << 20:].tobytes() == data[2 << 20:]
    chunk = bad[int(offs[1]):int(offs[2])].tobytes()
    ost, oout, _ = oracle.inflate(chunk, out_cap=1 << 20)
    if ost == 0 and len(oout) == 1 << 20:                              # the flip happened to stay decodable: bytes differ, CRC tells
        assert int(crcs[1]) != zlib.crc32(data[1 << 20:2 << 20])
    else:
        assert int(st[1]) != 0
        if ost > 0:
            assert int(st[1]) == ost


# ---- adaptive block splitting (the GPU counterpart of comp/BinarySplit.java) ----
def _mixed_local(rng, n):
    parts, total = [], 0
    while total < n:
        k = rng.choice((3000, 9000, 20000, 50000, 90000))
        kind = rng.randrange(4)
        p = _text(rng, k) if kind == 0 else rng.randbytes(k) if kind == 1 else bytes(k) if kind == 2 else bytes(range(256)) * (k // 256)
        parts.append(p)
        total += len(p)
    return b"".join(parts)[:n]


@pytest.mark.parametrize("leaf", [4096, 8192, 16384, 32768])
@pytest.mark.parametrize("mode", [0, 3])
def test_adaptive_split_roundtrip_and_gain(b2d, oracle, leaf, mode):
    """Every block_bytes span becomes the cheapest partition of its tree of pieces: the stream still decodes through the
    reference decoder and zlib, the block index still works (one entry per span), and on data whose statistics change
    inside 64 KiB it is smaller than one block per span."""
    rng = random.Random(leaf + mode)
    data = _mixed_local(rng, 3 * (1 << 20) + 70001)                          # ragged tail: partial span, partial piece
    plain = b2d.deflate_chunks(data, b2d.make_opts(mode=mode))
    opts = b2d.make_opts(mode=mode, split_min_bytes=leaf)
    comp, sizes, bits = _indexed_roundtrip(b2d, data, opts)
    assert _decode_both(oracle, comp, len(data)) == data
    assert len(comp) < len(plain) * 0.995, (len(comp), len(plain))
    # homogeneous text: nothing to gain, and nothing to lose beyond the matches cut at piece boundaries
    text = _text(rng, (1 << 20) + 5)
    a = b2d.deflate_chunks(text, b2d.make_opts(mode=mode))
    b = b2d.deflate_chunks(text, opts)
    assert _decode_both(oracle, b, len(text)) == text
    assert len(b) <= len(a) * 1.004, (len(b), len(a))


def test_adaptive_split_small_and_framings(b2d, oracle):
    rng = random.Random(5)
    for n in (0, 1, 4095, 4096, 4097, 16384, 65536, 65537, 100000, 200000):
        data = _mixed_local(rng, n)
        for framing in (0, 1):
            opts = b2d.make_opts(split_min_bytes=4096, framing=framing, chunk_bytes=1 << 17)
            comp = b2d.deflate_chunks(data, opts)
            assert _decode_both(oracle, comp, n) == data, (n, framing)
    # forced modes, greedy reference searches, small chunks
    data = _mixed_local(rng, 300000)
    for mode in (1, 2, 3):
        for search in (0, 2, 3):
            opts = b2d.make_opts(mode=mode, search=search, split_min_bytes=8192, chunk_bytes=1 << 17)
            assert _decode_both(oracle, b2d.deflate_chunks(data, opts), len(data)) == data, (mode, search)
    # invalid piece sizes are refused
    for bad in (1000, 4097, 12288, 2048):
        with pytest.raises(Exception):
            b2d.deflate_chunks(data, b2d.make_opts(split_min_bytes=bad))
    with pytest.raises(Exception):
        b2d.deflate_chunks(data, b2d.make_opts(split_min_bytes=4096, block_bytes=1 << 17))   # 32 pieces per span


@pytest.mark.parametrize("leaf", [4096, 8192, 16384])
def test_adaptive_split_size_against_the_restated_binarysplit(b2d, oracle, leaf):
    """Size parity of the GPU's adaptive splitting with the reference's own optimiser at the same minimum block length:
    BinarySplit(FULL_DYNAMIC, minimumBlockLength = leaf) as restated in the oracle (comp/BinarySplit.java:30-98), run per
    1 MiB chunk like the GPU stream (history reset per chunk, + the 5-byte chunk marker).  The bar is the north-star's:
    at most 1 % larger.  (Bytes cannot be compared: different parse, bottom-up instead of top-down decisions.)"""
    n_chunks, chunk = 4, 1 << 20
    data = b2d.corpus("mixed", 0xDEF1A7E + 77, n_chunks * chunk).tobytes()
    comp, crc, idx = b2d.deflate_chunks(data, b2d.make_opts(split_min_bytes=leaf, chunk_bytes=chunk), crc=0)
    assert _decode_both(oracle, bytes(comp), len(data)) == data
    ref_total, ref_blocks = 0, 0
    for c in range(n_chunks):
        piece = data[c * chunk:(c + 1) * chunk]
        stream, nb = oracle.deflate_split(piece, (oracle.FULL_DYNAMIC,), min_block_len=leaf)
        st, out, _ = oracle.inflate(stream, out_cap=chunk + 8)
        assert st == 0 and out == piece
        ref_total += len(stream) + 5
        ref_blocks += nb
    assert ref_blocks > n_chunks * 16, "the restated BinarySplit did not split anything: the comparison would be empty"
    assert len(comp) <= 1.01 * ref_total, (leaf, len(comp), ref_total, len(comp) / ref_total)
    # ... and the unsplit GPU stream against the unsplit reference strategy, for scale
    plain = b2d.deflate_chunks(data, b2d.make_opts(chunk_bytes=chunk))
    assert len(comp) < len(plain)
