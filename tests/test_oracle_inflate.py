"""CPU: the oracle decoder against the reference's own golden vectors (InflaterInputStreamTest.java:24-510),
its randomized constructions re-made with fixed seeds, and system zlib."""
import random
import zlib

import pytest

from util import BitWriter, bits_to_bytes, fixed_lit_code, golden_vectors, zlib_inflate_raw, zlib_raw

VECTORS = golden_vectors()


def test_vector_inventory():
    assert len(VECTORS) == 39
    assert sum(v["expect"] == "ok" for v in VECTORS) == 11


@pytest.mark.parametrize("v", VECTORS, ids=[v["name"] for v in VECTORS])
def test_golden_vector(oracle, v):
    rng = random.Random(v["line"])
    for pad in ("0", "1", "r", "r"):
        data = bits_to_bytes(v["bits"], pad, rng)
        for slow in (False, True):
            st, out, consumed = oracle.inflate(data, slow=slow)
            if v["expect"] == "ok":
                assert st == 0
                assert out.hex() == v["output_hex"]
                assert consumed == len(data)          # end-exactly, InflaterInputStreamTest.java:557-558
            else:
                assert oracle.status_name(st) == v["reason"]


def test_positive_vectors_agree_with_zlib(oracle):
    for v in VECTORS:
        if v["expect"] != "ok":
            continue
        data = bits_to_bytes(v["bits"], "0")
        out, used = zlib_inflate_raw(data)
        assert out.hex() == v["output_hex"] and used == len(data)


def _log_uniform(rng, limit):
    return rng.randrange(1 << rng.randrange(1, limit.bit_length() + 1)) % (limit + 1)


def test_random_stored_blocks(oracle):
    """InflaterInputStreamTest.java:131-163 with a fixed seed."""
    rng = random.Random(1311)
    for _ in range(60):
        bw, expect = BitWriter(), bytearray()
        nblocks = rng.randrange(1, 6)
        for k in range(nblocks):
            bw.put(1 if k == nblocks - 1 else 0, 1)
            bw.put(0, 2)
            bw.put(rng.randrange(32), 5)              # random padding bits
            n = _log_uniform(rng, 65535)
            payload = rng.randbytes(n)
            bw.put(n, 16)
            bw.put(n ^ 0xFFFF, 16)
            bw.put_bytes(payload)
            expect += payload
        data = bw.tobytes()
        st, out, consumed = oracle.inflate(data, out_cap=len(expect) + 8)
        assert st == 0 and out == bytes(expect) and consumed == len(data)


def test_random_mixed_start_positions(oracle):
    """Stored blocks after 19-bit one-literal fixed blocks hit every start bit position
    (InflaterInputStreamTest.java:166-208)."""
    rng = random.Random(1662)
    for _ in range(60):
        bw, expect = BitWriter(), bytearray()
        for _k in range(rng.randrange(1, 16)):
            if rng.random() < 0.5:
                bw.put(0, 1)
                bw.put(1, 2)
                sym = rng.randrange(144, 256)         # 9-bit literal: 3 + 9 + 7 = 19 bits
                bw.put_code(*fixed_lit_code(sym))
                bw.put_code(*fixed_lit_code(256))
                expect.append(sym)
            else:
                bw.put(0, 1)
                bw.put(0, 2)
                bw.align()
                n = rng.randrange(0, 300)
                payload = rng.randbytes(n)
                bw.put(n, 16)
                bw.put(n ^ 0xFFFF, 16)
                bw.put_bytes(payload)
                expect += payload
        bw.put(1, 1)
        bw.put(1, 2)
        bw.put_code(*fixed_lit_code(256))
        data = bw.tobytes()
        st, out, consumed = oracle.inflate(data, out_cap=len(expect) + 8)
        assert st == 0 and out == bytes(expect) and consumed == len(data)
        assert zlib_inflate_raw(data)[0] == bytes(expect)


def test_random_fixed_literal_blocks(oracle):
    """InflaterInputStreamTest.java:306-338."""
    rng = random.Random(3063)
    for _ in range(40):
        bw, expect = BitWriter(), bytearray()
        nblocks = rng.randrange(1, 4)
        for k in range(nblocks):
            bw.put(1 if k == nblocks - 1 else 0, 1)
            bw.put(1, 2)
            for _j in range(_log_uniform(rng, 32767)):
                b = rng.randrange(256)
                bw.put_code(*fixed_lit_code(b))
                expect.append(b)
            bw.put_code(*fixed_lit_code(256))
        data = bw.tobytes()
        st, out, consumed = oracle.inflate(data, out_cap=len(expect) + 8)
        assert st == 0 and out == bytes(expect) and consumed == len(data)


def test_reasons_not_covered_by_reference_tests(oracle):
    # COPY_FROM_BEFORE_DICTIONARY_START (Open.java:592-593): fixed block, length 3 distance 1 with no output yet
    bw = BitWriter()
    bw.put(1, 1); bw.put(1, 2)
    bw.put_code(*fixed_lit_code(257)); bw.put_code(0, 5)
    st, out, _ = oracle.inflate(bw.tobytes())
    assert oracle.status_name(st) == "COPY_FROM_BEFORE_DICTIONARY_START" and out == b""
    # one literal, then distance 2
    bw = BitWriter()
    bw.put(1, 1); bw.put(1, 2)
    bw.put_code(*fixed_lit_code(65)); bw.put_code(*fixed_lit_code(257)); bw.put_code(1, 5)
    st, out, _ = oracle.inflate(bw.tobytes())
    assert oracle.status_name(st) == "COPY_FROM_BEFORE_DICTIONARY_START" and out == b"A"


def test_zlib_streams_decode_identically(oracle):
    rng = random.Random(7)
    words = [bytes(rng.choices(b"etaoinshrdlu", k=rng.randrange(1, 9))) for _ in range(500)]
    for n in (0, 1, 100, 70000, 300000):
        text = b" ".join(rng.choice(words) for _ in range(n // 5 + 1))[:n]
        for level in (1, 6, 9):
            for strat in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_RLE, zlib.Z_HUFFMAN_ONLY, zlib.Z_FIXED):
                comp = zlib_raw(text, level, strat)
                st, out, consumed = oracle.inflate(comp, out_cap=n + 8)
                assert st == 0 and out == text and consumed == len(comp)
    rnd = rng.randbytes(200000)
    st, out, consumed = oracle.inflate(zlib_raw(rnd, 0), out_cap=len(rnd))
    assert st == 0 and out == rnd


def test_output_overflow_reports_prefix(oracle):
    data = b"abcdefgh" * 100
    comp = zlib_raw(data)
    st, out, _ = oracle.inflate(comp, out_cap=100)
    assert st == oracle.OUTPUT_OVERFLOW and out == data[:100]


def test_mutated_streams_accept_reject_like_zlib(oracle):
    """Differential check of the restatement against an independent decoder: 3000 zlib-made streams (levels 1/6/9,
    default / fixed / Huffman-only / RLE strategies) with 1-3 mutations each (bit flip, byte overwrite, truncation).
    zlib classifies errors more coarsely and later than Open.java, but the three verdicts line up: the oracle
    accepts exactly what zlib accepts (same bytes, same consumed input), reports UNEXPECTED_END_OF_STREAM exactly where
    zlib wants more input, and a format Reason exactly where zlib raises an error."""
    rng = random.Random(99)
    words = [bytes(rng.choices(b"etaoinshrdlu ,.\n", k=rng.randrange(1, 9))) for _ in range(300)]
    verdicts = {}
    for _ in range(3000):
        n = rng.choice([50, 300, 2000, 9000])
        text = bytearray()
        while len(text) < n:
            text += rng.choice(words)
        base = bytearray(zlib_raw(bytes(text[:n]), rng.choice([1, 6, 9]),
                                  rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE])))
        for _ in range(rng.randrange(1, 4)):
            if len(base) < 2:
                break
            m = rng.randrange(3)
            if m == 0:
                base[rng.randrange(len(base))] ^= 1 << rng.randrange(8)
            elif m == 1 and len(base) > 4:
                del base[rng.randrange(1, len(base)):]
            else:
                base[rng.randrange(len(base))] = rng.randrange(256)
        data = bytes(base)
        st, out, consumed = oracle.inflate(data, out_cap=1 << 20)
        d = zlib.decompressobj(-15)
        try:
            zout = d.decompress(data)
            z = "ok" if d.eof else "more"
        except zlib.error:
            z = "error"
        mine = "ok" if st == 0 else "more" if oracle.status_name(st) == "UNEXPECTED_END_OF_STREAM" else "error"
        assert mine == z, (oracle.status_name(st), z, data.hex()[:80])
        if st == 0:
            assert out == zout and consumed == len(data) - len(d.unused_data)
        verdicts[mine] = verdicts.get(mine, 0) + 1
    assert min(verdicts.get(k, 0) for k in ("ok", "more", "error")) > 300, verdicts


def test_mutated_gzip_and_zlib_members_accept_reject_like_python(oracle):
    """The container restatements (GzipInputStream.java:66-90, ZlibInputStream.java:64-83: body, then CRC-32 + ISIZE /
    Adler-32) against Python's gzip and zlib modules: 3000 members, most with 1-2 mutations behind the header (bit flip,
    byte overwrite, truncation).  Accepted by one <=> accepted by the other, with the same bytes."""
    import gzip
    rng = random.Random(5)
    words = [bytes(rng.choices(b"etaoinshrdlu ,.\n", k=rng.randrange(1, 9))) for _ in range(300)]
    seen = set()
    for _ in range(3000):
        text = bytearray()
        n = rng.choice([0, 1, 50, 300, 5000])
        while len(text) < n:
            text += rng.choice(words)
        data = bytes(text[:n])
        is_gzip = rng.random() < 0.5
        base = bytearray(gzip.compress(data, rng.choice([1, 6, 9]), mtime=0) if is_gzip else zlib.compress(data, rng.choice([1, 6, 9])))
        hdr = 10 if is_gzip else 2
        if rng.random() < 0.85:
            for _ in range(rng.randrange(1, 3)):
                m = rng.randrange(3)
                if m == 0:
                    base[rng.randrange(hdr, len(base))] ^= 1 << rng.randrange(8)
                elif m == 1 and len(base) > hdr + 2:
                    del base[rng.randrange(hdr + 1, len(base)):]
                else:
                    base[rng.randrange(hdr, len(base))] = rng.randrange(256)
        member = bytes(base)
        st, out, _ = (oracle.gunzip if is_gzip else oracle.unzlib)(member, out_cap=1 << 16)
        try:
            ref = gzip.decompress(member) if is_gzip else zlib.decompress(member)
        except Exception:
            ref = None
        assert (st == 0) == (ref is not None), (is_gzip, oracle.status_name(st), member.hex()[:80])
        if st == 0:
            assert out == ref
        seen.add((is_gzip, oracle.status_name(st)))
    for is_gzip in (True, False):
        assert {(is_gzip, "OK"), (is_gzip, "DECOMPRESSED_CHECKSUM_MISMATCH"), (is_gzip, "UNEXPECTED_END_OF_STREAM")} <= seen
    assert (True, "DECOMPRESSED_SIZE_MISMATCH") in seen
