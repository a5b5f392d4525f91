"""GPU parity: b2d_inflate_batch (through the C ABI, host pointers) against the oracle restating
decomp/Open.java, the reference's golden vectors, and zlib."""
import os
import random
import zlib

import numpy as np
import pytest

from util import BitWriter, bits_to_bytes, fixed_lit_code, golden_vectors, zlib_raw

pytestmark = pytest.mark.gpu
VECTORS = golden_vectors()


def _check_against_oracle(b2d, oracle, members, caps, flags=0):
    outs, out_len, consumed, crc, status = b2d.inflate_batch(members, caps, flags)
    if isinstance(caps, int):
        caps = [caps] * len(members)
    for i, m in enumerate(members):
        st, out, cons = oracle.inflate(bytes(m), out_cap=caps[i])
        assert int(status[i]) == st, (i, b2d.status_name(int(status[i])), oracle.status_name(st))
        assert outs[i] == out, (i, len(outs[i]), len(out))
        if st == 0:
            assert int(consumed[i]) == cons, i
            if flags & b2d.INFLATE_CRC32:
                assert int(crc[i]) == zlib.crc32(out)
    return outs, status


def test_golden_vectors_all_paddings(b2d, oracle):
    """All 39 vectors x (0-pad, 1-pad, random pad) in ONE batch: a bad member must not disturb its neighbours."""
    rng = random.Random(39)
    members, expect = [], []
    for v in VECTORS:
        for pad in ("0", "1", "r", "r"):
            members.append(bits_to_bytes(v["bits"], pad, rng))
            expect.append(v)
    outs, out_len, consumed, crc, status = b2d.inflate_batch(members, 1024)
    for i, v in enumerate(expect):
        if v["expect"] == "ok":
            assert status[i] == 0, (v["name"], b2d.status_name(int(status[i])))
            assert outs[i].hex() == v["output_hex"], v["name"]
            assert consumed[i] == len(members[i]), v["name"]       # end-exactly (InflaterInputStreamTest.java:557-558)
        else:
            assert b2d.status_name(int(status[i])) == v["reason"], v["name"]
    _check_against_oracle(b2d, oracle, members, 1024)
    _check_against_oracle(b2d, oracle, members, 64)      # and with a slot too small for some vectors


def test_empty_batch_and_empty_member(b2d):
    outs, out_len, consumed, crc, status = b2d.inflate_batch([], 0)
    assert outs == []
    outs, out_len, consumed, crc, status = b2d.inflate_batch([b""], 16)
    assert b2d.status_name(int(status[0])) == "UNEXPECTED_END_OF_STREAM"


def _text(rng, n):
    words = [bytes(rng.choices(b"etaoinshrdlucmfw", k=rng.randrange(1, 10))) for _ in range(700)]
    out = bytearray()
    while len(out) < n:
        out += rng.choice(words) + b" "
    return bytes(out[:n])


def test_zlib_members(b2d, oracle):
    rng = random.Random(101)
    members, datas = [], []
    for n in (0, 1, 2, 257, 258, 259, 4096, 70000, 262144, 300001):
        data = _text(rng, n)
        for level in (1, 6, 9):
            for strat in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_RLE, zlib.Z_HUFFMAN_ONLY, zlib.Z_FIXED):
                members.append(zlib_raw(data, level, strat))
                datas.append(data)
    rnd = rng.randbytes(150000)
    members.append(zlib_raw(rnd, 0)); datas.append(rnd)           # stored blocks
    members.append(zlib_raw(rnd, 6)); datas.append(rnd)
    z = bytes(300000)
    members.append(zlib_raw(z, 9)); datas.append(z)               # length 258 / distance 1
    caps = [len(d) + 7 for d in datas]
    outs, status = _check_against_oracle(b2d, oracle, members, caps, flags=b2d.INFLATE_CRC32)
    for o, d in zip(outs, datas):
        assert o == d


def test_oracle_encoded_members_all_presets(b2d, oracle):
    rng = random.Random(202)
    datas = [b"", b"A", b"abc" * 1000, bytes(1000), bytes(range(256)), _text(rng, 200000), rng.randbytes(70000),
             b"".join(bytes([rng.randrange(256)]) * rng.randrange(1, 600) for _ in range(300))]
    members, expect = [], []
    for d in datas:
        for s in range(7):
            members.append(oracle.deflate(d, (s,)))
            expect.append(d)
        members.append(oracle.deflate(d, (oracle.UNCOMPRESSED, oracle.FULL_STATIC, oracle.FULL_DYNAMIC)))
        expect.append(d)
    outs, status = _check_against_oracle(b2d, oracle, members, [len(d) + 3 for d in expect], flags=b2d.INFLATE_CRC32)
    for o, d in zip(outs, expect):
        assert o == d


def test_random_constructions_of_reference_tests(b2d, oracle):
    """Stored blocks with random padding / every start bit position / fixed literal blocks
    (InflaterInputStreamTest.java:131-163, 166-208, 306-338) with fixed seeds."""
    rng = random.Random(303)
    members, expect = [], []
    for _ in range(200):
        bw, exp = BitWriter(), bytearray()
        for _k in range(rng.randrange(1, 12)):
            kind = rng.randrange(3)
            bw.put(0, 1)
            if kind == 0:
                bw.put(0, 2)
                bw.put(rng.randrange(256), (-len(bw.bits)) % 8)   # random padding bits
                n = rng.randrange(1 << rng.randrange(1, 17))
                payload = rng.randbytes(n)
                bw.put(n, 16); bw.put(n ^ 0xFFFF, 16); bw.put_bytes(payload)
                exp += payload
            elif kind == 1:
                bw.put(1, 2)
                sym = rng.randrange(144, 256)
                bw.put_code(*fixed_lit_code(sym)); bw.put_code(*fixed_lit_code(256))
                exp.append(sym)
            else:
                bw.put(1, 2)
                for _j in range(rng.randrange(1 << rng.randrange(1, 13))):
                    b = rng.randrange(256)
                    bw.put_code(*fixed_lit_code(b)); exp.append(b)
                bw.put_code(*fixed_lit_code(256))
        bw.put(1, 1); bw.put(1, 2); bw.put_code(*fixed_lit_code(256))
        members.append(bw.tobytes()); expect.append(bytes(exp))
    outs, status = _check_against_oracle(b2d, oracle, members, [len(e) + 1 for e in expect])
    assert all(s == 0 for s in status)
    for o, e in zip(outs, expect):
        assert o == e


def test_truncations_and_corruptions_match_oracle(b2d, oracle):
    """Every prefix of a few streams, and random single-byte corruptions: status, delivered bytes and (for OK)
    consumed input must equal the oracle's."""
    rng = random.Random(404)
    base = [zlib_raw(_text(rng, 3000), 6), zlib_raw(_text(rng, 3000), 6, zlib.Z_FIXED), zlib_raw(rng.randbytes(500), 0),
            oracle.deflate(_text(rng, 2000), (oracle.FULL_DYNAMIC,))]
    members = []
    for s in base:
        for cut in range(0, len(s), max(1, len(s) // 97)):
            members.append(s[:cut])
        for _ in range(150):
            b = bytearray(s)
            b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
            members.append(bytes(b))
    _check_against_oracle(b2d, oracle, members, 8192)


def test_members_at_every_phase_of_a_128_byte_line(b2d, oracle):
    """The symbol loop keeps a 128-byte line of the member's input across the lanes of the warp, numbered from the line
    of the member's first byte (decode_block_fast, bitin_init): members starting at every byte phase of a line, long
    enough for several lines and short enough to end inside the first, and cut short at every length around a line's
    end, must decode like the oracle says."""
    rng = random.Random(128)
    text = _text(rng, 9000)
    body = zlib_raw(text[:6000], 6)                       # ~2.5 KB: some twenty lines
    short = zlib_raw(text[:40], 6)                        # ends inside its first line
    stored = zlib_raw(rng.randbytes(700), 0)
    members, offset = [], 0
    for phase in range(128):
        pad = (phase - offset) % 128
        if pad:
            members.append(rng.randbytes(pad))            # (garbage: whatever it decodes to, the oracle says the same)
            offset += pad
        m = (body, short, stored)[phase % 3] if phase % 5 else body[:len(body) - 1 - phase]
        members.append(m)
        offset += len(m)
    _check_against_oracle(b2d, oracle, members, 8192, flags=b2d.INFLATE_CRC32)
    # the end of input against the line: every length of the last 300 bytes, at two phases
    for lead in (0, 77):
        members = [rng.randbytes(lead)] if lead else []
        for cut in range(300):
            members.append(body[:len(body) - cut])
        _check_against_oracle(b2d, oracle, members, 8192)


def test_uncovered_reasons(b2d, oracle):
    bw = BitWriter()
    bw.put(1, 1); bw.put(1, 2); bw.put_code(*fixed_lit_code(257)); bw.put_code(0, 5)
    m1 = bw.tobytes()
    bw = BitWriter()
    bw.put(1, 1); bw.put(1, 2); bw.put_code(*fixed_lit_code(65)); bw.put_code(*fixed_lit_code(257)); bw.put_code(1, 5)
    m2 = bw.tobytes()
    outs, status = _check_against_oracle(b2d, oracle, [m1, m2], 64)
    assert [b2d.status_name(int(s)) for s in status] == ["COPY_FROM_BEFORE_DICTIONARY_START"] * 2
    assert outs == [b"", b"A"]


def test_output_overflow(b2d, oracle):
    data = b"abcdefgh" * 1000 + bytes(5000)
    m = zlib_raw(data)
    outs, out_len, consumed, crc, status = b2d.inflate_batch([m, m, m], [100, len(data), len(data) - 1])
    assert status[0] == b2d.ERR_OUTPUT_OVERFLOW and outs[0] == data[:100]
    assert status[1] == 0 and outs[1] == data
    assert status[2] == b2d.ERR_OUTPUT_OVERFLOW and outs[2] == data[:-1]


def test_batch_shape_of_config2_small(b2d, oracle):
    """The BASELINE config-2 shape at reduced count: independent 256 KiB members, fixed output stride."""
    n, size = 96, 256 * 1024
    datas = [b2d.corpus("text", 0xDEF1A7E + i, size).tobytes() for i in range(n)]
    members = [zlib_raw(d, 6) for d in datas]
    outs, out_len, consumed, crc, status = b2d.inflate_batch(members, size, flags=b2d.INFLATE_CRC32)
    assert all(s == 0 for s in status)
    for i in range(n):
        assert outs[i] == datas[i] and consumed[i] == len(members[i]) and crc[i] == zlib.crc32(datas[i])
    st, out, cons = oracle.inflate(members[0], out_cap=size)
    assert st == 0 and out == outs[0]


def test_gunzip_batch_matches_reference_gzip_reader(b2d, oracle):
    """b2d_gunzip_batch vs the oracle's restatement of GzipInputStream (GzipInputStream.java:38-90,
    GzipMetadata.java:73-146): good members from three writers, and every container failure reason."""
    import gzip as pygzip
    rng = random.Random(1952)
    datas = [b"", b"x", _text(rng, 70000), rng.randbytes(30000), bytes(100000), _text(rng, 300000)]
    members = []
    for d in datas:
        members.append(pygzip.compress(d, 6, mtime=0))                                           # Python/zlib writer
        members.append(oracle.gzip_member(d, file_name="a.txt", mtime=1700000000))               # gzip.java as restated (FNAME + FHCRC)
        members.append(oracle.gzip_member(d, file_name=None, mtime=0, extra=b"B2\x04\x00abcd"))   # FEXTRA
    good = pygzip.compress(datas[2], 6, mtime=0)
    hc = oracle.gzip_member(datas[2], file_name="h", mtime=5)
    bad = [
        b"", good[:1], good[:5], good[:9],                       # truncated header
        b"\x1f\x8c" + good[2:],                                  # magic
        good[:2] + b"\x07" + good[3:],                           # method
        good[:3] + b"\x20" + good[4:],                           # reserved flag
        good[:9] + b"\x0e" + good[10:],                          # operating system 14
        good[:9] + b"\xff" + good[10:],                          # operating system "unknown" is legal
        hc[:12] + bytes([hc[12] ^ 1]) + hc[13:],                 # header CRC-16 mismatch (file name byte flipped)
        good[:-8] + bytes([good[-8] ^ 1]) + good[-7:],           # CRC-32
        good[:-1] + bytes([good[-1] ^ 1]),                       # ISIZE
        good[:-3],                                               # truncated trailer
        good[:len(good) // 2],                                   # truncated body
        good + b"trailing garbage is ignored",                   # GzipInputStream.java:66-74
    ]
    members += bad
    caps = [400000] * len(members)
    outs, out_len, consumed, status = b2d.gunzip_batch(members, caps)
    for i, m in enumerate(members):
        st, out, cons = oracle.gunzip(m, out_cap=caps[i])
        assert int(status[i]) == st, (i, b2d.status_name(int(status[i])), oracle.status_name(st))
        assert outs[i] == out, (i, len(outs[i]), len(out))        # also on failure: the bytes delivered before it
        if st == 0:
            assert int(consumed[i]) == cons, i
    outs2, _, _, status2 = b2d.gunzip_batch(members[:18])     # capacities from ISIZE
    assert not status2.any() and outs2 == outs[:18]
    # the same batch into page-locked memory (bytes delivered by the decoding warps): identical in every respect
    outs3, out_len3, consumed3, status3 = b2d.gunzip_batch(members, caps, pinned_out=True)
    assert outs3 == outs and np.array_equal(out_len3, out_len) and np.array_equal(consumed3, consumed)
    assert np.array_equal(status3, status)


def test_gunzip_batch_member_never_reads_its_neighbour(b2d, oracle):
    """A truncated or corrupt member must end where ITS bytes end (GzipInputStream reading that member alone), not run
    on into the next member's header and body: status, out_len and the delivered bytes equal the oracle's for the
    member on its own, whatever follows it in the batch."""
    import gzip as pygzip
    rng = random.Random(77)
    a, b = _text(rng, 120000), _text(rng, 90000)
    ga, gb = pygzip.compress(a, 6, mtime=0), pygzip.compress(b, 6, mtime=0)
    raw_a = zlib_raw(a, 6)
    cases = [ga[:len(ga) // 2], ga[:-9], ga[:-8], ga[:-4], ga[:10 + len(raw_a) - 1], ga[:10 + 3], ga[:10]]
    for pinned in (False, True):
        members = []
        for c in cases:
            members += [c, gb]                                   # every damaged member is followed by a valid one
        outs, out_len, consumed, status = b2d.gunzip_batch(members, [200000] * len(members), pinned_out=pinned)
        for i, m in enumerate(members):
            st, out, cons = oracle.gunzip(m, out_cap=200000)
            assert int(status[i]) == st, (i, b2d.status_name(int(status[i])), oracle.status_name(st))
            assert int(out_len[i]) == len(out) and outs[i] == out, (i, int(out_len[i]), len(out))
            assert int(consumed[i]) <= len(m), (i, int(consumed[i]), len(m))
            if st == 0:
                assert int(consumed[i]) == cons
    # raw members through b2d_inflate_batch: a truncated member's range ends at the next member's start
    members = [raw_a[:len(raw_a) // 3], zlib_raw(b, 6), raw_a[:-1], zlib_raw(b, 1)]
    outs, out_len, consumed, crc, status = b2d.inflate_batch(members, 200000)
    for i, m in enumerate(members):
        st, out, cons = oracle.inflate(m, out_cap=200000)
        assert int(status[i]) == st and outs[i] == out and int(consumed[i]) <= len(m), i


def test_random_garbage_members(b2d, oracle):
    """Members of pure noise and of noise spliced into valid streams: whatever the reference decoder would report
    (status, bytes delivered before the failure), the batch reports the same, and nothing hangs or crosses slots."""
    rng = random.Random(666)
    valid = zlib_raw(_text(rng, 20000), 6)
    members = []
    for _ in range(1500):
        kind = rng.randrange(3)
        if kind == 0:
            members.append(rng.randbytes(rng.randrange(0, 300)))
        elif kind == 1:
            cut = rng.randrange(len(valid))
            members.append(valid[:cut] + rng.randbytes(rng.randrange(1, 200)))
        else:                                           # a plausible dynamic-block header followed by noise
            members.append(bytes([rng.choice([0x05, 0x0D, 0x04, 0x0C, 0xED, 0xBD])]) + rng.randbytes(rng.randrange(10, 400)))
    outs, out_len, consumed, crc, status = b2d.inflate_batch(members, 4096)
    n_ok = 0
    for i, m in enumerate(members):
        st, out, cons = oracle.inflate(m, out_cap=4096)
        assert int(status[i]) == st, (i, b2d.status_name(int(status[i])), oracle.status_name(st), m[:8].hex())
        assert outs[i] == out, i
        n_ok += st == 0
    assert n_ok < len(members) // 2


def test_pinned_output_is_delivered_by_the_kernel(b2d, oracle):
    """With a page-locked output buffer b2d_inflate_batch skips the D2H copy: the decoding warps write whole 128-byte
    lines to the mapped host address themselves.  Same bytes, lengths, statuses as the pageable path, for ragged
    slots at odd addresses, stored blocks, empty and failing members."""
    rng = random.Random(4242)
    members, caps = [], []
    for n in (0, 1, 100, 127, 128, 129, 4095, 4096, 4097, 8191, 70000, 262144, 300001):
        data = _text(rng, n)
        for level, strat in ((6, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_FIXED), (0, zlib.Z_DEFAULT_STRATEGY)):
            members.append(zlib_raw(data, level, strat))
            caps.append(n + rng.randrange(0, 40))
    rnd = rng.randbytes(200000)
    members.append(zlib_raw(rnd, 0)); caps.append(200000)                    # stored blocks only
    members.append(zlib_raw(bytes(500000), 9)); caps.append(500001)          # length 258 / distance 1
    members.append(zlib_raw(_text(rng, 50000), 6)); caps.append(30000)       # slot too small: overflow after 30000 bytes
    members.append(zlib_raw(_text(rng, 50000), 6)[:9000]); caps.append(50000)   # truncated stream
    for v in VECTORS[:12]:
        members.append(bits_to_bytes(v["bits"], "0", rng)); caps.append(64)
    order = list(range(len(members)))
    rng.shuffle(order)
    members = [members[i] for i in order]
    caps = [caps[i] for i in order]
    ref = b2d.inflate_batch(members, caps, b2d.INFLATE_CRC32)
    total = sum(caps)
    # page-locked input alone (copied to the device like pageable input, but asynchronously), and with a pinned output
    got = b2d.inflate_batch(members, caps, b2d.INFLATE_CRC32, pinned_in=True)
    assert got[0] == ref[0]
    for k in range(1, 5):
        assert np.array_equal(got[k], ref[k]), k
    for phase in (0, 3, 77):                                                 # the slots start at odd host addresses too
        buf = b2d.PinnedBuffer(total + 256)
        buf.array[:] = 0xEE
        got = b2d.inflate_batch(members, caps, b2d.INFLATE_CRC32, out=buf.array[phase:phase + max(total, 1)],
                                pinned_in=phase != 3)
        assert got[0] == ref[0]
        for k in range(1, 5):
            assert np.array_equal(got[k], ref[k]), k
        # nothing outside the delivered bytes is touched
        assert (buf.array[:phase] == 0xEE).all() and (buf.array[phase + total:] == 0xEE).all()
        off = phase
        for i, c in enumerate(caps):
            n = int(got[1][i])
            assert (buf.array[off + n:off + c] == 0xEE).all(), i
            off += c
    _check_against_oracle(b2d, oracle, members, caps, flags=b2d.INFLATE_CRC32)


def test_pinned_uniform_slots_are_delivered_by_the_copy_engine(b2d, oracle):
    """Slots of one size in a page-locked buffer: the decoding warps only REPORT progress (32 KiB pieces) and the host
    moves finished pieces of all members with strided copy-engine transfers while the decode runs.  Same bytes, lengths,
    consumed counts, checksums and statuses as the pageable path and as the oracle, for members that fill the slot, stop
    short, overflow it, fail, are stored-only, or are empty; and B2D_INFLATE_D2H picks the other paths for the same call."""
    rng = random.Random(31337)
    cap = 100000                                          # not a multiple of the piece size
    members = []
    for i in range(300):
        kind = i % 10
        if kind == 0:
            members.append(zlib_raw(rng.randbytes(cap), 0))                       # stored blocks only, fills the slot
        elif kind == 1:
            members.append(zlib_raw(_text(rng, rng.randrange(0, 3000)), 6))       # far short of the slot
        elif kind == 2:
            members.append(zlib_raw(_text(rng, cap + 5000), 6))                   # overflows the slot
        elif kind == 3:
            members.append(zlib_raw(_text(rng, 60000), 6)[:rng.randrange(1, 9000)])   # truncated
        elif kind == 4:
            members.append(zlib_raw(bytes(cap), 9))                               # length 258 / distance 1
        elif kind == 5:
            members.append(b"")
        else:
            members.append(zlib_raw(_text(rng, cap - rng.randrange(0, 40000)), rng.choice([1, 6, 9])))
    ref = b2d.inflate_batch(members, cap, b2d.INFLATE_CRC32)
    for i in (0, 1, 2, 3, 4, 5, 6, 17, 299):
        st, out, cons = oracle.inflate(members[i], out_cap=cap)
        assert int(ref[4][i]) == st and ref[0][i] == out, i
    for mode in ("", "ce", "mirror", "copy"):
        if mode:
            os.environ["B2D_INFLATE_D2H"] = mode
        try:
            buf = b2d.PinnedBuffer(cap * len(members) + 64)
            buf.array[:] = 0xEE
            got = b2d.inflate_batch(members, cap, b2d.INFLATE_CRC32, out=buf.array[7:7 + cap * len(members)], pinned_in=True)
        finally:
            os.environ.pop("B2D_INFLATE_D2H", None)
        assert got[0] == ref[0], mode
        for k in range(1, 5):
            assert np.array_equal(got[k], ref[k]), (mode, k)
        assert (buf.array[:7] == 0xEE).all() and (buf.array[7 + cap * len(members):] == 0xEE).all(), mode
    # gzip members through the same path (b2d_gunzip_batch, the bench's end-to-end call)
    import gzip as pygzip
    datas = [_text(rng, 65536 - (i % 7) * 1000) for i in range(130)]
    gz = [pygzip.compress(d, 6, mtime=0) for d in datas]
    outs, out_len, consumed, status = b2d.gunzip_batch(gz, 65536, pinned_out=True)
    assert not status.any() and outs == datas and [int(c) for c in consumed] == [len(g) for g in gz]
