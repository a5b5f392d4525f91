"""CPU: the oracle encoder.  The reference has no golden compressed bytes (DeflaterOutputStreamTest.java:24-115 only
round-trips, default strategy), so these mirror its five round-trip tests with fixed seeds, extend them to every
preset, and pin the survey's independent cross-check bytes."""
import random
import zlib

import pytest

from util import zlib_inflate_raw

STRATS = list(range(7))


def _roundtrip(oracle, data, strategies, **kw):
    comp = oracle.deflate(data, strategies, **kw)
    st, out, consumed = oracle.inflate(comp, out_cap=len(data) + 8)
    assert st == 0 and out == bytes(data) and consumed == len(comp)
    assert zlib_inflate_raw(comp)[0] == bytes(data)
    return comp


def test_empty(oracle):                                # DeflaterOutputStreamTest.java:24-29
    for s in STRATS:
        _roundtrip(oracle, b"", (s,))


def test_short_random(oracle):                         # :32-44
    rng = random.Random(32)
    for _ in range(300):
        _roundtrip(oracle, rng.randbytes(rng.randrange(100)), (oracle.RLE_DYNAMIC,))


def test_byte_runs(oracle):                            # :67-86
    rng = random.Random(67)
    data = b"".join(bytes([rng.randrange(256)]) * rng.randrange(1, 1001) for _ in range(1000))
    for s in STRATS:
        _roundtrip(oracle, data, (s,))


def test_long_random_multiblock(oracle):               # :89-115 (lengths up to 1 MB cross the 64 KiB block size)
    rng = random.Random(89)
    for _ in range(6):
        n = rng.randrange(1000000)
        data = bytes(rng.choices(range(256), weights=[1 + 40 * (i < 8) for i in range(256)], k=n))
        _roundtrip(oracle, data, (oracle.RLE_DYNAMIC,))
        _roundtrip(oracle, data[:200000], (oracle.FULL_DYNAMIC,))


def test_history_crosses_blocks(oracle):
    rng = random.Random(5)
    seg = rng.randbytes(20000)
    data = seg * 9                                      # repeats at distance 20000 across 64 KiB block boundaries
    full = _roundtrip(oracle, data, (oracle.FULL_DYNAMIC,))
    assert len(full) < len(data) // 4
    nohist = _roundtrip(oracle, data, (oracle.FULL_DYNAMIC,), lookahead=16384, history=0)
    assert len(nohist) > len(data) * 0.9


def test_brute_force_equals_hash_chains(oracle):
    rng = random.Random(11)
    words = [rng.randbytes(rng.randrange(1, 8)) for _ in range(64)]
    data = b"".join(rng.choice(words) for _ in range(3000))
    for s in (oracle.FULL_STATIC, oracle.FULL_DYNAMIC, oracle.RLE_DYNAMIC):
        assert oracle.deflate(data, (s,), brute_force=True) == oracle.deflate(data, (s,), brute_force=False)


def test_multistrategy_never_worse(oracle):
    rng = random.Random(13)
    for data in (rng.randbytes(70000), bytes(70000), b"abcabcabd" * 5000, b"x"):
        multi = _roundtrip(oracle, data, (oracle.UNCOMPRESSED, oracle.FULL_STATIC, oracle.FULL_DYNAMIC))
        for s in (oracle.UNCOMPRESSED, oracle.FULL_STATIC, oracle.FULL_DYNAMIC):
            assert len(multi) <= len(oracle.deflate(data, (s,)))


APPENDIX_F = [
    (b"", "LITERAL_STATIC", "0300"),
    (b"", "FULL_DYNAMIC", "05c0810800000000a0fda92f"),
    (b"A", "RLE_STATIC", "730400"),
    (b"A", "RLE_DYNAMIC", "05c081080000000020b6fda54e"),
    (b"abc" * 1000, "FULL_STATIC", "4b4c4a1e45a368148da251348a46d1281a45a368140d720400"),
    (b"abc" * 1000, "FULL_DYNAMIC", "edc3411100000c02a0ac9bfd3b58c30777701f0000605c01"),
    (bytes(1000), "RLE_STATIC", "631805a360140c7b0000"),
    (bytes(1000), "FULL_DYNAMIC", "edc1010d000000c220fba7b6c7070cc83b"),
    (b"a" * 10 + b"b" * 10, "LITERAL_STATIC", "4b4c4c4c4c4c4c4c4c4c4c4a4a4a4a4a4a4a4a4a0200"),
    (b"a" * 10 + b"b" * 10, "FULL_STATIC", "4b848324380000"),
    (b"a" * 10 + b"b" * 10, "LITERAL_DYNAMIC", "05c0810c0000008030d6ee0fd1aaaa0a8001"),
    (b"a" * 10 + b"b" * 10, "RLE_DYNAMIC", "3dc13101000000c2a0acda3fc43e609c00"),
]
APPENDIX_F_CRC = [
    (b"abc" * 1000, "RLE_DYNAMIC", 763, 0xF664BED9),
    (bytes(1000), "LITERAL_DYNAMIC", 137, 0xC375BE96),
    (bytes(range(256)), "FULL_STATIC", 272, 0x8557FBB7),
    (bytes(range(256)), "FULL_DYNAMIC", 280, 0xFF540EEC),
]


@pytest.mark.parametrize("data,strat,hexout", APPENDIX_F)
def test_cross_check_bytes(oracle, data, strat, hexout):
    assert oracle.deflate(data, (getattr(oracle, strat),)).hex() == hexout


@pytest.mark.parametrize("data,strat,n,crc", APPENDIX_F_CRC)
def test_cross_check_crc(oracle, data, strat, n, crc):
    comp = oracle.deflate(data, (getattr(oracle, strat),))
    assert len(comp) == n and zlib.crc32(comp) == crc


def test_package_merge_is_optimal_and_complete(oracle):
    rng = random.Random(3)
    for _ in range(200):
        n = rng.randrange(2, 286)
        hist = [rng.randrange(0, 1 << rng.randrange(1, 16)) if rng.random() < 0.7 else 0 for _ in range(n)]
        if sum(h > 0 for h in hist) < 2:
            continue
        for limit in (15, 7) if sum(h > 0 for h in hist) <= 128 else (15,):
            lens = oracle.package_merge(hist, limit)
            assert all((l > 0) == (h > 0) for l, h in zip(lens, hist))
            assert max(lens) <= limit
            assert sum(2 ** (limit - l) for l in lens if l) == 2 ** limit     # complete code


def test_crc32_and_gzip_container(oracle, tmp_path):
    import gzip
    import subprocess
    assert oracle.crc32(b"123456789") == 0xCBF43926
    rng = random.Random(17)
    data = rng.randbytes(1000) + b"hello " * 30000
    assert oracle.crc32(data) == zlib.crc32(data)
    member = oracle.gzip_member(data, file_name="in.txt", mtime=1700000000)
    assert gzip.decompress(member) == data                                      # Python's gzip (zlib)
    st, out, consumed = oracle.gunzip(member, out_cap=len(data) + 8)
    assert st == 0 and out == data and consumed == len(member)
    p = tmp_path / "x.gz"
    p.write_bytes(member)
    assert subprocess.run(["gzip", "-t", str(p)]).returncode == 0               # system gzip 1.12
    bad = bytearray(member); bad[-5] ^= 1
    assert oracle.status_name(oracle.gunzip(bytes(bad), out_cap=len(data) + 8)[0]) == "DECOMPRESSED_CHECKSUM_MISMATCH"
    bad = bytearray(member); bad[-1] ^= 1
    assert oracle.status_name(oracle.gunzip(bytes(bad), out_cap=len(data) + 8)[0]) == "DECOMPRESSED_SIZE_MISMATCH"
    bad = bytearray(member); bad[0] ^= 1
    assert oracle.status_name(oracle.gunzip(bytes(bad), out_cap=len(data) + 8)[0]) == "GZIP_INVALID_MAGIC_NUMBER"


def test_zlib_container_oracle(oracle):
    """ZlibInputStream / Adler-32 restatement against Python's zlib (the reference has no tests for the containers)."""
    rng = random.Random(1950)
    for n in (0, 1, 5552, 5553, 70000, 300000):
        d = rng.randbytes(n // 2) + bytes(n - n // 2)
        assert oracle.adler32(d) == zlib.adler32(d)
        z = zlib.compress(d, 6)
        st, out, consumed = oracle.unzlib(z, out_cap=n + 8)
        assert st == 0 and out == d and consumed == len(z)
    z = zlib.compress(b"hello hello hello hello", 9)
    assert oracle.status_name(oracle.unzlib(z[:1])[0]) == "UNEXPECTED_END_OF_STREAM"
    assert oracle.status_name(oracle.unzlib(bytes([z[0], z[1] ^ 1]) + z[2:])[0]) == "HEADER_CHECKSUM_MISMATCH"
    bad = bytes([0x77, 0x77 ^ 0]) + z[2:]
    fix = (31 - ((0x77 << 8) % 31)) % 31
    assert oracle.status_name(oracle.unzlib(bytes([0x77, fix]) + z[2:])[0]) == "UNSUPPORTED_COMPRESSION_METHOD"
    assert oracle.status_name(oracle.unzlib(z[:-1] + bytes([z[-1] ^ 1]))[0]) == "DECOMPRESSED_CHECKSUM_MISMATCH"
    assert oracle.status_name(oracle.unzlib(z[:-2])[0]) == "UNEXPECTED_END_OF_STREAM"


# ---- BinarySplit (comp/BinarySplit.java): the reference ships it, never uses it by default and has no test for it ----
def _mixed(rng):
    words = [bytes(rng.choices(b"etaoinshrdlucmfw", k=rng.randrange(1, 10))) for _ in range(400)]
    def text(n):
        out = bytearray()
        while len(out) < n:
            out += rng.choice(words) + b" "
        return bytes(out[:n])
    return text(40000) + rng.randbytes(30000) + bytes(20000) + text(25000) + bytes(range(256)) * 60


@pytest.mark.parametrize("strategies", [(3,), (5,), (6, 4, 5)], ids=["RLE_DYNAMIC", "FULL_DYNAMIC", "Multi"])
def test_binary_split_roundtrip_and_never_worse(oracle, strategies):
    """BinarySplit.decide keeps a cut only where it is cheaper (:60-69), so the stream is never larger than the
    substrategy's own, it decodes to the input through both decoders, and on data whose statistics change inside a
    block it is strictly smaller; with a minimum length of half the block no cut is allowed (:42)."""
    rng = random.Random(77)
    data = _mixed(rng)
    plain = oracle.deflate(data, strategies)
    prev_blocks = None
    for min_len in (1, 1000, 4096, 16384):
        comp, blocks = oracle.deflate_split(data, strategies, min_len)
        st, out, cons = oracle.inflate(comp, out_cap=len(data) + 8)
        assert st == 0 and out == data and cons == len(comp)
        assert zlib_inflate_raw(comp)[0] == data
        assert len(comp) < len(plain)
        assert prev_blocks is None or blocks <= prev_blocks          # a larger minimum allows fewer cuts
        prev_blocks = blocks
    comp, blocks = oracle.deflate_split(data, strategies, 32768)
    assert comp == plain                                             # halves of a 64 KiB block are not > 32768: no cut
    for data in (b"", b"a", bytes(1000), rng.randbytes(70000)):
        comp, _ = oracle.deflate_split(data, strategies, 8)
        st, out, _ = oracle.inflate(comp, out_cap=len(data) + 8)
        assert st == 0 and out == data
        assert len(comp) <= len(oracle.deflate(data, strategies))


def test_binary_split_brute_force_equals_hash_chains(oracle):
    rng = random.Random(78)
    data = _mixed(rng)[:30000]
    a = oracle.deflate_split(data, (5,), 512, brute_force=True)
    b = oracle.deflate_split(data, (5,), 512, brute_force=False)
    assert a == b
