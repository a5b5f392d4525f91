"""GPU: the host-side mirror of the reference's stream classes and CLIs (deflate-library-java_b200/host), driven the
way src/gzip.java / src/gunzip.java are: files in, files out, exit code 1 + message on error.  Interop both ways with
system gzip and with the oracle's restatement of GzipInputStream."""
import gzip as pygzip
import os
import subprocess
import zlib

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "deflate-library-java_b200", "bin")
GZIP, GUNZIP = os.path.join(BIN, "gzip"), os.path.join(BIN, "gunzip")

pytestmark = pytest.mark.gpu


def run(*args, env=None):
    e = dict(os.environ)
    if env:
        e.update(env)
    return subprocess.run(list(args), capture_output=True, text=True, env=e)


# (every CLI run is a process that creates a CUDA context, 2-3 s on the GPU boxes: the matrix is kept small)
@pytest.mark.parametrize("kind,n,index", [("text", 0, "1"), ("text", 0, "0"), ("text", 1, "1"), ("text", 70000, "0"), ("text", 70000, "split"),
                                          ("mixed", (5 << 20) + 12345, "1"), ("mixed", (5 << 20) + 12345, "0"),
                                          ("mixed", (5 << 20) + 12345, "split"), ("random", 3 << 20, "1")])
def test_cli_roundtrip_and_interop(b2d, oracle, tmp_path, kind, n, index):
    data = b2d.corpus(kind, 0xDEF1A7E, n).tobytes()
    src, gz, back = tmp_path / "in.bin", tmp_path / "out.gz", tmp_path / "back.bin"
    src.write_bytes(data)
    env = {"B2D_GZIP_INDEX": index}
    if index == "split":                                         # adaptive block splitting, with the index
        env = {"B2D_GZIP_INDEX": "1", "B2D_GZIP_SPLIT": "8192"}
        index = "1"
    r = run(GZIP, str(src), str(gz), env=env)
    assert r.returncode == 0, r.stderr
    assert "Input  speed:" in r.stderr and "Output speed:" in r.stderr              # gzip.java:73-74
    member = gz.read_bytes()
    # the reference's reader (restated) and two independent gzip readers accept the file
    st, out, consumed = oracle.gunzip(member, out_cap=n + 16)
    assert st == 0 and out == data and consumed == len(member)
    assert pygzip.decompress(member) == data
    assert run("gzip", "-t", str(gz)).returncode == 0
    # header as gzip.java writes it: FNAME + FHCRC (+ FEXTRA with the chunk index), OS = Unix
    assert member[:3] == b"\x1f\x8b\x08" and member[9] == 3
    assert member[3] & 0x0A == 0x0A and bool(member[3] & 4) == (index == "1")
    if index == "1" and n > 0:                                   # chunk sizes ("B2") and block bit offsets ("B3") in FEXTRA
        xlen = member[10] | member[11] << 8
        extra = member[12:12 + xlen]
        assert extra[:2] == b"B2" and b"B3" in extra
    r = run(GUNZIP, str(gz), str(back))
    assert r.returncode == 0, r.stderr
    assert back.read_bytes() == data
    assert "File name: in.bin" in r.stderr and "Operating system: Unix" in r.stderr and "File mode: Binary" in r.stderr


def test_gunzip_reads_system_gzip_files(b2d, tmp_path):
    data = b2d.corpus("text", 7, 3 << 20).tobytes()
    src, back = tmp_path / "t.txt", tmp_path / "t.out"
    src.write_bytes(data)
    subprocess.check_call(["gzip", "-k", "-6", str(src)])
    r = run(GUNZIP, str(src) + ".gz", str(back))
    assert r.returncode == 0, r.stderr
    assert back.read_bytes() == data
    assert "File name: t.txt" in r.stderr


def test_gunzip_error_convention(b2d, tmp_path):
    data = b2d.corpus("text", 9, 200000).tobytes()
    good = pygzip.compress(data, 6, mtime=0)
    cases = {
        "DataFormatException: Decompression CRC-32 mismatch": good[:-8] + bytes([good[-8] ^ 1]) + good[-7:],
        "DataFormatException: Decompressed size mismatch": good[:-1] + bytes([good[-1] ^ 1]),
        "DataFormatException: Invalid GZIP magic number": b"\x1f\x8c" + good[2:],
        "DataFormatException: Unexpected end of stream": good[:len(good) // 2],
        "DataFormatException: Reserved flags are set": good[:3] + b"\x80" + good[4:],
    }
    for msg, blob in cases.items():
        p = tmp_path / "bad.gz"
        p.write_bytes(blob)
        r = run(GUNZIP, str(p), str(tmp_path / "bad.out"))
        assert r.returncode == 1 and msg in r.stderr, (msg, r.stderr)
    r = run(GUNZIP, str(tmp_path / "missing.gz"), str(tmp_path / "x"))
    assert r.returncode == 1 and "Input path does not exist" in r.stderr
    r = run(GUNZIP)
    assert r.returncode == 1 and r.stderr.startswith("Usage:")
    r = run(GZIP, str(tmp_path), str(tmp_path / "x.gz"))
    assert r.returncode == 1 and "Input path is a directory" in r.stderr


def test_adler32_on_gpu(b2d):
    import random
    rng = random.Random(1950)
    assert b2d.adler32(b"Wikipedia") == 0x11E60398
    for n in (0, 1, 15, 16, 17, 4095, 65536, (1 << 20) + 3, 5_000_001):
        d = rng.randbytes(n // 2) + b"\xff" * (n - n // 2)
        assert b2d.adler32(d) == zlib.adler32(d), n
        assert b2d.adler32(d, 0x12345678 % 65521 | (0x4321 << 16)) == zlib.adler32(d, 0x12345678 % 65521 | (0x4321 << 16)), n


ZPIPE = os.path.join(BIN, "zpipe")


@pytest.mark.parametrize("kind,n", [("text", 0), ("text", 100), ("mixed", (3 << 20) + 17), ("random", 1 << 20)])
def test_zlib_container_roundtrip_and_interop(b2d, oracle, tmp_path, kind, n):
    """ZlibOutputStream / ZlibInputStream of the host mirror against the oracle's restatement and
    Python's zlib, both directions."""
    data = b2d.corpus(kind, 0xDEF1A7E, n).tobytes()
    src, zz, back = tmp_path / "in.bin", tmp_path / "out.zz", tmp_path / "back.bin"
    src.write_bytes(data)
    r = run(ZPIPE, "-c", str(src), str(zz))
    assert r.returncode == 0, r.stderr
    blob = zz.read_bytes()
    assert blob[:2] == b"\x78\x9c"                                   # ZlibMetadata.DEFAULT: deflate, 32 KiB window, level DEFAULT
    assert zlib.decompress(blob) == data
    st, out, consumed = oracle.unzlib(blob, out_cap=n + 16)
    assert st == 0 and out == data and consumed == len(blob)
    r = run(ZPIPE, "-d", str(zz), str(back))
    assert r.returncode == 0 and back.read_bytes() == data
    foreign = tmp_path / "py.zz"
    foreign.write_bytes(zlib.compress(data, 6))
    r = run(ZPIPE, "-d", str(foreign), str(back))
    assert r.returncode == 0 and back.read_bytes() == data


def test_zlib_container_errors(b2d, tmp_path):
    z = zlib.compress(b2d.corpus("text", 3, 100000).tobytes(), 6)
    fix = (31 - ((0x77 << 8) % 31)) % 31
    cases = {
        "Header checksum mismatch": bytes([z[0], z[1] ^ 1]) + z[2:],
        "Unsupported compression method: 7": bytes([0x77, fix]) + z[2:],
        "Decompression Adler-32 mismatch": z[:-1] + bytes([z[-1] ^ 1]),
        "Unexpected end of stream": z[:len(z) // 2],
    }
    for msg, blob in cases.items():
        p = tmp_path / "bad.zz"
        p.write_bytes(blob)
        r = run(ZPIPE, "-d", str(p), str(tmp_path / "o"))
        assert r.returncode == 1 and msg in r.stderr, (msg, r.stderr)


def test_reference_stream_tests_on_host_mirror(b2d, tmp_path):
    """The reference's InflaterInputStreamTest harness (both read patterns, endExactly position check, all 39 vectors
    under 0-, 1- and random padding) and its five DeflaterOutputStreamTest round trips, run against the C++ host mirror
    (host/stream_tests.cpp), plus the constructor / state contract of the stream classes."""
    import random
    from util import bits_to_bytes, golden_vectors
    rng = random.Random(5)
    lines = []
    for v in golden_vectors():
        for pad in ("0", "1", "r"):
            data = bits_to_bytes(v["bits"], pad, rng)
            if v["expect"] == "ok":
                lines.append(f"{v['name']} {data.hex() or '-'} ok {v['output_hex'] or '-'}")
            else:
                lines.append(f"{v['name']} {data.hex() or '-'} fail {v['reason']}")
    p = tmp_path / "vectors.txt"
    p.write_text("\n".join(lines) + "\n")
    r = run(os.path.join(BIN, "stream_tests"), str(p))
    assert r.returncode == 0, r.stderr[-3000:] + r.stdout
    assert "passed" in r.stdout and "of" in r.stdout


@pytest.mark.parametrize("batch_mib", ["1", "2"])
def test_gzip_streams_the_file_in_batches(b2d, oracle, tmp_path, batch_mib):
    """bin/gzip reads, compresses and writes in batches (reader and writer threads next to the GPU call) and rewrites the
    header in place with the chunk index once the sizes are known: whatever the batch size, the file is the same member
    byte for byte, both gzip readers accept it, and bin/gunzip uses the index."""
    n = (5 << 20) + 54321
    data = b2d.corpus("mixed", 0xDEF1A7E + 5, n).tobytes()
    src, gz, back = tmp_path / "in.bin", tmp_path / f"out{batch_mib}.gz", tmp_path / "back.bin"
    src.write_bytes(data)
    os.utime(src, (1700000000, 1700000000))
    r = run(GZIP, str(src), str(gz), env={"B2D_GZIP_BATCH": batch_mib})
    assert r.returncode == 0, r.stderr
    member = gz.read_bytes()
    ref = tmp_path / "ref.gz"
    r = run(GZIP, str(src), str(ref), env={"B2D_GZIP_BATCH": "4096"})
    assert r.returncode == 0 and ref.read_bytes() == member                       # batching does not change a byte
    st, out, consumed = oracle.gunzip(member, out_cap=n + 16)
    assert st == 0 and out == data and consumed == len(member)
    assert run("gzip", "-t", str(gz)).returncode == 0
    xlen = member[10] | member[11] << 8
    extra = member[12:12 + xlen]
    assert extra[:2] == b"B2" and int.from_bytes(extra[2:4], "little") == 4 + 4 * 6    # six chunk sizes, none of them zero
    sizes = [int.from_bytes(extra[8 + 4 * i:12 + 4 * i], "little") for i in range(6)]
    assert all(sizes) and sum(sizes) == len(member) - (12 + xlen) - len(b"in.bin\0") - 2 - 8
    r = run(GUNZIP, str(gz), str(back))
    assert r.returncode == 0 and back.read_bytes() == data


def test_gunzip_decodes_a_system_gzip_file_in_parallel(b2d, tmp_path):
    """A file from system gzip is one DEFLATE stream without an index: bin/gunzip -> GzipInputStream -> InflaterInputStream
    -> b2d_inflate_stream (speculative parallel decode)."""
    data = b2d.corpus("text", 21, 20 << 20).tobytes()
    src, back = tmp_path / "big.txt", tmp_path / "big.out"
    src.write_bytes(data)
    subprocess.check_call(["gzip", "-k", "-6", str(src)])
    r = run(GUNZIP, str(src) + ".gz", str(back), env={"B2D_TRACE": "1"})
    assert r.returncode == 0, r.stderr
    assert back.read_bytes() == data
    assert "inflate_stream: result on the host" in r.stderr                  # the parallel path ran


@pytest.mark.parametrize("batch_chunks", ["1", "2"])
def test_gunzip_decodes_indexed_files_in_batches(b2d, tmp_path, batch_chunks):
    """bin/gunzip on a file with the block index decodes batch k + 1 on the GPU while batch k is being written: the output,
    the CRC / ISIZE verdict and the error convention are those of the one-shot decode, whatever the batch size -- also
    when a chunk in a later batch is damaged (the bytes in front of the damage are delivered, then the exception)."""
    n = (5 << 20) + 4321
    data = b2d.corpus("mixed", 0xDEF1A7E + 9, n).tobytes()
    src, gz, back = tmp_path / "in.bin", tmp_path / "out.gz", tmp_path / "back.bin"
    src.write_bytes(data)
    assert run(GZIP, str(src), str(gz)).returncode == 0
    env = {"B2D_GUNZIP_BATCH": batch_chunks}
    r = run(GUNZIP, str(gz), str(back), env=env)
    assert r.returncode == 0, r.stderr
    assert back.read_bytes() == data
    member = bytearray(gz.read_bytes())
    member[-6] ^= 1                                               # CRC-32 in the trailer
    bad = tmp_path / "badcrc.gz"
    bad.write_bytes(bytes(member))
    r = run(GUNZIP, str(bad), str(back), env=env)
    assert r.returncode == 1 and "Decompression CRC-32 mismatch" in r.stderr
    member = bytearray(gz.read_bytes())
    member[len(member) * 3 // 4] ^= 0x10                          # inside a chunk of a later batch
    bad2 = tmp_path / "badbody.gz"
    bad2.write_bytes(bytes(member))
    ref = run(GUNZIP, str(bad2), str(tmp_path / "ref.out"), env={"B2D_GUNZIP_BATCH": "100000"})
    r = run(GUNZIP, str(bad2), str(back), env=env)
    assert r.returncode == 1 and ref.returncode == 1
    assert r.stderr.splitlines()[-1] == ref.stderr.splitlines()[-1]                  # the same exception line
