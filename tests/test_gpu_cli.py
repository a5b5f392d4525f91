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


@pytest.mark.parametrize("kind,n", [("text", 0), ("text", 1), ("text", 70000), ("mixed", (5 << 20) + 12345), ("random", 3 << 20)])
@pytest.mark.parametrize("index", ["1", "0"])
def test_cli_roundtrip_and_interop(b2d, oracle, tmp_path, kind, n, index):
    data = b2d.corpus(kind, 0xDEF1A7E, n).tobytes()
    src, gz, back = tmp_path / "in.bin", tmp_path / "out.gz", tmp_path / "back.bin"
    src.write_bytes(data)
    r = run(GZIP, str(src), str(gz), env={"B2D_GZIP_INDEX": index})
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
