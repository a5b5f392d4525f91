"""CPU: the C-ABI library loads, exports every symbol include/b2deflate.h declares, and refuses to compute
without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b2deflate.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2d_\w+)\s*\(", text)))


def test_header_declares_the_boundary():
    syms = _declared_symbols()
    for must in ("b2d_init", "b2d_inflate_batch", "b2d_deflate_chunks", "b2d_crc32", "b2d_crc32_combine",
                 "b2d_alloc_pinned", "b2d_strerror"):
        assert must in syms


def test_library_exports_every_declared_symbol(b2d_nogpu):
    L = b2d_nogpu.lib()
    for s in _declared_symbols():
        assert hasattr(L, s), f"libb2deflate.so does not export {s}"


def test_strerror_matches_reference_messages(b2d_nogpu):
    # message strings of Open.java / DataFormatException call sites
    assert b2d_nogpu.strerror(1) == "Unexpected end of stream"
    assert b2d_nogpu.strerror(2) == "Reserved block type"
    assert b2d_nogpu.strerror(3) == "len/nlen mismatch in uncompressed block"
    assert b2d_nogpu.strerror(6) == "No code length value to copy"
    assert b2d_nogpu.strerror(7) == "Run exceeds number of codes"
    assert b2d_nogpu.strerror(8) == "End-of-block symbol has zero code length"
    assert b2d_nogpu.strerror(11) == "Length symbol encountered with empty distance code"
    assert len(b2d_nogpu.REASONS) == 19


def test_crc32_combine_is_host_math(b2d_nogpu, oracle):
    import random
    rng = random.Random(1)
    a, b = rng.randbytes(1234), rng.randbytes(77777)
    assert b2d_nogpu.crc32_combine(oracle.crc32(a), oracle.crc32(b), len(b)) == oracle.crc32(a + b)
    assert b2d_nogpu.crc32_combine(oracle.crc32(a), 0, 0) == oracle.crc32(a)


def test_corpus_generators_are_deterministic(b2d_nogpu):
    t1 = b2d_nogpu.corpus("text", 0xDEF1A7E, 100000)
    t2 = b2d_nogpu.corpus("text", 0xDEF1A7E, 100000)
    assert np.array_equal(t1, t2) and t1[:50000].tobytes() == b2d_nogpu.corpus("text", 0xDEF1A7E, 50000).tobytes()
    m = b2d_nogpu.corpus("mixed", 0xDEF1A7E, 1 << 20)
    assert m.size == 1 << 20
    r = b2d_nogpu.corpus("random", 1, 4096)
    assert len(set(r.tolist())) > 200


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="GPU present")
def test_no_cpu_fallback(b2d_nogpu):
    L = b2d_nogpu.lib()
    assert L.b2d_init(0) == b2d_nogpu.ERR_NO_DEVICE
    out = np.zeros(64, np.uint8)
    offs = np.array([0, 2], np.uint64)
    ooffs = np.array([0, 64], np.uint64)
    scratch = np.zeros(4, np.uint64)
    st = np.zeros(1, np.int32)
    data = np.frombuffer(b"\x03\x00", np.uint8)
    r = L.b2d_inflate_batch(data.ctypes.data, offs.ctypes.data, 1, out.ctypes.data, ooffs.ctypes.data,
                            scratch.ctypes.data, scratch.ctypes.data + 8, None, st.ctypes.data, 0)
    assert r == b2d_nogpu.ERR_NO_DEVICE
    opts = b2d_nogpu.make_opts()
    r = L.b2d_deflate_chunks(data.ctypes.data, 2, ctypes.byref(opts), out.ctypes.data, 64, None, None)
    assert r == b2d_nogpu.ERR_NO_DEVICE
    with pytest.raises(b2d_nogpu.B2dError):
        b2d_nogpu.init(0)
    with pytest.raises(b2d_nogpu.B2dError):
        b2d_nogpu.crc32(b"123456789")                 # checksums too: no device, no answer
    with pytest.raises(b2d_nogpu.B2dError):
        b2d_nogpu.adler32(b"123456789")


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="GPU present")
def test_cli_fails_loudly_without_gpu(tmp_path):
    import subprocess
    exe = os.path.join(ROOT, "deflate-library-java_b200", "bin", "gzip")
    if not os.path.exists(exe):
        pytest.skip("host CLIs not built")
    src = tmp_path / "a.txt"
    src.write_bytes(b"hello")
    r = subprocess.run([exe, str(src), str(tmp_path / "a.gz")], capture_output=True, text=True)
    assert r.returncode == 1 and "No usable sm_100 GPU" in r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr.startswith("Usage:")


def test_adler32_combine_is_host_math(b2d_nogpu):
    import random
    import zlib
    rng = random.Random(2)
    for la, lb in ((0, 0), (1, 0), (0, 5), (1234, 77777), (65521, 65521 * 3 + 7)):
        a, b = rng.randbytes(la), rng.randbytes(lb)
        assert b2d_nogpu.adler32_combine(zlib.adler32(a), zlib.adler32(b), len(b)) == zlib.adler32(a + b)
