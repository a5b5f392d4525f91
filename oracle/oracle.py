"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product package (deflate-library-java_b200) never does.
"""
import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "build", "liboracle.so")
_lock = threading.Lock()
_lib = None

# DataFormatException.Reason (DataFormatException.java:61-83); status = 1 + ordinal
REASONS = [
    "UNEXPECTED_END_OF_STREAM", "RESERVED_BLOCK_TYPE", "UNCOMPRESSED_BLOCK_LENGTH_MISMATCH",
    "HUFFMAN_CODE_UNDER_FULL", "HUFFMAN_CODE_OVER_FULL", "NO_PREVIOUS_CODE_LENGTH_TO_COPY",
    "CODE_LENGTH_CODE_OVER_FULL", "END_OF_BLOCK_CODE_ZERO_LENGTH", "RESERVED_LENGTH_SYMBOL",
    "RESERVED_DISTANCE_SYMBOL", "LENGTH_ENCOUNTERED_WITH_EMPTY_DISTANCE_CODE",
    "COPY_FROM_BEFORE_DICTIONARY_START", "HEADER_CHECKSUM_MISMATCH", "UNSUPPORTED_COMPRESSION_METHOD",
    "DECOMPRESSED_CHECKSUM_MISMATCH", "DECOMPRESSED_SIZE_MISMATCH", "GZIP_INVALID_MAGIC_NUMBER",
    "GZIP_RESERVED_FLAGS_SET", "GZIP_UNSUPPORTED_OPERATING_SYSTEM",
]
OUTPUT_OVERFLOW = -1

LITERAL_STATIC, LITERAL_DYNAMIC, RLE_STATIC, RLE_DYNAMIC, FULL_STATIC, FULL_DYNAMIC, UNCOMPRESSED = range(7)


def status_name(st):
    if st == 0:
        return "OK"
    if 1 <= st <= len(REASONS):
        return REASONS[st - 1]
    return {-1: "OUTPUT_OVERFLOW", -2: "BAD_ARGUMENT"}.get(st, f"UNKNOWN({st})")


def build(force=False):
    """Compile oracle/build/liboracle.so with gcc (a few seconds)."""
    srcs = [os.path.join(_HERE, f) for f in ("oracle_inflate.c", "oracle_deflate.c", "oracle_misc.c", "oracle.h", "Makefile",
                                             os.path.join("..", "deflate-library-java_b200", "csrc", "corpus.c"))]
    if not force and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "-s", "-B" if force else "-s"])
    return _SO


def lib():
    global _lib
    with _lock:
        if _lib is None:
            build()
            L = ctypes.CDLL(_SO)
            u8p, szp = ctypes.c_char_p, ctypes.POINTER(ctypes.c_size_t)
            for name in ("oracle_inflate", "oracle_inflate_slow", "oracle_gunzip", "oracle_unzlib"):
                f = getattr(L, name)
                f.restype = ctypes.c_int
                f.argtypes = [u8p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, szp, szp]
            L.oracle_deflate.restype = ctypes.c_size_t
            L.oracle_deflate.argtypes = [u8p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int), ctypes.c_int,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]
            L.oracle_deflate_split.restype = ctypes.c_size_t
            L.oracle_deflate_split.argtypes = [u8p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int), ctypes.c_int,
                                               ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                               ctypes.c_size_t, szp]
            L.oracle_deflate_bound.restype = ctypes.c_size_t
            L.oracle_deflate_bound.argtypes = [ctypes.c_size_t, ctypes.c_int]
            L.oracle_package_merge.restype = None
            L.oracle_package_merge.argtypes = [ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
            L.oracle_adler32.restype = ctypes.c_uint32
            L.oracle_adler32.argtypes = [ctypes.c_uint32, u8p, ctypes.c_size_t]
            L.oracle_crc32.restype = ctypes.c_uint32
            L.oracle_crc32.argtypes = [ctypes.c_uint32, u8p, ctypes.c_size_t]
            L.oracle_gzip_header.restype = ctypes.c_size_t
            L.oracle_gzip_header.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_uint32,
                                             ctypes.c_char_p, ctypes.c_size_t]
            L.oracle_gzip_parse_header.restype = ctypes.c_int
            L.oracle_gzip_parse_header.argtypes = [u8p, ctypes.c_size_t, szp]
            for name in ("b2d_corpus_random", "b2d_corpus_text", "b2d_corpus_mixed"):   # SURVEY Appendix D generators
                f = getattr(L, name)
                f.restype = None
                f.argtypes = [ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t]
            _lib = L
    return _lib


def _inflate(fn, data, out_cap):
    data = bytes(data)
    out = ctypes.create_string_buffer(max(out_cap, 1))
    out_len, consumed = ctypes.c_size_t(0), ctypes.c_size_t(0)
    st = fn(data, len(data), out, out_cap, ctypes.byref(out_len), ctypes.byref(consumed))
    return st, out.raw[:out_len.value], consumed.value


def inflate(data, out_cap=1 << 20, slow=False):
    """-> (status, output bytes, consumed input bytes).  Open.java semantics."""
    L = lib()
    return _inflate(L.oracle_inflate_slow if slow else L.oracle_inflate, data, out_cap)


def gunzip(data, out_cap=1 << 20):
    """-> (status, output bytes, consumed input bytes).  GzipInputStream semantics (first member only)."""
    return _inflate(lib().oracle_gunzip, data, out_cap)


def unzlib(data, out_cap=1 << 20):
    """-> (status, output bytes, consumed input bytes).  ZlibInputStream semantics."""
    return _inflate(lib().oracle_unzlib, data, out_cap)


def adler32(data, adler=1):
    data = bytes(data)
    return lib().oracle_adler32(adler, data, len(data))


def deflate(data, strategies=(RLE_DYNAMIC,), lookahead=64 * 1024, history=32 * 1024, brute_force=False):
    """DeflaterOutputStream(out, lookahead, history, strategy).write(data).finish() -> raw DEFLATE bytes.
    Several strategies = MultiStrategy(strategies...)."""
    L = lib()
    data = bytes(data)
    if isinstance(strategies, int):
        strategies = (strategies,)
    arr = (ctypes.c_int * len(strategies))(*strategies)
    cap = L.oracle_deflate_bound(len(data), lookahead)
    out = ctypes.create_string_buffer(cap)
    n = L.oracle_deflate(data, len(data), arr, len(strategies), lookahead, history, 1 if brute_force else 0, out, cap)
    if n == ctypes.c_size_t(-1).value:
        raise RuntimeError("oracle_deflate failed (bad arguments or bound too small)")
    return out.raw[:n]


def deflate_split(data, strategies=(RLE_DYNAMIC,), min_block_len=4096, lookahead=64 * 1024, history=32 * 1024,
                  brute_force=False):
    """BinarySplit(substrategy, min_block_len) as the stream's strategy -> (compressed bytes, blocks chosen)."""
    data = bytes(data)
    L = lib()
    strat = (ctypes.c_int * len(strategies))(*strategies)
    cap = L.oracle_deflate_bound(len(data), min(lookahead, max(min_block_len, 1))) + len(data) // max(min_block_len, 1) * 400
    out = ctypes.create_string_buffer(cap)
    nb = ctypes.c_size_t(0)
    n = L.oracle_deflate_split(data, len(data), strat, len(strategies), lookahead, history, 1 if brute_force else 0,
                               min_block_len, out, cap, ctypes.byref(nb))
    if n == ctypes.c_size_t(-1).value:
        raise ValueError("oracle_deflate_split failed")
    return out.raw[:n], nb.value


def package_merge(hist, max_len):
    L = lib()
    n = len(hist)
    arr = (ctypes.c_int * n)(*hist)
    out = ctypes.create_string_buffer(max(n, 1))
    L.oracle_package_merge(arr, n, max_len, out)
    return list(out.raw[:n])


def crc32(data, crc=0):
    data = bytes(data)
    return lib().oracle_crc32(crc, data, len(data))


def gzip_header(file_name=None, mtime=0, extra=None):
    buf = ctypes.create_string_buffer(70000 + (len(file_name) if file_name else 0))
    n = lib().oracle_gzip_header(buf, len(buf), file_name.encode("latin-1") if file_name is not None else None,
                                 mtime, extra, len(extra) if extra else 0)
    return buf.raw[:n]


def gzip_parse_header(data):
    data = bytes(data)
    hl = ctypes.c_size_t(0)
    st = lib().oracle_gzip_parse_header(data, len(data), ctypes.byref(hl))
    return st, hl.value


def gzip_member(data, file_name="data", mtime=0, strategies=(RLE_DYNAMIC,), extra=None):
    """What `java gzip In Out.gz` writes (gzip.java:52-68): header, raw DEFLATE, CRC-32, ISIZE (both LE)."""
    import struct
    body = deflate(data, strategies)
    return gzip_header(file_name, mtime, extra) + body + struct.pack("<II", crc32(data), len(data) & 0xFFFFFFFF)
