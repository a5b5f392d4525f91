"""ctypes binding of libb2deflate.so (include/b2deflate.h) -- the harness-side twin of the Panama FFM binding
shown in INTEGRATION.md.  Every function here is a thin call through the C ABI; there is no CPU codec in this
package: if the library or a B200 is missing, calls raise (they never fall back).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libb2deflate.so")

# 1 + DataFormatException.Reason.ordinal() (DataFormatException.java:61-83)
REASONS = [
    "UNEXPECTED_END_OF_STREAM", "RESERVED_BLOCK_TYPE", "UNCOMPRESSED_BLOCK_LENGTH_MISMATCH",
    "HUFFMAN_CODE_UNDER_FULL", "HUFFMAN_CODE_OVER_FULL", "NO_PREVIOUS_CODE_LENGTH_TO_COPY",
    "CODE_LENGTH_CODE_OVER_FULL", "END_OF_BLOCK_CODE_ZERO_LENGTH", "RESERVED_LENGTH_SYMBOL",
    "RESERVED_DISTANCE_SYMBOL", "LENGTH_ENCOUNTERED_WITH_EMPTY_DISTANCE_CODE",
    "COPY_FROM_BEFORE_DICTIONARY_START", "HEADER_CHECKSUM_MISMATCH", "UNSUPPORTED_COMPRESSION_METHOD",
    "DECOMPRESSED_CHECKSUM_MISMATCH", "DECOMPRESSED_SIZE_MISMATCH", "GZIP_INVALID_MAGIC_NUMBER",
    "GZIP_RESERVED_FLAGS_SET", "GZIP_UNSUPPORTED_OPERATING_SYSTEM",
]
ERR_OUTPUT_OVERFLOW, ERR_BAD_ARGUMENT, ERR_NO_DEVICE, ERR_CUDA, ERR_OUT_OF_MEMORY = -1, -2, -3, -4, -5
INFLATE_CRC32, INFLATE_CHUNK_INDEXED, INFLATE_ADLER32 = 1, 2, 4
CHECKSUM_CRC32, CHECKSUM_ADLER32 = 0, 1
MODE_AUTO, MODE_STORED, MODE_FIXED, MODE_DYNAMIC = 0, 1, 2, 3
SEARCH_DEFAULT, SEARCH_LITERAL, SEARCH_RLE, SEARCH_FULL = 0, 1, 2, 3
FRAMING_CHUNKED, FRAMING_REFERENCE = 0, 1

EXPORTS = [
    "b2d_init", "b2d_shutdown", "b2d_strerror", "b2d_last_error", "b2d_device_sm_count", "b2d_alloc_pinned",
    "b2d_free_pinned", "b2d_inflate_batch", "b2d_inflate_batch_dev", "b2d_deflate_bound", "b2d_deflate_chunks",
    "b2d_deflate_chunks_dev", "b2d_crc32", "b2d_crc32_dev", "b2d_crc32_combine", "b2d_corpus_random",
    "b2d_corpus_text", "b2d_corpus_mixed", "b2d_gzip_isize", "b2d_gunzip_batch",
    "b2d_adler32", "b2d_adler32_combine", "b2d_crc32_update", "b2d_adler32_update",
    "b2d_deflate_chunks_indexed", "b2d_deflate_chunks_indexed_dev", "b2d_inflate_chunks", "b2d_inflate_chunks_dev",
    "b2d_init_devices", "b2d_device_count", "b2d_kernel_launches", "b2d_inflate_stream", "b2d_inflate_stream_dev",
]


def status_name(st):
    if st == 0:
        return "OK"
    if 1 <= st <= len(REASONS):
        return REASONS[st - 1]
    return {-1: "ERR_OUTPUT_OVERFLOW", -2: "ERR_BAD_ARGUMENT", -3: "ERR_NO_DEVICE", -4: "ERR_CUDA",
            -5: "ERR_OUT_OF_MEMORY"}.get(st, f"UNKNOWN({st})")


class B2dError(RuntimeError):
    def __init__(self, code, what=""):
        self.code = code
        super().__init__(f"{what}: {status_name(code)} ({code})")


class DeflateOpts(ctypes.Structure):
    _fields_ = [("chunk_bytes", ctypes.c_uint32), ("block_bytes", ctypes.c_uint32), ("mode", ctypes.c_int32),
                ("search", ctypes.c_int32), ("chain_depth", ctypes.c_int32), ("lazy", ctypes.c_int32),
                ("is_last", ctypes.c_int32), ("framing", ctypes.c_int32), ("checksum", ctypes.c_int32),
                ("split_min_bytes", ctypes.c_uint32)]


_lib = None


def lib():
    """Loads libb2deflate.so.  Raises if it has not been built (python deflate-library-java_b200/build.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(f"{SO_PATH} is missing: build it with `python deflate-library-java_b200/build.py` "
                          "(there is no CPU fallback)")
    L = ctypes.CDLL(SO_PATH)
    vp, u32, u64, i32 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int
    L.b2d_init.restype = i32
    L.b2d_init.argtypes = [i32]
    L.b2d_shutdown.restype = None
    L.b2d_init_devices.restype = i32
    L.b2d_init_devices.argtypes = [vp, i32]
    L.b2d_device_count.restype = i32
    L.b2d_kernel_launches.restype = u64
    L.b2d_strerror.restype = ctypes.c_char_p
    L.b2d_strerror.argtypes = [i32]
    L.b2d_last_error.restype = ctypes.c_char_p
    L.b2d_device_sm_count.restype = i32
    L.b2d_alloc_pinned.restype = vp
    L.b2d_alloc_pinned.argtypes = [ctypes.c_size_t]
    L.b2d_free_pinned.restype = None
    L.b2d_free_pinned.argtypes = [vp]
    L.b2d_inflate_batch.restype = i32
    L.b2d_inflate_batch.argtypes = [vp, vp, u32, vp, vp, vp, vp, vp, vp, u32]
    L.b2d_inflate_batch_dev.restype = i32
    L.b2d_inflate_batch_dev.argtypes = [vp, vp, u32, vp, vp, vp, vp, vp, vp, u32, vp]
    L.b2d_inflate_stream.restype = i32
    L.b2d_inflate_stream.argtypes = [vp, u64, vp, u64, vp, vp, vp, vp, u32, vp]
    L.b2d_inflate_stream_dev.restype = i32
    L.b2d_inflate_stream_dev.argtypes = [vp, u64, vp, u64, vp, vp]
    L.b2d_gzip_isize.restype = i32
    L.b2d_gzip_isize.argtypes = [vp, vp, u32, vp]
    L.b2d_gunzip_batch.restype = i32
    L.b2d_gunzip_batch.argtypes = [vp, vp, u32, vp, vp, vp, vp, vp]
    L.b2d_deflate_bound.restype = u64
    L.b2d_deflate_bound.argtypes = [u64, u32]
    L.b2d_deflate_chunks.restype = ctypes.c_int64
    L.b2d_deflate_chunks.argtypes = [vp, u64, ctypes.POINTER(DeflateOpts), vp, u64, vp, vp]
    L.b2d_deflate_chunks_dev.restype = i32
    L.b2d_deflate_chunks_dev.argtypes = [vp, u64, ctypes.POINTER(DeflateOpts), vp, u64, vp, vp, vp, vp]
    L.b2d_deflate_chunks_indexed.restype = ctypes.c_int64
    L.b2d_deflate_chunks_indexed.argtypes = [vp, u64, ctypes.POINTER(DeflateOpts), vp, u64, vp, vp, vp]
    L.b2d_deflate_chunks_indexed_dev.restype = i32
    L.b2d_deflate_chunks_indexed_dev.argtypes = [vp, u64, ctypes.POINTER(DeflateOpts), vp, u64, vp, vp, vp, vp, vp]
    L.b2d_inflate_chunks.restype = i32
    L.b2d_inflate_chunks.argtypes = [vp, vp, u32, vp, u32, u32, vp, u64, vp, vp, u32]
    L.b2d_inflate_chunks_dev.restype = i32
    L.b2d_inflate_chunks_dev.argtypes = [vp, vp, u32, vp, u32, u32, u64, vp, vp, vp, u32, vp]
    L.b2d_crc32.restype = u32
    L.b2d_crc32.argtypes = [u32, vp, u64]
    L.b2d_crc32_dev.restype = i32
    L.b2d_crc32_dev.argtypes = [vp, u64, vp, vp]
    L.b2d_crc32_combine.restype = u32
    L.b2d_crc32_combine.argtypes = [u32, u32, u64]
    L.b2d_crc32_update.restype = i32
    L.b2d_crc32_update.argtypes = [vp, u64, vp]
    L.b2d_adler32_update.restype = i32
    L.b2d_adler32_update.argtypes = [vp, u64, vp]
    L.b2d_adler32.restype = u32
    L.b2d_adler32.argtypes = [u32, vp, u64]
    L.b2d_adler32_combine.restype = u32
    L.b2d_adler32_combine.argtypes = [u32, u32, u64]
    for name in ("b2d_corpus_random", "b2d_corpus_text", "b2d_corpus_mixed"):
        f = getattr(L, name)
        f.restype = None
        f.argtypes = [u64, vp, ctypes.c_size_t]
    _lib = L
    return L


def _check(code, what):
    if code < 0:
        detail = lib().b2d_last_error().decode() if code == ERR_CUDA else ""
        raise B2dError(code, what + (f" [{detail}]" if detail else ""))
    return code


def init(device=0):
    _check(lib().b2d_init(device), "b2d_init")


def init_devices(devices=None):
    """Binds several GPUs (list of CUDA ordinals; None = all of the box); the host entry points then shard over them."""
    if not devices:
        _check(lib().b2d_init_devices(None, 0), "b2d_init_devices")
    else:
        arr = (ctypes.c_int * len(devices))(*devices)
        _check(lib().b2d_init_devices(arr, len(devices)), "b2d_init_devices")
    return int(lib().b2d_device_count())


def kernel_launches():
    return int(lib().b2d_kernel_launches())


def shutdown():
    lib().b2d_shutdown()


def strerror(st):
    return lib().b2d_strerror(st).decode()


def _u8(a):
    a = np.frombuffer(a, dtype=np.uint8) if isinstance(a, (bytes, bytearray, memoryview)) else a
    return np.ascontiguousarray(a, dtype=np.uint8)


def _ptr(a):
    return a.ctypes.data if a is not None and a.size else (a.ctypes.data if a is not None else None)


def make_opts(chunk_bytes=0, block_bytes=0, mode=MODE_AUTO, search=SEARCH_DEFAULT, chain_depth=0, lazy=-1, is_last=1,
              framing=FRAMING_CHUNKED, checksum=0, split_min_bytes=0):
    return DeflateOpts(chunk_bytes, block_bytes, mode, search, chain_depth, lazy, is_last, framing, checksum,
                       split_min_bytes)


# ---- corpora (host) ----
def corpus(kind, seed, n):
    out = np.empty(n, dtype=np.uint8)
    getattr(lib(), "b2d_corpus_" + kind)(seed, out.ctypes.data, n)
    return out


class PinnedBuffer:
    """Page-locked host memory from b2d_alloc_pinned as a numpy u8 array (`.array`); freed with the object.  Handing
    such a buffer to b2d_inflate_batch as the output lets the kernel deliver the bytes itself (no D2H copy)."""

    def __init__(self, n):
        self._p = lib().b2d_alloc_pinned(max(int(n), 1))
        if not self._p:
            raise B2dError(-5, "b2d_alloc_pinned")
        self.array = np.ctypeslib.as_array((ctypes.c_uint8 * max(int(n), 1)).from_address(self._p))

    def __del__(self):
        if getattr(self, "_p", None):
            lib().b2d_free_pinned(self._p)
            self._p = None


# ---- host-pointer entry points ----
def inflate_batch(members, out_caps, flags=0, out=None, pinned_in=False):
    """members: list of bytes-like raw-DEFLATE members; out_caps: int or list of output capacities.
    out: optional output array (e.g. PinnedBuffer(...).array); pinned_in: stage the members in page-locked memory.
    -> (outputs list[bytes], out_len, in_consumed, crc32, status) with numpy arrays for the last four."""
    n = len(members)
    if isinstance(out_caps, int):
        out_caps = [out_caps] * n
    in_off = np.zeros(n + 1, dtype=np.uint64)
    out_off = np.zeros(n + 1, dtype=np.uint64)
    for i, m in enumerate(members):
        in_off[i + 1] = in_off[i] + np.uint64(len(m))
        out_off[i + 1] = out_off[i] + np.uint64(out_caps[i])
    blob = np.frombuffer(b"".join(bytes(m) for m in members), dtype=np.uint8) if n else np.zeros(0, np.uint8)
    keep = None
    if pinned_in and blob.size:
        keep = PinnedBuffer(blob.size + 5)
        keep.array[1:1 + blob.size] = blob                       # odd start address on purpose
        blob = keep.array[1:1 + blob.size]
    out, out_len, in_consumed, crc, status = inflate_batch_raw(blob, in_off, out_off, flags, out)
    outs = [bytes(out[int(out_off[i]):int(out_off[i]) + int(out_len[i])]) for i in range(n)]
    return outs, out_len, in_consumed, crc, status


def inflate_batch_raw(blob, in_off, out_off, flags=0, out=None):
    """Contiguous form: blob (u8), in_off/out_off (u64[n+1]).  -> (out u8, out_len, in_consumed, crc32, status)."""
    blob = _u8(blob)
    n = len(in_off) - 1
    in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
    out_off = np.ascontiguousarray(out_off, dtype=np.uint64)
    total = int(out_off[n]) if n >= 0 else 0
    if out is None:
        out = np.zeros(max(total, 1), dtype=np.uint8)
    out_len = np.zeros(max(n, 1), dtype=np.uint64)
    in_consumed = np.zeros(max(n, 1), dtype=np.uint64)
    crc = np.zeros(max(n, 1), dtype=np.uint32)
    status = np.zeros(max(n, 1), dtype=np.int32)
    r = lib().b2d_inflate_batch(blob.ctypes.data if blob.size else None, in_off.ctypes.data, n, out.ctypes.data,
                                out_off.ctypes.data, out_len.ctypes.data, in_consumed.ctypes.data,
                                crc.ctypes.data, status.ctypes.data, flags)
    _check(r, "b2d_inflate_batch")
    return out, out_len[:n], in_consumed[:n], crc[:n], status[:n]


def inflate_stream(data, out_cap, flags=INFLATE_CRC32):
    """ONE raw-DEFLATE stream of any origin (speculative parallel decode with a sequential fallback).
    -> (output bytes, in_consumed, checksum, status, parallel)."""
    data = _u8(data)
    out = np.zeros(max(int(out_cap), 1), dtype=np.uint8)
    out_len, cons = ctypes.c_uint64(0), ctypes.c_uint64(0)
    crc, st, par = ctypes.c_uint32(0), ctypes.c_int32(0), ctypes.c_int32(0)
    r = lib().b2d_inflate_stream(data.ctypes.data if data.size else None, data.size, out.ctypes.data, int(out_cap), ctypes.byref(out_len),
                                 ctypes.byref(cons), ctypes.byref(crc), ctypes.byref(st), flags, ctypes.byref(par))
    _check(r, "b2d_inflate_stream")
    return out[:out_len.value], cons.value, crc.value, st.value, par.value


def deflate_bound(n, chunk_bytes=0):
    return int(lib().b2d_deflate_bound(n, chunk_bytes))


def deflate_chunks(data, opts=None, crc=None, want_index=False):
    """-> compressed bytes (np.uint8 array) [, crc32, chunk_out_len]."""
    data = _u8(data)
    opts = opts if opts is not None else make_opts()
    cap = deflate_bound(data.size, opts.chunk_bytes)
    out = np.empty(cap, dtype=np.uint8)
    cb = opts.chunk_bytes or (1 << 20)
    n_chunks = max(1, (data.size + cb - 1) // cb)
    idx = np.zeros(n_chunks, dtype=np.uint64)
    c = ctypes.c_uint32(crc if crc is not None else 0)
    r = lib().b2d_deflate_chunks(data.ctypes.data if data.size else None, data.size, ctypes.byref(opts),
                                 out.ctypes.data, cap, ctypes.byref(c), idx.ctypes.data)
    _check(int(r), "b2d_deflate_chunks")
    res = out[:int(r)]
    if want_index or crc is not None:
        return res, c.value, idx
    return res


def crc32(data, crc=0):
    data = _u8(data)
    v = ctypes.c_uint32(crc)
    _check(lib().b2d_crc32_update(data.ctypes.data if data.size else None, data.size, ctypes.byref(v)), "b2d_crc32_update")
    return v.value


def adler32(data, adler=1):
    data = _u8(data)
    v = ctypes.c_uint32(adler)
    _check(lib().b2d_adler32_update(data.ctypes.data if data.size else None, data.size, ctypes.byref(v)), "b2d_adler32_update")
    return v.value


def adler32_combine(a, b, len_b):
    return int(lib().b2d_adler32_combine(a, b, len_b))


def crc32_combine(a, b, len_b):
    return int(lib().b2d_crc32_combine(a, b, len_b))


def gunzip_batch(members, out_caps=None, pinned_out=False):
    """members: list of complete gzip members.  out_caps: capacities (default: each member's ISIZE).  pinned_out: decode
    into page-locked memory (the kernel then delivers the bytes itself, no D2H copy).
    -> (outputs list[bytes], out_len, in_consumed, status)."""
    n = len(members)
    in_off = np.zeros(n + 1, dtype=np.uint64)
    for i, m in enumerate(members):
        in_off[i + 1] = in_off[i] + np.uint64(len(m))
    blob = np.frombuffer(b"".join(bytes(m) for m in members), dtype=np.uint8) if n else np.zeros(0, np.uint8)
    if out_caps is None:
        caps = np.zeros(max(n, 1), dtype=np.uint64)
        _check(lib().b2d_gzip_isize(blob.ctypes.data if blob.size else None, in_off.ctypes.data, n, caps.ctypes.data), "b2d_gzip_isize")
        out_caps = [int(c) for c in caps[:n]]
    elif isinstance(out_caps, int):
        out_caps = [out_caps] * n
    out_off = np.zeros(n + 1, dtype=np.uint64)
    for i in range(n):
        out_off[i + 1] = out_off[i] + np.uint64(out_caps[i])
    keep = PinnedBuffer(max(int(out_off[n]), 1)) if pinned_out else None
    out = keep.array if pinned_out else np.zeros(max(int(out_off[n]), 1), dtype=np.uint8)
    out_len = np.zeros(max(n, 1), dtype=np.uint64)
    consumed = np.zeros(max(n, 1), dtype=np.uint64)
    status = np.zeros(max(n, 1), dtype=np.int32)
    r = lib().b2d_gunzip_batch(blob.ctypes.data if blob.size else None, in_off.ctypes.data, n, out.ctypes.data, out_off.ctypes.data,
                               out_len.ctypes.data, consumed.ctypes.data, status.ctypes.data)
    _check(r, "b2d_gunzip_batch")
    outs = [bytes(out[int(out_off[i]):int(out_off[i]) + int(out_len[i])]) for i in range(n)]
    return outs, out_len[:n], consumed[:n], status[:n]


def deflate_chunks_indexed(data, opts=None, crc=0):
    """-> (compressed np.uint8, checksum, chunk sizes u64[], block bit offsets u32[])."""
    data = _u8(data)
    opts = opts if opts is not None else make_opts()
    cb, bb = opts.chunk_bytes or (1 << 20), opts.block_bytes or (1 << 16)
    cap = deflate_bound(data.size, cb)
    out = np.empty(cap, dtype=np.uint8)
    idx = np.zeros(max(1, (data.size + cb - 1) // cb), dtype=np.uint64)
    bits = np.zeros(max(1, (data.size + bb - 1) // bb), dtype=np.uint32)
    c = ctypes.c_uint32(crc)
    r = lib().b2d_deflate_chunks_indexed(data.ctypes.data if data.size else None, data.size, ctypes.byref(opts), out.ctypes.data,
                                         cap, ctypes.byref(c), idx.ctypes.data, bits.ctypes.data)
    _check(int(r), "b2d_deflate_chunks_indexed")
    return out[:int(r)], c.value, idx[:(data.size + cb - 1) // cb], bits[:(data.size + bb - 1) // bb]


def inflate_chunks(comp, chunk_sizes, block_bits, out_total, chunk_bytes=1 << 20, block_bytes=1 << 16, flags=INFLATE_CRC32):
    """-> (out np.uint8[out_total], per-chunk checksum u32[], per-chunk status i32[])."""
    comp = _u8(comp)
    sizes = np.ascontiguousarray(chunk_sizes, dtype=np.uint64)
    bits = np.ascontiguousarray(block_bits, dtype=np.uint32)
    n = len(sizes)
    out = np.zeros(max(out_total, 1), dtype=np.uint8)
    crc = np.zeros(max(n, 1), dtype=np.uint32)
    st = np.zeros(max(n, 1), dtype=np.int32)
    r = lib().b2d_inflate_chunks(comp.ctypes.data if comp.size else None, sizes.ctypes.data, n, bits.ctypes.data, chunk_bytes,
                                 block_bytes, out.ctypes.data, out_total, crc.ctypes.data, st.ctypes.data, flags)
    _check(r, "b2d_inflate_chunks")
    return out[:out_total], crc[:n], st[:n]
