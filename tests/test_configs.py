"""The BASELINE.json configurations as tests: config 1 on the CPU (oracle = restated reference CLIs), configs 2/3/5 on
the GPU at full or near-full size through size-independent properties (encode -> decode round trip, a checksum of
checksums against zlib, chunk sizes that sum to the stream length, stored-block size formula)."""
import hashlib
import subprocess
import zlib

import numpy as np
import pytest

SEED = 0xDEF1A7E


def test_config1_cpu_gzip_gunzip_roundtrip_64mib(oracle, b2d_nogpu, tmp_path):
    """configs[0]: 64 MiB text-like corpus through gzip.java then gunzip.java on the host CPU (here: the oracle's
    restatement of DeflaterOutputStream(RLE_DYNAMIC) + GzipMetadata/GzipOutputStream, then GzipInputStream), byte-exact."""
    n = 64 << 20
    data = b2d_nogpu.corpus("text", SEED, n).tobytes()
    member = oracle.gzip_member(data, file_name="corpus.txt", mtime=1700000000)
    st, out, consumed = oracle.gunzip(member, out_cap=n + 64)
    assert st == 0 and consumed == len(member)
    assert hashlib.sha256(out).digest() == hashlib.sha256(data).digest()
    p = tmp_path / "corpus.txt.gz"
    p.write_bytes(member)
    assert subprocess.run(["gzip", "-t", str(p)]).returncode == 0
    assert 1.5 < n / len(member) < 2.2                        # RLE_DYNAMIC on G_TEXT is Huffman-only territory


def _roundtrip_gpu(b2d, data, opts, chunk):
    comp, crc, idx = b2d.deflate_chunks(data, opts, crc=0)
    assert int(idx.sum()) == comp.size
    n_chunks = len(idx)
    in_off = np.zeros(n_chunks + 1, np.uint64)
    in_off[1:] = np.cumsum(idx)
    out_off = np.arange(n_chunks + 1, dtype=np.uint64) * np.uint64(chunk)
    out, out_len, consumed, crcs, status = b2d.inflate_batch_raw(comp, in_off, out_off, b2d.INFLATE_CHUNK_INDEXED | b2d.INFLATE_CRC32)
    assert not status.any(), [b2d.status_name(int(s)) for s in status[status != 0][:4]]
    assert np.array_equal(consumed, idx)
    assert int(out_len.sum()) == data.size and np.array_equal(out[:data.size], data)
    c = 0
    for i in range(n_chunks):                                  # checksum of checksums
        c = b2d.crc32_combine(c, int(crcs[i]), int(out_len[i]))
    assert c == crc
    return comp, crc, idx


def _oracle_sample(oracle, data, comp, idx, chunk, n_sample=64):
    """A strided sample of the stream's chunks through the ORACLE's decoder (the restated Open.java): status, bytes and
    consumed input of every sampled chunk.  The last chunk of a stream carries the final marker; the others end in a
    non-final empty stored block, after which the reference decoder asks for the next block header and reports the end
    of input -- having delivered the whole chunk."""
    n_chunks = len(idx)
    off = np.zeros(n_chunks + 1, np.int64)
    off[1:] = np.cumsum(idx)
    picked = sorted(set(list(range(0, n_chunks, max(1, n_chunks // n_sample))) + [n_chunks - 1]))
    for c in picked:
        body = comp[int(off[c]):int(off[c + 1])].tobytes()
        want = data[c * chunk:min(data.size, (c + 1) * chunk)].tobytes()
        st, out, consumed = oracle.inflate(body, out_cap=len(want) + 8)
        assert out == want, f"oracle decode of chunk {c} differs"
        if c == n_chunks - 1:
            assert st == 0 and consumed == len(body), (c, st, consumed, len(body))
        else:
            assert st == 1, (c, st)                        # UNEXPECTED_END_OF_STREAM after the whole chunk: no BFINAL yet
    return len(picked)


@pytest.mark.gpu
def test_config3_chunked_deflate_1gib(b2d, oracle):
    """configs[2]: 1 GiB G_MIXED, 1 MiB chunks + sync-flush markers + CRC-32."""
    n = 1 << 30
    data = np.concatenate([b2d.corpus("mixed", SEED + k, 64 << 20) for k in range(16)])
    comp, crc, idx = _roundtrip_gpu(b2d, data, b2d.make_opts(), 1 << 20)
    assert crc == zlib.crc32(data.data)
    assert 2.3 < n / comp.size < 3.3
    assert _oracle_sample(oracle, data, comp, idx, 1 << 20) >= 64
    for c in (0, 333, 1023):                                   # zlib (the java.util.zip.Inflater stand-in) reads chunks alone
        off = int(idx[:c].sum())
        d = zlib.decompressobj(-15)
        assert d.decompress(comp[off:off + int(idx[c])].tobytes()) == data[c << 20:(c + 1) << 20].tobytes()


@pytest.mark.gpu
def test_config2_batch_inflate_4096_members(b2d, oracle):
    """configs[1]: 4096 independent 256 KiB members (here encoded by the GPU encoder, one stream per member)."""
    n, size = 4096, 256 * 1024
    data = np.concatenate([b2d.corpus("text", SEED + 1000 * k, 64 << 20) for k in range(16)])
    comp, crc, idx = _roundtrip_gpu(b2d, data, b2d.make_opts(chunk_bytes=size, is_last=1), size)
    assert len(idx) == n
    assert _oracle_sample(oracle, data, comp, idx, size) >= 64


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["random", "zeros", "fixed"])
def test_config5_edge_stress(b2d, oracle, kind):
    """configs[4] at 512 MiB per case: incompressible bytes -> stored blocks, zeros -> 258/distance-1 runs,
    fixed-Huffman-only streams."""
    n = 512 << 20
    if kind == "random":
        data = np.concatenate([b2d.corpus("random", SEED + k, 64 << 20) for k in range(8)])
        comp, crc, idx = _roundtrip_gpu(b2d, data, b2d.make_opts(), 1 << 20)
        assert comp.size == n + 10 * (n >> 16) + 5 * (n >> 20)     # two stored pieces per 64 KiB block + a marker per chunk
    elif kind == "zeros":
        data = np.zeros(n, np.uint8)
        comp, crc, idx = _roundtrip_gpu(b2d, data, b2d.make_opts(), 1 << 20)
        assert comp.size < n // 800
    else:
        data = np.concatenate([b2d.corpus("text", SEED + k, 64 << 20) for k in range(8)])
        comp, crc, idx = _roundtrip_gpu(b2d, data, b2d.make_opts(mode=b2d.MODE_FIXED), 1 << 20)
        off = 0
        for c in range(0, len(idx), 97):                      # every chunk starts with BFINAL=0, BTYPE=01
            assert comp[int(idx[:c].sum())] & 7 == 0b010
    assert crc == zlib.crc32(data.data)
    assert _oracle_sample(oracle, data, comp, idx, 1 << 20, 32) >= 32
