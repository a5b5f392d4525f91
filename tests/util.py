"""Shared helpers for the parity tests."""
import json
import os
import random
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))


def golden_vectors():
    with open(os.path.join(HERE, "golden", "inflate_vectors.json")) as f:
        return json.load(f)


def bits_to_bytes(bits, pad, rng=None):
    """StringInputStream packing (StringInputStream.java:40-47): LSB-first per byte; the harness pads the last byte
    with 0s, 1s or random bits (InflaterInputStreamTest.java:523-531)."""
    b = bits.replace(" ", "")
    while len(b) % 8:
        b += pad if pad in "01" else (rng or random).choice("01")
    return bytes(int(b[i:i + 8][::-1], 2) for i in range(0, len(b), 8))


def zlib_raw(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=-15, memlevel=8):
    c = zlib.compressobj(level, zlib.DEFLATED, wbits, memlevel, strategy)
    return c.compress(bytes(data)) + c.flush()


def zlib_inflate_raw(data):
    d = zlib.decompressobj(-15)
    out = d.decompress(bytes(data))
    assert d.eof, "zlib: stream did not end"
    return out, len(data) - len(d.unused_data)


class BitWriter:
    """LSB-first bit string builder for hand-made streams."""

    def __init__(self):
        self.bits = []

    def put(self, value, n):
        for i in range(n):
            self.bits.append((value >> i) & 1)

    def put_code(self, code, n):  # Huffman codes go MSB-of-code first
        for i in range(n - 1, -1, -1):
            self.bits.append((code >> i) & 1)

    def align(self):
        while len(self.bits) % 8:
            self.bits.append(0)

    def put_bytes(self, b):
        for x in b:
            self.put(x, 8)

    def tobytes(self):
        bits = self.bits + [0] * (-len(self.bits) % 8)
        return bytes(sum(bits[i + k] << k for k in range(8)) for i in range(0, len(bits), 8))


def fixed_lit_code(sym):
    """(code, nbits) of the fixed lit/len code (RFC 1951 3.2.6; Open.java:812-830)."""
    if sym < 144:
        return 0x30 + sym, 8
    if sym < 256:
        return 0x190 + sym - 144, 9
    if sym < 280:
        return sym - 256, 7
    return 0xC0 + sym - 280, 8
