"""CPU: the LZ77 layer of the restated encoder, checked token by token against the reference's rule.

The reference holds no golden compressed bytes, so the encoder restatement is pinned by properties.  This
one is independent of the C code: a small Python inflate splits the oracle's streams into blocks and tokens, and every
token must be what comp/Lz77Huffman.java:68-92 picks at that position -- the longest run over the distances
searchMinimumDistance .. min(searchMaximumDistance, bytes available behind the position: the block's history plus what
the block has produced), ties to the smaller distance, runs cut at the block's end and at 258, a literal when the best
run is shorter than 3; one block per `lookahead` bytes (DeflaterOutputStream.java:119-137)."""
import random

import pytest

LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
LEN_EXTRA = [0] * 8 + [1] * 4 + [2] * 4 + [3] * 4 + [4] * 4 + [5] * 4 + [0]
DIST_BASE = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145,
             8193, 12289, 16385, 24577]
DIST_EXTRA = [0, 0, 0, 0] + [i // 2 for i in range(2, 28)]
CL_ORDER = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]


class Bits:
    def __init__(self, data):
        self.d, self.pos = data, 0

    def get(self, n):
        v = 0
        for i in range(n):
            v |= ((self.d[self.pos >> 3] >> (self.pos & 7)) & 1) << i
            self.pos += 1
        return v


def canonical(lengths):
    """code lengths -> {(nbits, code): symbol} (RFC 1951 3.2.2)."""
    table, code = {}, 0
    for n in range(1, 16):
        for sym, ln in enumerate(lengths):
            if ln == n:
                table[(n, code)] = sym
                code += 1
        code <<= 1
    return table


def read_symbol(bits, table):
    code = 0
    for n in range(1, 16):
        code = code << 1 | bits.get(1)
        if (n, code) in table:
            return table[(n, code)]
    raise AssertionError("no such code")


def tokens_per_block(stream):
    """-> [(btype, [literal byte | (length, distance), ...])] and the decoded bytes."""
    bits, out, blocks = Bits(stream), bytearray(), []
    while True:
        final, btype = bits.get(1), bits.get(2)
        toks = []
        if btype == 0:
            bits.pos = (bits.pos + 7) & ~7
            n, nn = bits.get(16), bits.get(16)
            assert n ^ nn == 0xFFFF
            for _ in range(n):
                toks.append(bits.get(8))
                out.append(toks[-1])
        else:
            if btype == 1:
                lit = canonical([8] * 144 + [9] * 112 + [7] * 24 + [8] * 8)
                dist = canonical([5] * 32)
            else:
                hlit, hdist, hclen = bits.get(5) + 257, bits.get(5) + 1, bits.get(4) + 4
                cl = [0] * 19
                for i in range(hclen):
                    cl[CL_ORDER[i]] = bits.get(3)
                clt, lens = canonical(cl), []
                while len(lens) < hlit + hdist:
                    s = read_symbol(bits, clt)
                    if s < 16:
                        lens.append(s)
                    elif s == 16:
                        lens += [lens[-1]] * (3 + bits.get(2))
                    else:
                        lens += [0] * (3 + bits.get(3) if s == 17 else 11 + bits.get(7))
                assert len(lens) == hlit + hdist
                lit, dist = canonical(lens[:hlit]), canonical(lens[hlit:])
            while True:
                s = read_symbol(bits, lit)
                if s < 256:
                    toks.append(s)
                    out.append(s)
                elif s == 256:
                    break
                else:
                    ln = LEN_BASE[s - 257] + bits.get(LEN_EXTRA[s - 257])
                    ds = read_symbol(bits, dist)
                    d = DIST_BASE[ds] + bits.get(DIST_EXTRA[ds])
                    toks.append((ln, d))
                    for _ in range(ln):
                        out.append(out[-d])
        blocks.append((btype, toks))
        if final:
            return blocks, bytes(out)


def greedy(data, index, block_end, avail, min_dist, max_dist):
    """Lz77Huffman.java:71-86 at `index`: (run, dist) of the best candidate, run 0 if there is none."""
    best_run = best_dist = 0
    for dist in range(min_dist, min(max_dist, avail) + 1):
        if best_run >= 258:
            break
        run = 0
        while run < 258 and index + run < block_end and data[index + run] == data[index + run - dist]:
            run += 1               # (the reference wraps its history index at `index`: the same bytes, data[i] == data[i - dist])
        if run > best_run:         # ascending distances: a tie keeps the smaller one
            best_run, best_dist = run, dist
    return best_run, best_dist


def _inputs():
    rng = random.Random(2026)
    words = [bytes(rng.choices(b"abcdefgh", k=rng.randrange(1, 6))) for _ in range(40)]
    text = b"".join(rng.choice(words) for _ in range(700))
    runs = b"".join(bytes([rng.randrange(4)]) * rng.randrange(1, 400) for _ in range(30))
    mixed = text[:900] + rng.randbytes(300) + runs[:900] + text[200:1100]
    return {"text": text[:2600], "runs": runs[:3000], "mixed": mixed, "period3": b"xyz" * 700, "one": b"q", "two": b"qq", "zeros": bytes(1500)}


@pytest.mark.parametrize("name", ["text", "runs", "mixed", "period3", "one", "two", "zeros"])
@pytest.mark.parametrize("lookahead,history", [(1 << 16, 1 << 15), (700, 1 << 15), (512, 200)])
def test_every_token_is_the_reference_rule_choice(oracle, name, lookahead, history):
    data = _inputs()[name]
    presets = [(oracle.LITERAL_STATIC, 1, 0, 1), (oracle.LITERAL_DYNAMIC, 1, 0, 2), (oracle.RLE_STATIC, 1, 1, 1), (oracle.RLE_DYNAMIC, 1, 1, 2),
               (oracle.FULL_STATIC, 1, 32768, 1), (oracle.FULL_DYNAMIC, 1, 32768, 2)]
    for strat, min_dist, max_dist, btype in presets:
        comp = oracle.deflate(data, (strat,), lookahead=lookahead, history=history)
        blocks, out = tokens_per_block(comp)
        assert out == data
        assert len(blocks) == max(1, -(-len(data) // lookahead))            # one block per `lookahead` bytes, the last one final
        pos = 0
        for bi, (bt, toks) in enumerate(blocks):
            assert bt == btype
            start, end = bi * lookahead, min(len(data), (bi + 1) * lookahead)
            assert pos == start
            hist0 = start - min(history, start)                             # the block sees at most `history` bytes in front of it
            for t in toks:
                run, dist = greedy(data, pos, end, pos - hist0, min_dist, max_dist)
                if run < 3:
                    assert t == data[pos], (name, strat, pos, t)
                    pos += 1
                else:
                    assert t == (run, dist), (name, strat, pos, t, (run, dist))
                    pos += run
            assert pos == end


@pytest.mark.parametrize("name", ["text", "runs", "mixed", "period3", "one", "two", "zeros", "empty", "bytes256"])
@pytest.mark.parametrize("lookahead,history", [(1 << 16, 1 << 15), (700, 1 << 15), (512, 200)])
def test_oracle_bytes_equal_the_python_restatement(oracle, name, lookahead, history):
    """Byte equality of the two restatements of comp/Lz77Huffman.java + DeflaterOutputStream.java (tests/ref_model.py vs
    oracle/oracle_deflate.c) for all six presets: search, symbolisation, histogram fix-ups, package-merge with the
    reference's tie order, code-length run-length coding, header, canonical codes, bit order, block framing."""
    import ref_model
    inputs = dict(_inputs(), empty=b"", bytes256=bytes(range(256)) * 3)
    data = inputs[name]
    for preset in ref_model.PRESETS:
        mine = ref_model.deflate(data, preset, lookahead, history)
        theirs = oracle.deflate(data, (getattr(oracle, preset),), lookahead=lookahead, history=history)
        assert mine == theirs, (name, preset, mine.hex()[:60], theirs.hex()[:60])


@pytest.mark.parametrize("lookahead,history", [(1 << 16, 1 << 15), (700, 1 << 15), (512, 200)])
def test_uncompressed_and_multistrategy_bytes_equal_the_python_restatement(oracle, lookahead, history):
    """comp/Uncompressed.java:19-48 (cost per start bit, blocks of at most 65535 bytes) and comp/MultiStrategy.java:31-57
    (cheapest substrategy for the bit position the block starts at) in both restatements: equal bytes."""
    import ref_model
    rng = random.Random(31)
    inputs = dict(_inputs(), empty=b"", noise=rng.randbytes(2500), noise_then_text=rng.randbytes(1500) + _inputs()["text"][:1500],
                  long_noise=rng.randbytes(66000 if lookahead > 65535 else 3000))
    combos = [("UNCOMPRESSED",), ("UNCOMPRESSED", "FULL_STATIC", "FULL_DYNAMIC"), ("RLE_DYNAMIC", "UNCOMPRESSED"),
              ("LITERAL_STATIC", "LITERAL_DYNAMIC", "UNCOMPRESSED")]
    for name, data in inputs.items():
        if len(data) > 4000 and lookahead > 65535:
            use = [c for c in combos if all(x in ("UNCOMPRESSED", "RLE_DYNAMIC", "LITERAL_STATIC", "LITERAL_DYNAMIC") for x in c)]
        else:
            use = combos
        for combo in use:
            mine = ref_model.deflate(data, list(combo), lookahead, history)
            theirs = oracle.deflate(data, tuple(getattr(oracle, x) for x in combo), lookahead=lookahead, history=history)
            assert mine == theirs, (name, combo, len(mine), len(theirs), mine.hex()[:60], theirs.hex()[:60])


@pytest.mark.parametrize("min_block", [100, 400])
def test_binary_split_bytes_equal_the_python_restatement(oracle, min_block):
    """comp/BinarySplit.java:30-98 in both restatements (recursive halving while a split is cheaper, kept per start bit
    position, costs summed from position 0 as the reference does): equal bytes, on data whose statistics change."""
    import ref_model
    rng = random.Random(41)
    inp = _inputs()
    datas = {"mixed": inp["mixed"][600:2600], "noise_then_text": rng.randbytes(900) + inp["text"][:900],
             "runs_noise_runs": inp["runs"][:700] + rng.randbytes(500) + inp["runs"][1200:1900]}
    n_split = 0
    for name, data in datas.items():
        for combo in (("FULL_DYNAMIC",), ("RLE_DYNAMIC",), ("UNCOMPRESSED", "FULL_STATIC", "FULL_DYNAMIC")):
            for lookahead in (1 << 16, 1100):
                mine = ref_model.deflate(data, list(combo), lookahead, 1 << 15, split_min_block_len=min_block)
                theirs, blocks = oracle.deflate_split(data, tuple(getattr(oracle, x) for x in combo), min_block_len=min_block,
                                                      lookahead=lookahead)
                assert mine == theirs, (name, combo, lookahead, len(mine), len(theirs))
                n_split += blocks > -(-len(data) // lookahead)
    assert n_split > 0                                     # (the splitter did cut somewhere)
