"""A second, independent restatement of the reference ENCODER in plain Python, for small inputs only.

TEST INFRASTRUCTURE.  The reference holds no golden compressed bytes and no JVM exists here, so the C
oracle's encoder cannot be pinned against the real thing.  What can be done is to state the same Java twice, in two
languages and two shapes, and demand equal bytes: this file restates comp/Lz77Huffman.java rule by rule in its simplest
form (brute-force search, Python's stable sort standing in for Collections.sort, lists of symbols standing in for the node
objects), where oracle/oracle_deflate.c uses hash chains, rank arithmetic and bit buffers.
Paths relative to /root/reference/src/io/nayuki/deflate/.
"""

CL_ORDER = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]      # Lz77Huffman.java:367-368


class BitOut:
    """BitOutputStream as DeflaterOutputStream.java:141-171 implements it: LSB-first, zero padding at the end."""

    def __init__(self):
        self.bits = []

    def write(self, value, n):
        for i in range(n):
            self.bits.append((value >> i) & 1)

    def tobytes(self):
        b = self.bits + [0] * (-len(self.bits) % 8)
        return bytes(sum(b[i + k] << k for k in range(8)) for i in range(0, len(b), 8))


def code_lengths(hist, max_len):
    """calcHuffmanCodeLengths, Lz77Huffman.java:309-335: package-merge; a tie in the sort keeps the packages of the
    previous round in front of the leaves, and the leaves in symbol order (the sort is stable)."""
    leaves = [(f, (s,)) for s, f in enumerate(hist) if f > 0]
    nodes = []
    for _ in range(max_len):
        nodes = sorted(nodes + leaves, key=lambda n: n[0])
        nodes = [(nodes[j][0] + nodes[j + 1][0], nodes[j][1] + nodes[j + 1][1]) for j in range(0, len(nodes) - 1, 2)]
    lens = [0] * len(hist)
    for n in nodes[:len(leaves) - 1]:
        for s in n[1]:
            lens[s] += 1
    return lens


def codes(lens, max_len):
    """codeLengthsToCodes, :371-391 -> [(code, nbits)], the code to be written MSB first."""
    out, nxt = [None] * len(lens), 0
    for n in range(1, max_len + 1):
        nxt <<= 1
        for s, ln in enumerate(lens):
            if ln == n:
                assert nxt >> n == 0
                out[s] = (nxt, n)
                nxt += 1
    assert nxt == 1 << max_len
    return out


def put_code(out, pair):
    code, n = pair
    for i in range(n - 1, -1, -1):
        out.write((code >> i) & 1, 1)


def length_symbol(run):                                                          # :93-107
    r = run - 3
    if run < 11:
        return r + 257, 0, 0
    if run == 258:
        return 285, 0, 0
    ne = r.bit_length() - 3
    return (ne << 2) + (r >> ne) + 257, ne, r & ((1 << ne) - 1)


def distance_symbol(dist):                                                       # :112-124
    d = dist - 1
    if dist < 5:
        return d, 0, 0
    ne = d.bit_length() - 2
    return (ne << 1) + (d >> ne), ne, d & ((1 << ne) - 1)


def search(data, index, end, avail, min_run, max_run, min_dist, max_dist):
    """:68-86 at one position."""
    best_run = best_dist = 0
    dist = min_dist
    while dist <= min(max_dist, avail) and best_run < max_run:
        run = 0
        while run < max_run and index + run < end and data[index + run] == data[index + run - dist]:
            run += 1
        if run > best_run or (run == best_run and dist < best_dist):
            best_run, best_dist = run, dist
        dist += 1
    return best_run, best_dist


_PARSED = {}


def block(out, data, start, end, hist_len, preset, is_final):
    """Decision.compressTo, :61-288.  preset = (dynamic, min_run, max_run, min_dist, max_dist)."""
    dynamic, min_run, max_run, min_dist, max_dist = preset
    key = (data, start, end, hist_len, preset[1:])
    if key not in _PARSED:                       # (the same span is compressed to count its bits, then again to emit them)
        toks, lit_hist, dist_hist = [], [0] * 286, [0] * 30
        i = start
        while i < end:
            run, dist = search(data, i, end, i - (start - hist_len), min_run, max_run, min_dist, max_dist)
            if run == 0 or run < min_run:
                toks.append((data[i],))
                lit_hist[data[i]] += 1
                i += 1
            else:
                ls, ds = length_symbol(run), distance_symbol(dist)
                toks.append((ls, ds))
                lit_hist[ls[0]] += 1
                dist_hist[ds[0]] += 1
                i += run
        lit_hist[256] += 1
        if len(_PARSED) > 4096:
            _PARSED.clear()
        _PARSED[key] = (toks, lit_hist, dist_hist)
    toks, lit_hist, dist_hist = _PARSED[key]
    lit_hist, dist_hist = list(lit_hist), list(dist_hist)
    out.write(1 if is_final else 0, 1)
    out.write(2 if dynamic else 1, 2)
    if not dynamic:                                                              # :394-410
        lit_code = codes([8] * 144 + [9] * 112 + [7] * 24 + [8] * 8, 9)
        dist_code = codes([5] * 32, 5)
    else:
        if end == start:                                                         # :147-148
            lit_hist[0] += 1
        n = len(lit_hist)
        while n > 257 and lit_hist[n - 1] == 0:
            n -= 1
        lit_len = code_lengths(lit_hist[:n], 15)
        if sum(1 for x in dist_hist if x > 0) == 1:                              # :157-172
            k = next(k for k, x in enumerate(dist_hist) if x > 0)
            if len(dist_hist) - k > 1:
                dist_hist[k + 1] = 1
            else:
                dist_hist[k - 1] = 1
        n = len(dist_hist)
        while n > 1 and dist_hist[n - 1] == 0:
            n -= 1
        dist_hist = dist_hist[:n]
        dist_len = [0] if (len(dist_hist) == 1 and dist_hist[0] == 0) else code_lengths(dist_hist, 15)
        lens = lit_len + dist_len
        syms, k = [], 0                                                          # :187-223 greedy run-length coding
        while k < len(lens):
            v = lens[k]
            if v == 0:
                r = 1
                while r < 138 and k + r < len(lens) and lens[k + r] == 0:
                    r += 1
                if r < 3:
                    syms.append((0,)); k += 1
                elif r < 11:
                    syms.append((17, r - 3, 3)); k += r
                else:
                    syms.append((18, r - 11, 7)); k += r
                continue
            if k > 0:
                r = 0
                while r < 6 and k + r < len(lens) and lens[k + r] == lens[k - 1]:
                    r += 1
                if r >= 3:
                    syms.append((16, r - 3, 2)); k += r
                    continue
            syms.append((v,)); k += 1
        cl_hist = [0] * 19
        for s in syms:
            cl_hist[s[0]] += 1
        cl_len = code_lengths(cl_hist, 7)
        reordered = [cl_len[o] for o in CL_ORDER]
        ncl = 19
        while ncl > 4 and reordered[ncl - 1] == 0:
            ncl -= 1
        out.write(len(lit_len) - 257, 5)
        out.write(len(dist_len) - 1, 5)
        out.write(ncl - 4, 4)
        for v in reordered[:ncl]:
            out.write(v, 3)
        cl_code = codes(cl_len, 7)
        for s in syms:
            put_code(out, cl_code[s[0]])
            if len(s) == 3:
                out.write(s[1], s[2])
        lit_code = codes(lit_len, 15)
        dist_code = None if dist_len == [0] else codes(dist_len, 15)
    for t in toks:                                                               # :267-285
        if len(t) == 1:
            put_code(out, lit_code[t[0]])
        else:
            (ls, lne, lex), (ds, dne, dex) = t
            put_code(out, lit_code[ls])
            out.write(lex, lne)
            put_code(out, dist_code[ds])
            out.write(dex, dne)
    put_code(out, lit_code[256])


PRESETS = {                                                                      # Lz77Huffman.java:298-306
    "LITERAL_STATIC": (False, 0, 0, 0, 0), "LITERAL_DYNAMIC": (True, 0, 0, 0, 0),
    "RLE_STATIC": (False, 3, 258, 1, 1), "RLE_DYNAMIC": (True, 3, 258, 1, 1),
    "FULL_STATIC": (False, 3, 258, 1, 32768), "FULL_DYNAMIC": (True, 3, 258, 1, 32768),
}


def stored(out, data, start, end, is_final):
    """comp/Uncompressed.java:33-46: stored blocks of at most 65535 bytes, the last one carries the final flag."""
    i = start
    while True:
        n = min(end - i, 65535)
        out.write(1 if (is_final and n == end - i) else 0, 1)
        out.write(0, 2)
        out.write(0, (8 - len(out.bits) % 8) % 8)
        out.write(n, 16)
        out.write(n ^ 0xFFFF, 16)
        for x in data[i:i + n]:
            out.write(x, 8)
        i += n
        if i >= end:
            break


def stored_bit_lengths(n):
    """comp/Uncompressed.java:21-25: the cost for each of the eight bit positions the block may start at."""
    blocks = max(-(-n // 65535), 1)
    return [n * 8 + blocks * 40 + ((13 - i) % 8 - 5) for i in range(8)]


class Lz77Decision:
    """What Lz77Huffman.decide returns (:42-58): the bit length, counted by compressing, is the same for all eight start
    positions."""

    def __init__(self, data, off, hist_len, n, preset):
        self.args = (data, off + hist_len, off + hist_len + n, hist_len, preset)
        tmp = BitOut()
        block(tmp, *self.args, False)
        self.bit_lengths = [len(tmp.bits)] * 8

    def compress_to(self, out, is_final):
        block(out, *self.args, is_final)


class StoredDecision:
    def __init__(self, data, off, hist_len, n):
        self.args = (data, off + hist_len, off + hist_len + n)
        self.bit_lengths = stored_bit_lengths(n)

    def compress_to(self, out, is_final):
        stored(out, *self.args, is_final)


class Chosen:
    """A decision made of one list of decisions per start bit position (MultiStrategy: one each; BinarySplit: one or two)."""

    def __init__(self, bit_lengths, per_position):
        self.bit_lengths, self.per_position = bit_lengths, per_position

    def compress_to(self, out, is_final):
        decs = self.per_position[len(out.bits) % 8]
        for k, d in enumerate(decs):
            d.compress_to(out, is_final and k == len(decs) - 1)


def strategy(names):
    """names -> decide(data, off, hist_len, n): one preset, "UNCOMPRESSED", or several = comp/MultiStrategy.java:31-57
    (per start bit position the cheapest substrategy, the earlier one on a tie)."""
    def one(name):
        if name == "UNCOMPRESSED":
            return lambda data, off, h, n: StoredDecision(data, off, h, n)
        return lambda data, off, h, n: Lz77Decision(data, off, h, n, PRESETS[name])
    subs = [one(x) for x in names]
    if len(subs) == 1:
        return subs[0]

    def decide(data, off, h, n):
        best, chosen = [None] * 8, [None] * 8
        for st in subs:
            d = st(data, off, h, n)
            for i in range(8):
                if best[i] is None or d.bit_lengths[i] < best[i]:
                    best[i], chosen[i] = d.bit_lengths[i], [d]
        return Chosen(best, chosen)
    return decide


def binary_split(sub, min_block_len):
    """comp/BinarySplit.java:30-98.  As there, the cost of the two halves is summed from bit position 0 whatever the
    position it is compared for (:52-56, :63-67)."""
    def split_cost(decs):
        total = 0
        for d in decs:
            total += d.bit_lengths[total % 8]
        return total

    def refine(data, off, h, n, cur):
        lengths, per_position = list(cur.bit_lengths), [[cur]] * 8
        first = (n + 1) // 2
        second = n - first
        if min(first, second) > min_block_len:
            halves = [sub(data, off, h, first), sub(data, off, h + first, second)]
            if any(split_cost(halves) < lengths[i] for i in range(8)):
                halves = [refine(data, off, h, first, halves[0]), refine(data, off, h + first, second, halves[1])]
            cost = split_cost(halves)
            for i in range(8):
                if cost < lengths[i]:
                    lengths[i], per_position[i] = cost, halves
        return Chosen(lengths, per_position)

    return lambda data, off, h, n: refine(data, off, h, n, sub(data, off, h, n))


def deflate(data, preset, lookahead=1 << 16, history=1 << 15, split_min_block_len=None):
    """new DeflaterOutputStream(out, lookahead, history, strategy).write(data).close() -- DeflaterOutputStream.java:76-137:
    a block per `lookahead` bytes, each seeing at most `history` bytes in front of it, the last one final.
    preset: a name from PRESETS, "UNCOMPRESSED", or a list of those = new MultiStrategy(...);
    split_min_block_len = n wraps that strategy in new BinarySplit(strategy, n)."""
    decide = strategy([preset] if isinstance(preset, str) else list(preset))
    if split_min_block_len is not None:
        decide = binary_split(decide, split_min_block_len)
    out = BitOut()
    n_blocks = max(1, -(-len(data) // lookahead))
    for b in range(n_blocks):
        start, end = b * lookahead, min(len(data), (b + 1) * lookahead)
        hist = min(history, start)
        decide(data, start - hist, hist, end - start).compress_to(out, b == n_blocks - 1)
    return out.tobytes()
