"""Development aid: the decoder's symbol loop on synthetic tables, one warp (needs a library built with -DB2D_LOOPBENCH:
tools/variants.py inflate.cu loopbench=-DB2D_LOOPBENCH; B2D_SO=<that library>).  Prints cycles per symbol."""
import ctypes, os
L = ctypes.CDLL(os.environ["B2D_SO"])
assert L.b2d_init(0) == 0
out = (ctypes.c_uint32 * 4)()
for mode, name in ((0, "literals (5 bits)"), (1, "pairs (7 + 5 bits)"), (2, "mixed")):
    assert L.b2d_loop_bench(mode, 20000, out) == 0
    cyc, nbytes, bits, events = out[0], out[1], out[2], out[3]
    # literal: 1 byte, 5 bits; pair: 5 bytes, 12 bits
    n_pair = (5 * nbytes - bits) / 13.0 if mode else 0
    n_lit = nbytes - 5 * n_pair
    print(f"{name:20s} {cyc} cycles, {events} loop entries, {n_lit:.0f} literals + {n_pair:.0f} pairs: {cyc / max(n_lit + n_pair, 1):.1f} cycles per symbol")
