/*
 * oracle.h -- CPU restatement of the reference DEFLATE codec (TEST INFRASTRUCTURE ONLY).
 *
 * This library is the parity oracle for the B200 kernels.  It restates, in plain C, the
 * algorithms of nayuki/DEFLATE-library-Java for the hot path named in BASELINE.json:
 *   decoder  : src/io/nayuki/deflate/decomp/Open.java
 *   encoder  : src/io/nayuki/deflate/comp/{Lz77Huffman,Uncompressed,MultiStrategy}.java
 *              framed as src/io/nayuki/deflate/DeflaterOutputStream.java does
 *   container: src/io/nayuki/deflate/{GzipMetadata,GzipOutputStream,GzipInputStream}.java
 *   checksum : java.util.zip.CRC32 (JDK, not reference source; IEEE 802.3 reflected CRC-32)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product library (libb2deflate.so) never links, loads or calls anything in here.
 *
 * Parity pinning: the decoder is pinned by the reference's 39 deterministic test vectors
 * (tests/golden/inflate_vectors.json, extracted from InflaterInputStreamTest.java:24-510) and
 * cross-checked against system zlib.  The encoder has NO golden bytes in the reference
 * (DeflaterOutputStreamTest.java only round-trips) and no JVM exists in this image, so encoder
 * byte-level parity is UNPINNED by the reference; what pins it here: round trips through oracle_inflate
 * + zlib, the cross-check rows of SURVEY.md Appendix F, and byte equality with a second, independent
 * restatement of the same Java in plain Python (tests/ref_model.py, tests/test_oracle_tokens.py).
 */
#ifndef B2D_ORACLE_H
#define B2D_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes: 0 = OK, otherwise 1 + DataFormatException.Reason.ordinal()
 * (DataFormatException.java:61-83).  Negative = conditions that are not reference Reasons. */
enum {
	ORC_OK = 0,
	ORC_UNEXPECTED_END_OF_STREAM = 1,
	ORC_RESERVED_BLOCK_TYPE = 2,
	ORC_UNCOMPRESSED_BLOCK_LENGTH_MISMATCH = 3,
	ORC_HUFFMAN_CODE_UNDER_FULL = 4,
	ORC_HUFFMAN_CODE_OVER_FULL = 5,
	ORC_NO_PREVIOUS_CODE_LENGTH_TO_COPY = 6,
	ORC_CODE_LENGTH_CODE_OVER_FULL = 7,
	ORC_END_OF_BLOCK_CODE_ZERO_LENGTH = 8,
	ORC_RESERVED_LENGTH_SYMBOL = 9,
	ORC_RESERVED_DISTANCE_SYMBOL = 10,
	ORC_LENGTH_ENCOUNTERED_WITH_EMPTY_DISTANCE_CODE = 11,
	ORC_COPY_FROM_BEFORE_DICTIONARY_START = 12,
	ORC_HEADER_CHECKSUM_MISMATCH = 13,
	ORC_UNSUPPORTED_COMPRESSION_METHOD = 14,
	ORC_DECOMPRESSED_CHECKSUM_MISMATCH = 15,
	ORC_DECOMPRESSED_SIZE_MISMATCH = 16,
	ORC_GZIP_INVALID_MAGIC_NUMBER = 17,
	ORC_GZIP_RESERVED_FLAGS_SET = 18,
	ORC_GZIP_UNSUPPORTED_OPERATING_SYSTEM = 19,
	ORC_OUTPUT_OVERFLOW = -1,   /* caller's output capacity exhausted (no reference analogue) */
	ORC_BAD_ARGUMENT = -2
};

/* ---- decoder (Open.java) ---- */

/* Decodes one raw DEFLATE stream.  *out_len = bytes produced (on failure: bytes produced before the
 * failing symbol, i.e. what byte-at-a-time read() calls would have delivered before the exception).
 * *in_consumed = ceil(bits consumed / 8), the end-exactly position (Open.java:113-124); only
 * meaningful when the status is ORC_OK. */
int oracle_inflate(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap,
                   size_t *out_len, size_t *in_consumed);

/* Same, decoding bit-by-bit through the code tree only (Open.java:634-646 slow path), no LUT.
 * Used to cross-check the fast path. */
int oracle_inflate_slow(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap,
                        size_t *out_len, size_t *in_consumed);

/* ---- encoder (Lz77Huffman.java, Uncompressed.java, MultiStrategy.java, DeflaterOutputStream.java) ---- */

enum {
	ORC_STRAT_LITERAL_STATIC = 0,   /* Lz77Huffman.java:298 */
	ORC_STRAT_LITERAL_DYNAMIC = 1,  /* :299 */
	ORC_STRAT_RLE_STATIC = 2,       /* :301 */
	ORC_STRAT_RLE_DYNAMIC = 3,      /* :302  (the DeflaterOutputStream default, DeflaterOutputStream.java:51) */
	ORC_STRAT_FULL_STATIC = 4,      /* :304 */
	ORC_STRAT_FULL_DYNAMIC = 5,     /* :305 */
	ORC_STRAT_UNCOMPRESSED = 6      /* Uncompressed.java */
};

/* Compresses in[0,n) the way `new DeflaterOutputStream(out, lookahead, history, strategy)` followed by
 * write(in) and finish() would: one block per `lookahead` bytes, <=`history` bytes of look-behind,
 * final (possibly empty) block on finish, zero padding to a byte.  If n_strategies > 1 the strategies
 * are combined as `new MultiStrategy(strategies...)`.  search: 0 = exhaustive hash chains (bit-identical
 * to the brute-force scan of Lz77Huffman.java:71-84), 1 = the literal brute-force scan.
 * Returns the number of bytes written, or (size_t)-1 if cap is too small / bad arguments. */
size_t oracle_deflate(const uint8_t *in, size_t n, const int *strategies, int n_strategies,
                      int lookahead, int history, int search, uint8_t *out, size_t cap);

/* The same with `new BinarySplit(substrategy, min_block_len)` as the stream's strategy (comp/BinarySplit.java:30-98),
 * substrategy = the single strategy or the MultiStrategy of `strategies`: every lookahead block is recursively cut in
 * halves while a cut makes it smaller, halves no shorter than min_block_len + 1.  *n_blocks (optional) = number of
 * blocks chosen, judged at bit position 0.  The reference never uses BinarySplit by default and has no test for it:
 * pinned by round trips, by never being larger than the unsplit stream, and by byte equality with the Python
 * restatement of tests/ref_model.py. */
size_t oracle_deflate_split(const uint8_t *in, size_t n, const int *strategies, int n_strategies,
                            int lookahead, int history, int search, int min_block_len,
                            uint8_t *out, size_t cap, size_t *n_blocks);

/* Upper bound for oracle_deflate's output. */
size_t oracle_deflate_bound(size_t n, int lookahead);

/* calcHuffmanCodeLengths (Lz77Huffman.java:309-335): package-merge; hist[n] -> lens[n]. */
void oracle_package_merge(const int *hist, int n, int max_len, uint8_t *lens);

/* ---- checksum (java.util.zip.CRC32) ---- */
uint32_t oracle_crc32(uint32_t crc, const uint8_t *p, size_t n);

/* ---- gzip container as src/gzip.java writes it (GzipMetadata.java:164-212, GzipOutputStream.java:62-70) ---- */

/* Header with FNAME + FHCRC, OS=UNIX, XFL=0, optional mtime (0 = absent).  Returns header length. */
size_t oracle_gzip_header(uint8_t *out, size_t cap, const char *file_name, uint32_t mtime,
                          const uint8_t *extra, size_t extra_len);

/* Parses a gzip member header the way GzipMetadata.read does (GzipMetadata.java:73-146).
 * Returns status; *header_len = bytes consumed. */
int oracle_gzip_parse_header(const uint8_t *in, size_t n, size_t *header_len);

/* Whole-member decode as GzipInputStream (GzipInputStream.java:38-90): header, inflate end-exactly,
 * CRC-32 then ISIZE check. */
int oracle_gunzip(const uint8_t *in, size_t n, uint8_t *out, size_t out_cap, size_t *out_len,
                  size_t *in_consumed);

/* ---- zlib container (ZlibMetadata.java, ZlibInputStream.java, ZlibOutputStream.java; java.util.zip.Adler32) ---- */
uint32_t oracle_adler32(uint32_t adler, const uint8_t *p, size_t n);
int oracle_unzlib(const uint8_t *in, size_t n, uint8_t *out, size_t out_cap, size_t *out_len, size_t *in_consumed);

#ifdef __cplusplus
}
#endif
#endif
