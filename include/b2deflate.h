/*
 * b2deflate.h -- C ABI of libb2deflate.so, the B200 (sm_100a) DEFLATE codec.
 *
 * This is the drop-in boundary for the hot path of nayuki/DEFLATE-library-Java (paths below are
 * relative to the reference's src/io/nayuki/deflate/).  Each entry point names the reference
 * interface it replaces; INTEGRATION.md shows the Panama FFM (java.lang.foreign) binding a
 * maintainer of the reference would add.  Plain pointers and sizes only -- no CUDA or torch types.
 *
 * Conventions
 *   - One process drives one GPU (b2d_init(device)) or several GPUs of the box (b2d_init_devices): the host
 *     entry points then partition the independent units (gzip members, chunks) into contiguous ranges, one per
 *     GPU, and join the results.  Calls are blocking and serialised per GPU by an internal mutex,
 *     so they may arrive from any thread (the reference classes are not thread-safe and spawn no threads; FFM
 *     downcalls may come from any JVM thread).
 *   - Host entry points take HOST pointers and include the host<->device copies.  The *_dev entry
 *     points take DEVICE pointers (e.g. torch tensor data_ptr()) and run on the given CUDA stream
 *     (a cudaStream_t passed as void*; NULL = the CUDA legacy default stream, which is what
 *     torch.cuda.current_stream().cuda_stream is unless the caller switched streams) without synchronising it.
 *   - There is no CPU fallback: every call fails with B2D_ERR_NO_DEVICE if no sm_100 GPU is usable.
 *   - Format errors are per-member status codes: 0 = OK, otherwise 1 + Reason.ordinal() of the
 *     reference's DataFormatException.Reason (DataFormatException.java:61-83).  The Java wrapper
 *     re-throws `new DataFormatException(Reason.values()[status - 1], b2d_strerror(status))`.
 */
#ifndef B2DEFLATE_H
#define B2DEFLATE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ---- */
enum {
	B2D_OK = 0,
	/* 1 + DataFormatException.Reason.ordinal()  (DataFormatException.java:61-83) */
	B2D_UNEXPECTED_END_OF_STREAM = 1,
	B2D_RESERVED_BLOCK_TYPE = 2,
	B2D_UNCOMPRESSED_BLOCK_LENGTH_MISMATCH = 3,
	B2D_HUFFMAN_CODE_UNDER_FULL = 4,
	B2D_HUFFMAN_CODE_OVER_FULL = 5,
	B2D_NO_PREVIOUS_CODE_LENGTH_TO_COPY = 6,
	B2D_CODE_LENGTH_CODE_OVER_FULL = 7,
	B2D_END_OF_BLOCK_CODE_ZERO_LENGTH = 8,
	B2D_RESERVED_LENGTH_SYMBOL = 9,
	B2D_RESERVED_DISTANCE_SYMBOL = 10,
	B2D_LENGTH_ENCOUNTERED_WITH_EMPTY_DISTANCE_CODE = 11,
	B2D_COPY_FROM_BEFORE_DICTIONARY_START = 12,
	B2D_HEADER_CHECKSUM_MISMATCH = 13,
	B2D_UNSUPPORTED_COMPRESSION_METHOD = 14,
	B2D_DECOMPRESSED_CHECKSUM_MISMATCH = 15,
	B2D_DECOMPRESSED_SIZE_MISMATCH = 16,
	B2D_GZIP_INVALID_MAGIC_NUMBER = 17,
	B2D_GZIP_RESERVED_FLAGS_SET = 18,
	B2D_GZIP_UNSUPPORTED_OPERATING_SYSTEM = 19,
	/* not reference Reasons (the reference's streams are unbounded / have no device) */
	B2D_ERR_OUTPUT_OVERFLOW = -1,   /* a member needs more output than its out_off[] slot holds */
	B2D_ERR_BAD_ARGUMENT = -2,      /* maps to IllegalArgumentException / NullPointerException */
	B2D_ERR_NO_DEVICE = -3,         /* no usable sm_100 GPU / b2d_init not called */
	B2D_ERR_CUDA = -4,              /* CUDA runtime failure; see b2d_last_error() */
	B2D_ERR_OUT_OF_MEMORY = -5
};

/* ---- lifetime ---- */

/* Binds the process to GPU `device` (CUDA ordinal), creates the stream and scratch pools.
 * Idempotent for the same device.  Reference objects "only use memory and no OS resources"
 * (InflaterInputStream.java:22-23), so nothing below requires per-stream teardown. */
int b2d_init(int device);
/* Binds the process to n GPUs (CUDA ordinals, distinct); devices[0] is "GPU 0", the one single-device work runs on and
 * the gather target.  n <= 0: every GPU of the box.  The host entry points (b2d_inflate_batch, b2d_gunzip_batch,
 * b2d_deflate_chunks[_indexed], b2d_inflate_chunks) then shard their units over the GPUs, one host thread per GPU,
 * with the same results byte for byte as on one GPU; the *_dev entry points run on the GPU their output pointer
 * lives on.  Compressed payloads are joined at the scanned offsets: straight into the caller's host buffer over every
 * GPU's own PCIe link, or -- environment B2D_MULTI_GATHER=peer -- with one cudaMemcpyPeerAsync per GPU into GPU 0's
 * memory (NVLink) and one copy from there.  The reference has no counterpart (DeflaterOutputStream.java:119-137 and
 * Open.java:83-110 are sequential). */
int b2d_init_devices(const int *devices, int n);
int b2d_device_count(void);                  /* GPUs bound by b2d_init / b2d_init_devices */
void b2d_shutdown(void);
uint64_t b2d_kernel_launches(void);          /* kernels this library has launched so far (process-wide counter) */
const char *b2d_strerror(int status);        /* message strings of Open.java / GzipInputStream.java */
const char *b2d_last_error(void);            /* last CUDA error text for B2D_ERR_CUDA */
int b2d_device_sm_count(void);

/* Pinned staging memory the Java side wraps as a MemorySegment (ownership: caller frees). */
void *b2d_alloc_pinned(size_t bytes);
void b2d_free_pinned(void *p);

/* ---- decompression: replaces decomp/Open.java (Open.read :83-110) behind
 *      InflaterInputStream.read (InflaterInputStream.java:147-164) ---- */

#define B2D_INFLATE_CRC32        1u   /* also compute CRC-32 of every member's output (GzipInputStream.java:72) */
#define B2D_INFLATE_ADLER32      4u   /* the crc32[] slot receives Adler-32 instead (ZlibInputStream.java:69); wins over CRC32 */
#define B2D_INFLATE_CHUNK_INDEXED 2u  /* unit = chunk of a b2d_deflate_chunks stream: ends at BFINAL *or* exactly
                                         at the end of its byte range on a block boundary */

/* Decodes n independent raw-DEFLATE members.  Member i occupies in[in_off[i], in_off[i+1]) and is
 * written to out[out_off[i], out_off[i+1]) (the slot size is its capacity).
 *   out_len[i]     bytes produced (on failure: bytes produced before the failing symbol)
 *   in_consumed[i] ceil(bits consumed / 8): the end-exactly position of Open.finish (Open.java:113-124)
 *   crc32[i]       CRC-32 of the member's output if B2D_INFLATE_CRC32 (may be NULL otherwise)
 *   status[i]      0 or 1 + Reason.ordinal(); a bad member does not disturb the others
 * If `out` is page-locked memory (b2d_alloc_pinned, cudaHostAlloc, cudaHostRegister) the decoding kernel delivers the
 * bytes to it directly while it decodes and only out[out_off[i], out_off[i] + out_len[i]) is written; otherwise the
 * whole slots are copied back after the kernels.
 * Returns B2D_OK if the batch ran (inspect status[]), or a negative B2D_ERR_*. */
int b2d_inflate_batch(const uint8_t *in, const uint64_t *in_off, uint32_t n,
                      uint8_t *out, const uint64_t *out_off,
                      uint64_t *out_len, uint64_t *in_consumed, uint32_t *crc32, int32_t *status,
                      uint32_t flags);

/* Same with device pointers.  `in` must be readable up to the next 4-byte boundary past in_off[n]. */
int b2d_inflate_batch_dev(const uint8_t *d_in, const uint64_t *d_in_off, uint32_t n,
                          uint8_t *d_out, const uint64_t *d_out_off,
                          uint64_t *d_out_len, uint64_t *d_in_consumed, uint32_t *d_crc32, int32_t *d_status,
                          uint32_t flags, void *stream);

/* ---- one stream of any origin, decoded in parallel.  What `new InflaterInputStream(in)` read to the end does for a
 *      stream nobody indexed (a file made by system gzip, zlib, or the reference's own DeflaterOutputStream): one member on
 *      one warp is the sequential decoder of Open.java:83-110 at a few tens of MB/s, so the stream is decoded
 *      speculatively instead -- block starts found by a header-plausibility scan (the checks of Open.java:336-431 are
 *      the filter), one warp per found start, back-references into the unknown 32 KiB in front of each unit carried as
 *      markers and resolved along the chain of units (csrc/inflate.cu).  Whenever that cannot vouch for the result --
 *      an invalid stream, long stretches without a dynamic block, a unit outgrowing its buffer -- the stream is decoded
 *      again sequentially, so bytes, out_len, in_consumed and status are always the sequential decoder's.
 *      *parallel (may be NULL) receives 1 if the parallel decode produced the result. ---- */
int b2d_inflate_stream(const uint8_t *in, uint64_t in_len, uint8_t *out, uint64_t out_cap, uint64_t *out_len,
                       uint64_t *in_consumed, uint32_t *crc32, int32_t *status, uint32_t flags, int32_t *parallel);
/* Device pointers (d_in 4-byte aligned, readable to the next 16-byte boundary past in_len).  d_result: five uint64
 * {out_len, in_consumed, status, 0, units}; status != 0: decode with b2d_inflate_batch_dev for the exact outcome. */
int b2d_inflate_stream_dev(const uint8_t *d_in, uint64_t in_len, uint8_t *d_out, uint64_t out_cap, uint64_t *d_result, void *stream);

/* ---- gzip members: GzipInputStream.java:38-90 over a batch ---- */

/* ISIZE (mod 2^32) from each member's trailer: the output capacity a caller needs for members below 4 GiB. */
int b2d_gzip_isize(const uint8_t *in, const uint64_t *in_off, uint32_t n, uint64_t *isize);

/* Decodes n independent gzip members (host pointers): header checks of GzipMetadata.read (GzipMetadata.java:73-146) in
 * the reference's order, body + CRC-32 on the GPU, then the trailer checks of GzipInputStream.java:73-88 (CRC, then
 * ISIZE mod 2^32).  status[i] = 0 or 1 + Reason.ordinal() (incl. the container reasons 13..19); in_consumed[i] counts
 * header + body + 8 trailer bytes; only the first member of each range is read, trailing bytes are ignored like
 * GzipInputStream.java:66-74 does. */
int b2d_gunzip_batch(const uint8_t *in, const uint64_t *in_off, uint32_t n, uint8_t *out, const uint64_t *out_off,
                     uint64_t *out_len, uint64_t *in_consumed, int32_t *status);

/* ---- compression: replaces comp/Lz77Huffman.java (decide/compressTo :42-288), comp/Uncompressed.java,
 *      comp/MultiStrategy.java behind DeflaterOutputStream.writeBuffer (DeflaterOutputStream.java:119-137) ---- */

enum {
	B2D_MODE_AUTO = 0,      /* per block the cheapest of stored / fixed / dynamic (MultiStrategy.java:35-44 rule) */
	B2D_MODE_STORED = 1,    /* comp/Uncompressed.java */
	B2D_MODE_FIXED = 2,     /* Lz77Huffman *_STATIC  (Lz77Huffman.java:298,301,304) */
	B2D_MODE_DYNAMIC = 3    /* Lz77Huffman *_DYNAMIC (Lz77Huffman.java:299,302,305) */
};

/* Match search of the Lz77Huffman record (Lz77Huffman.java:20-25): which back-references are tried. */
enum {
	B2D_SEARCH_DEFAULT = 0, /* hash chains, depth-limited, lazy parse (ratio >= FULL_DYNAMIC on text/mixed data) */
	B2D_SEARCH_LITERAL = 1, /* no matches            = LITERAL_*  (0,0,0,0) */
	B2D_SEARCH_RLE = 2,     /* distance 1 only, greedy = RLE_*    (3,258,1,1), the DeflaterOutputStream default */
	B2D_SEARCH_FULL = 3     /* exhaustive chains, 3-byte minimum, greedy = FULL_* (3,258,1,32768): finds exactly
	                           the matches of the brute-force scan Lz77Huffman.java:71-84 (slow; parity/ratio bar) */
};

/* Stream framing. */
enum {
	B2D_FRAMING_CHUNKED = 0,   /* independent chunks, each closed by an empty stored block (sync-flush marker) */
	B2D_FRAMING_REFERENCE = 1  /* exactly what DeflaterOutputStream emits (DeflaterOutputStream.java:119-137): history
	                              carried across blocks, no markers, BFINAL on the last (possibly empty) block when
	                              is_last, zero padding to a byte.  chunk_bytes is ignored; the input is ONE unit, so
	                              only block-level parallelism remains.  With search RLE/FULL/LITERAL and lazy = 0 the
	                              output is byte-identical to the reference strategy of the same name. */
};

/* Checksum of the uncompressed data computed alongside compression. */
enum {
	B2D_CHECKSUM_CRC32 = 0,    /* gzip: java.util.zip.CRC32 at GzipOutputStream.java:57; running value starts at 0 */
	B2D_CHECKSUM_ADLER32 = 1   /* zlib: java.util.zip.Adler32 at ZlibOutputStream.java:56; running value starts at 1 */
};

typedef struct b2d_deflate_opts {
	uint32_t chunk_bytes;   /* independent unit, history reset at its start; 0 = 1 MiB */
	uint32_t block_bytes;   /* one DEFLATE block per this many input bytes (dataLookaheadLimit,
	                           DeflaterOutputStream.java:50-52); 0 = 64 KiB; must divide chunk_bytes */
	int32_t mode;           /* B2D_MODE_* */
	int32_t search;         /* B2D_SEARCH_* */
	int32_t chain_depth;    /* candidates examined per position; 0 = default (4) */
	int32_t lazy;           /* -1 = default (on), 0 = greedy, 1 = lazy */
	int32_t is_last;        /* 1: the stream ends after this call (final block emitted) */
	int32_t framing;        /* B2D_FRAMING_* */
	int32_t checksum;       /* B2D_CHECKSUM_*: which checksum of the input crc32_inout / d_chunk_crc32 carry */
	uint32_t split_min_bytes; /* 0 = one block per block_bytes.  Otherwise adaptive block splitting, the counterpart of
	                           comp/BinarySplit.java:30-98: the input is parsed in pieces of this many bytes (a power of
	                           two >= 4096 dividing block_bytes, at most 16 pieces per block_bytes) and every
	                           block_bytes span becomes the cheapest partition of its binary tree of pieces -- a piece,
	                           a pair, ... or the whole span as one block -- priced with the real codes of every node.
	                           The block index then has one entry per block_bytes span (its first block). */
} b2d_deflate_opts;

/* Worst-case output size of b2d_deflate_chunks for in_len bytes (any mode, any framing). */
uint64_t b2d_deflate_bound(uint64_t in_len, uint32_t chunk_bytes);

/* Compresses in[0,in_len) as ceil(in_len / chunk_bytes) independent chunks.  The output is raw DEFLATE,
 * byte-aligned: every chunk ends with an empty stored block (bits 0|00, zero pad, 00 00 FF FF -- the
 * Z_SYNC_FLUSH marker), whose BFINAL bit is set only on the last chunk when opts->is_last.  The
 * concatenation of all calls' outputs is ONE valid DEFLATE stream for the reference's sequential
 * decoder (Open.java:83-110, stored path :232-241) and for zlib.
 *   crc32_inout    running CRC-32 of the uncompressed data (GzipOutputStream.java:57), updated; may be NULL
 *   chunk_out_len  per-chunk compressed byte counts (the chunk index), ceil(in_len/chunk_bytes) entries; may be NULL
 * Returns bytes written, or a negative B2D_ERR_*. */
int64_t b2d_deflate_chunks(const uint8_t *in, uint64_t in_len, const b2d_deflate_opts *opts,
                           uint8_t *out, uint64_t out_cap, uint32_t *crc32_inout, uint64_t *chunk_out_len);

/* Same with device pointers (d_in 16-byte aligned, d_out 4-byte aligned; d_out need not be cleared, the call
 * clears it).  *d_out_len_total (device, 8 bytes) receives the byte count; d_chunk_out_len / d_chunk_crc32
 * receive the per-chunk compressed sizes and per-chunk CRC-32s of the input (either may be NULL). */
int b2d_deflate_chunks_dev(const uint8_t *d_in, uint64_t in_len, const b2d_deflate_opts *opts,
                           uint8_t *d_out, uint64_t out_cap, uint64_t *d_out_len_total,
                           uint64_t *d_chunk_out_len, uint32_t *d_chunk_crc32, void *stream);

/* ---- block-indexed streams: the same stream as b2d_deflate_chunks, plus the bit offset of every DEFLATE block inside
 *      its chunk (one uint32 per block_bytes of input).  With it the decoder runs one warp per BLOCK (16 x the units of
 *      the chunk index): Huffman decode of all blocks in parallel, back-references replayed per chunk afterwards.  There
 *      is no reference counterpart (DeflaterOutputStream.java:119-137 / Open.java:83-110 are sequential); the stream itself
 *      stays one valid DEFLATE stream for the reference's decoder. ---- */
int64_t b2d_deflate_chunks_indexed(const uint8_t *in, uint64_t in_len, const b2d_deflate_opts *opts, uint8_t *out,
                                   uint64_t out_cap, uint32_t *crc32_inout, uint64_t *chunk_out_len,
                                   uint32_t *block_bits /* ceil(in_len / block_bytes) entries */);
int b2d_deflate_chunks_indexed_dev(const uint8_t *d_in, uint64_t in_len, const b2d_deflate_opts *opts, uint8_t *d_out,
                                   uint64_t out_cap, uint64_t *d_out_len_total, uint64_t *d_chunk_out_len,
                                   uint32_t *d_chunk_crc32, uint32_t *d_block_bits, void *stream);
/* Chunk c occupies chunk_in_len[c] bytes of `in` (back to back) and decodes to out[c * chunk_bytes, ...); out_total is the
 * exact decompressed size.  chunk_status[c] = 0 or 1 + Reason.ordinal(); the host form re-decodes a failing chunk
 * sequentially so that status and bytes are the sequential decoder's.  flags: B2D_INFLATE_CRC32 / _ADLER32 fill
 * chunk_crc32[] with per-chunk checksums (fold them with b2d_crc32_combine / b2d_adler32_combine). */
int b2d_inflate_chunks(const uint8_t *in, const uint64_t *chunk_in_len, uint32_t n_chunks, const uint32_t *block_bits,
                       uint32_t chunk_bytes, uint32_t block_bytes, uint8_t *out, uint64_t out_total,
                       uint32_t *chunk_crc32, int32_t *chunk_status, uint32_t flags);
int b2d_inflate_chunks_dev(const uint8_t *d_in, const uint64_t *d_chunk_in_off /* n_chunks + 1 */, uint32_t n_chunks,
                           const uint32_t *d_block_bits, uint32_t chunk_bytes, uint32_t block_bytes, uint64_t out_total,
                           uint8_t *d_out, uint32_t *d_chunk_crc32, int32_t *d_chunk_status, uint32_t flags, void *stream);

/* ---- CRC-32: replaces java.util.zip.CRC32 at GzipOutputStream.java:25,57 / GzipInputStream.java:32,72 ---- */
int b2d_crc32_update(const uint8_t *data, uint64_t len, uint32_t *crc_inout);    /* host pointer, computed on the GPU; B2D_OK or B2D_ERR_* */
uint32_t b2d_crc32(uint32_t crc, const uint8_t *data, uint64_t len);            /* convenience: on failure the value is returned
                                                                                    unchanged and b2d_last_error() says why */
int b2d_crc32_dev(const uint8_t *d_data, uint64_t len, uint32_t *d_crc_out, void *stream);   /* crc of d_data[0,len), init 0 */
uint32_t b2d_crc32_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b);     /* crc(A||B) from crc(A), crc(B), |B| */

/* ---- Adler-32: replaces java.util.zip.Adler32 at ZlibOutputStream.java:25,56,65 / ZlibInputStream.java:30,69,78
 *      (SURVEY.md 8f, row N4).  A fresh checksum starts at 1. ---- */
int b2d_adler32_update(const uint8_t *data, uint64_t len, uint32_t *adler_inout);
uint32_t b2d_adler32(uint32_t adler, const uint8_t *data, uint64_t len);         /* convenience form, like b2d_crc32 */
uint32_t b2d_adler32_combine(uint32_t adler_a, uint32_t adler_b, uint64_t len_b);

/* ---- synthetic corpora (host; SURVEY.md Appendix D -- the reference ships no data) ---- */
void b2d_corpus_random(uint64_t seed, uint8_t *out, size_t n);
void b2d_corpus_text(uint64_t seed, uint8_t *out, size_t n);
void b2d_corpus_mixed(uint64_t seed, uint8_t *out, size_t n);

#ifdef __cplusplus
}
#endif
#endif
