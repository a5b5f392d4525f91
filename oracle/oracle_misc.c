/*
 * oracle_misc.c -- CRC-32 and gzip container (TEST INFRASTRUCTURE ONLY, see oracle.h).
 *
 * CRC-32: the reference delegates to java.util.zip.CRC32 (JDK; GzipOutputStream.java:25,57,67,
 *   GzipInputStream.java:32,72,83) -- IEEE 802.3, reflected polynomial 0xEDB88320, init/xorout 0xFFFFFFFF.
 *   Restated bit-serially through a 256-entry table; pinned by crc32("123456789") == 0xCBF43926 and
 *   against Python's zlib.crc32 in tests/test_oracle_misc.py.
 * gzip: GzipMetadata.java:73-146 (read), :164-212 (write); GzipOutputStream.java:62-70 (trailer);
 *   GzipInputStream.java:66-90 (trailer check order: CRC first, then ISIZE mod 2^32);
 *   src/gzip.java:52-62 (the metadata the CLI writes: FNAME + FHCRC, OS = UNIX, XFL = 0).
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

/* ---------- CRC-32 ---------- */
static uint32_t crc_table[256];
static int crc_ready = 0;

static void crc_init(void) {
	if (crc_ready) return;
	for (uint32_t i = 0; i < 256; i++) {
		uint32_t c = i;
		for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
		crc_table[i] = c;
	}
	__sync_synchronize();
	crc_ready = 1;
}

uint32_t oracle_crc32(uint32_t crc, const uint8_t *p, size_t n) {
	crc_init();
	uint32_t c = ~crc;
	for (size_t i = 0; i < n; i++) c = crc_table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
	return ~c;
}

/* ---------- gzip container ---------- */
size_t oracle_gzip_header(uint8_t *out, size_t cap, const char *file_name, uint32_t mtime,
                          const uint8_t *extra, size_t extra_len) {
	size_t name_len = file_name ? strlen(file_name) + 1 : 0;
	size_t need = 10 + (extra ? 2 + extra_len : 0) + name_len + 2;
	if (need > cap || extra_len > 0xFFFF) return (size_t)-1;
	size_t p = 0;
	out[p++] = 0x1F; out[p++] = 0x8B;                 /* GzipMetadata.java:169 */
	out[p++] = 8;                                     /* DEFLATE :171-174 */
	out[p++] = (uint8_t)(0x02 | (extra ? 0x04 : 0) | (file_name ? 0x08 : 0));   /* FHCRC | FEXTRA | FNAME :176-187 */
	out[p++] = (uint8_t)mtime; out[p++] = (uint8_t)(mtime >> 8);
	out[p++] = (uint8_t)(mtime >> 16); out[p++] = (uint8_t)(mtime >> 24);       /* :189 */
	out[p++] = 0;                                     /* XFL :191 */
	out[p++] = 3;                                     /* OS = UNIX.ordinal() :193-196 */
	if (extra) {
		out[p++] = (uint8_t)extra_len; out[p++] = (uint8_t)(extra_len >> 8);    /* :198-202 */
		memcpy(out + p, extra, extra_len); p += extra_len;
	}
	if (file_name) { memcpy(out + p, file_name, name_len); p += name_len; }     /* :204-205 */
	uint32_t crc = oracle_crc32(0, out, p);
	out[p++] = (uint8_t)crc; out[p++] = (uint8_t)(crc >> 8);                    /* :210-211 low 16 bits, LE */
	return p;
}

int oracle_gzip_parse_header(const uint8_t *in, size_t n, size_t *header_len) {
	size_t p = 0;
#define NEED(k) do { if (p + (size_t)(k) > n) return ORC_UNEXPECTED_END_OF_STREAM; } while (0)
	NEED(2);
	if (in[0] != 0x1F || in[1] != 0x8B) return ORC_GZIP_INVALID_MAGIC_NUMBER;   /* :81-82 */
	p = 2;
	NEED(1);
	if (in[p++] != 8) return ORC_UNSUPPORTED_COMPRESSION_METHOD;                /* :84-86 */
	NEED(1);
	int flags = in[p++];
	if (flags & 0xE0) return ORC_GZIP_RESERVED_FLAGS_SET;                       /* :93-95 */
	NEED(4); p += 4;                                                            /* mtime */
	NEED(1); p += 1;                                                            /* XFL */
	NEED(1);
	int os = in[p++];
	if (!(os < 14 || os == 0xFF)) return ORC_GZIP_UNSUPPORTED_OPERATING_SYSTEM; /* :104-111 */
	if (flags & 0x04) {                                                         /* FEXTRA :116-122 */
		NEED(2);
		size_t len = (size_t)in[p] | (size_t)in[p + 1] << 8;
		p += 2;
		NEED(len); p += len;
	}
	if (flags & 0x08) { do { NEED(1); } while (in[p++] != 0); }                 /* FNAME :124-126 */
	if (flags & 0x10) { do { NEED(1); } while (in[p++] != 0); }                 /* FCOMMENT :128-130 */
	if (flags & 0x02) {                                                         /* FHCRC :132-138 */
		uint32_t expect = oracle_crc32(0, in, p) & 0xFFFF;
		NEED(2);
		uint32_t actual = (uint32_t)in[p] | (uint32_t)in[p + 1] << 8;
		p += 2;
		if (actual != expect) return ORC_HEADER_CHECKSUM_MISMATCH;
	}
#undef NEED
	*header_len = p;
	return 0;
}

int oracle_gunzip(const uint8_t *in, size_t n, uint8_t *out, size_t out_cap, size_t *out_len,
                  size_t *in_consumed) {
	size_t hl = 0, ol = 0, used = 0;
	if (out_len) *out_len = 0;
	int st = oracle_gzip_parse_header(in, n, &hl);
	if (st) return st;
	st = oracle_inflate(in + hl, n - hl, out, out_cap, &ol, &used);
	if (out_len) *out_len = ol;
	if (st) return st;
	size_t p = hl + used;
	if (p + 8 > n) return ORC_UNEXPECTED_END_OF_STREAM;                          /* GzipInputStream.java:77-81 */
	uint32_t crc = (uint32_t)in[p] | (uint32_t)in[p + 1] << 8 | (uint32_t)in[p + 2] << 16 | (uint32_t)in[p + 3] << 24;
	uint32_t isz = (uint32_t)in[p + 4] | (uint32_t)in[p + 5] << 8 | (uint32_t)in[p + 6] << 16 | (uint32_t)in[p + 7] << 24;
	if (in_consumed) *in_consumed = p + 8;
	if (oracle_crc32(0, out, ol) != crc) return ORC_DECOMPRESSED_CHECKSUM_MISMATCH;   /* :82-83 */
	if ((uint32_t)ol != isz) return ORC_DECOMPRESSED_SIZE_MISMATCH;                   /* :85-86 */
	return 0;
}


/* ---- zlib container (ZlibMetadata.java:47-104, ZlibInputStream.java:64-83, ZlibOutputStream.java:60-67) ---- */

/* java.util.zip.Adler32 (JDK, not reference source): RFC 1950 section 9.  Fresh value = 1. */
uint32_t oracle_adler32(uint32_t adler, const uint8_t *p, size_t n) {
	uint32_t s1 = adler & 0xFFFF, s2 = adler >> 16;
	for (size_t i = 0; i < n; i++) {
		s1 += p[i];
		if (s1 >= 65521) s1 -= 65521;
		s2 += s1;
		if (s2 >= 65521) s2 -= 65521;
	}
	return s2 << 16 | s1;
}

/* ZlibInputStream read to the end: header checks in ZlibMetadata.read's order, inflate end-exactly, big-endian
 * Adler-32 trailer. */
int oracle_unzlib(const uint8_t *in, size_t n, uint8_t *out, size_t out_cap, size_t *out_len, size_t *in_consumed) {
	if (out_len) *out_len = 0;
	if (n < 2) return ORC_UNEXPECTED_END_OF_STREAM;                              /* :50-52 */
	int cmf = in[0], flg = in[1];
	if ((cmf << 8 | flg) % 31 != 0) return ORC_HEADER_CHECKSUM_MISMATCH;         /* :53-54 */
	int method = cmf & 0xF;
	if (method != 8 && method != 15) return ORC_UNSUPPORTED_COMPRESSION_METHOD;  /* :56-61 */
	size_t hl = 2;
	if ((flg >> 5) & 1) {                                                        /* :65-75 preset dictionary id */
		if (n < 6) return ORC_UNEXPECTED_END_OF_STREAM;
		hl = 6;
	}
	if (method == 8 && (cmf >> 4) > 7) return ORC_BAD_ARGUMENT;                  /* IllegalArgumentException from the record (:24-25) */
	size_t ol = 0, used = 0;
	int st = oracle_inflate(in + hl, n - hl, out, out_cap, &ol, &used);
	if (out_len) *out_len = ol;
	if (st) return st;
	size_t p = hl + used;
	if (p + 4 > n) return ORC_UNEXPECTED_END_OF_STREAM;
	uint32_t expect = (uint32_t)in[p] << 24 | (uint32_t)in[p + 1] << 16 | (uint32_t)in[p + 2] << 8 | (uint32_t)in[p + 3];
	if (in_consumed) *in_consumed = p + 4;
	if (oracle_adler32(1, out, ol) != expect) return ORC_DECOMPRESSED_CHECKSUM_MISMATCH;
	return 0;
}
