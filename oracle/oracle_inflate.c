/*
 * oracle_inflate.c -- CPU restatement of the reference decoder (TEST INFRASTRUCTURE ONLY, see oracle.h).
 *
 * Follows src/io/nayuki/deflate/decomp/Open.java of nayuki/DEFLATE-library-Java:
 *   block loop                     Open.java:83-110
 *   bit reader (LSB-first)         Open.java:137-170
 *   stored block                   Open.java:227-306
 *   dynamic header                 Open.java:336-431
 *   Huffman data loop              Open.java:438-620  (LUT + residual tree walk :483-493)
 *   slow-path symbol decode        Open.java:634-676
 *   codeLengthsToCodeTree          Open.java:705-756
 *   codeTreeToCodeTable            Open.java:771-789
 *   tables                         Open.java:794-886
 * The validation ORDER is the reference's (first failing check wins).  The stream abstraction is
 * replaced by a flat in-memory buffer, so "end of stream" means "a bit is needed past in_len".
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

#define CODE_TABLE_BITS 9                       /* Open.java:803 */
#define CODE_TABLE_MASK ((1 << CODE_TABLE_BITS) - 1)
#define DICTIONARY_LENGTH 32768                 /* Open.java:201 */

typedef struct {
	const uint8_t *in;
	size_t in_len, in_pos;      /* in_pos = next byte to load into the bit buffer */
	uint64_t bitbuf;
	int bitlen;                 /* 0..64, always whole bytes + leftover of the current byte */
	uint8_t *out;
	size_t out_cap, out_pos;
	int use_table;
	int err;
} St;

/* Open.java:137-170.  Returns -1 on end of stream (err set). */
static int read_bits(St *s, int n) {
	while (s->bitlen < n) {
		if (s->in_pos >= s->in_len) {
			s->err = ORC_UNEXPECTED_END_OF_STREAM;   /* Open.java:186-187 */
			return -1;
		}
		s->bitbuf |= (uint64_t)s->in[s->in_pos++] << s->bitlen;
		s->bitlen += 8;
	}
	int r = (int)(s->bitbuf & ((1u << n) - 1));
	s->bitbuf >>= n;
	s->bitlen -= n;
	return r;
}

/* Best-effort refill used by the data loop (Open.java:451-475): never fails, just loads what exists. */
static void refill(St *s) {
	while (s->bitlen <= 56 && s->in_pos < s->in_len) {
		s->bitbuf |= (uint64_t)s->in[s->in_pos++] << s->bitlen;
		s->bitlen += 8;
	}
}

/* ---- code trees (Open.java:705-756) ---- */

static int cmp_short(const void *a, const void *b) {
	return (int)*(const int16_t *)a - (int)*(const int16_t *)b;
}

/* tree must hold 2*(n-1) shorts at most (n<=288 -> 574).  Returns 0 or an ORC_ status; *tree_len set. */
static int code_lengths_to_tree(const uint8_t *lens, int n, int16_t *tree, int *tree_len) {
	int16_t pairs[320];
	for (int i = 0; i < n; i++)
		pairs[i] = (int16_t)(lens[i] << 11 | i);      /* :713-717 */
	qsort(pairs, (size_t)n, sizeof(int16_t), cmp_short);

	int idx = 0;
	while (idx < n && (pairs[idx] >> 11) == 0)        /* :720-722 skip unused symbols */
		idx++;
	int num_codes = n - idx;
	if (num_codes < 2)
		return ORC_HUFFMAN_CODE_UNDER_FULL;           /* :725-726 */

	int result_len = (num_codes - 1) * 2;
	int next = 0, end = 2, cur_len = 1;
	for (; idx < n; idx++) {
		int pair = pairs[idx];
		for (int code_len = pair >> 11; cur_len < code_len; cur_len++) {
			for (int e = end; next < e; next++) {     /* :736-743 double every open slot */
				if (end >= result_len)
					return ORC_HUFFMAN_CODE_UNDER_FULL;
				tree[next] = (int16_t)end;
				end += 2;
			}
		}
		if (next >= end)
			return ORC_HUFFMAN_CODE_OVER_FULL;        /* :745-746 */
		int symbol = pair & 0x7FF;
		tree[next] = (int16_t)~symbol;
		next++;
	}
	if (next < end)
		return ORC_HUFFMAN_CODE_UNDER_FULL;           /* :753-754 */
	*tree_len = result_len;
	return 0;
}

/* Open.java:771-789 */
static void code_tree_to_table(const int16_t *tree, int16_t *table) {
	for (int i = 0; i < (1 << CODE_TABLE_BITS); i++) {
		int node = 0, consumed = 0;
		do {
			node = tree[node + ((i >> consumed) & 1)];
			consumed++;
		} while (node >= 0 && consumed < CODE_TABLE_BITS);
		table[i] = (int16_t)(node << 4 | consumed);
	}
}

/* ---- constant tables (Open.java:843-886) ---- */
static int16_t RUN_LENGTH_TABLE[29];
static int32_t DISTANCE_TABLE[30];
static int16_t FIXED_LL_TREE[574], FIXED_LL_TABLE[512];
static int16_t FIXED_D_TREE[62], FIXED_D_TABLE[512];
static int tables_ready = 0;

static void init_tables(void) {
	if (tables_ready)
		return;
	for (int i = 0; i < 29; i++) {
		int sym = i + 257, run, eb;
		if (sym <= 264) { eb = 0; run = sym - 254; }
		else if (sym <= 284) { eb = (sym - 261) / 4; run = (((sym - 1) % 4 + 4) << eb) + 3; }
		else { eb = 0; run = 258; }
		RUN_LENGTH_TABLE[i] = (int16_t)(run << 3 | eb);
	}
	for (int sym = 0; sym < 30; sym++) {
		int dist, eb;
		if (sym <= 3) { eb = 0; dist = sym + 1; }
		else { eb = sym / 2 - 1; dist = ((sym % 2 + 2) << eb) + 1; }
		DISTANCE_TABLE[sym] = dist << 4 | eb;
	}
	uint8_t ll[288], dl[32];
	memset(ll, 8, 144); memset(ll + 144, 9, 112); memset(ll + 256, 7, 24); memset(ll + 280, 8, 8);   /* :812-818 */
	memset(dl, 5, 32);
	int tl;
	code_lengths_to_tree(ll, 288, FIXED_LL_TREE, &tl);
	code_lengths_to_tree(dl, 32, FIXED_D_TREE, &tl);
	code_tree_to_table(FIXED_LL_TREE, FIXED_LL_TABLE);
	code_tree_to_table(FIXED_D_TREE, FIXED_D_TABLE);
	__sync_synchronize();
	tables_ready = 1;
}

/* Open.java:634-646: bit-by-bit tree walk.  Returns symbol or -1 (err set). */
static int decode_symbol_slow(St *s, const int16_t *tree) {
	int node = 0;
	while (node >= 0) {
		int b = read_bits(s, 1);
		if (b < 0)
			return -1;
		node = tree[node + b];
	}
	return ~node;
}

/* LUT + residual walk on buffered bits (Open.java:483-493), made safe for short streams:
 * a lookup that would consume more bits than exist is an end-of-stream. */
static int decode_symbol(St *s, const int16_t *tree, const int16_t *table) {
	if (!s->use_table)
		return decode_symbol_slow(s, tree);
	if (s->bitlen < 15 + 13)
		refill(s);
	int temp = table[(int)s->bitbuf & CODE_TABLE_MASK];
	int consumed = temp & 0xF;
	if (consumed > s->bitlen) {
		s->err = ORC_UNEXPECTED_END_OF_STREAM;
		return -1;
	}
	s->bitbuf >>= consumed;
	s->bitlen -= consumed;
	int node = temp >> 4;
	while (node >= 0) {
		if (s->bitlen == 0) {
			s->err = ORC_UNEXPECTED_END_OF_STREAM;
			return -1;
		}
		node = tree[node + ((int)s->bitbuf & 1)];
		s->bitbuf >>= 1;
		s->bitlen--;
	}
	return ~node;
}

/* ---- stored block (Open.java:227-306) ---- */
static int stored_block(St *s) {
	if (read_bits(s, s->bitlen % 8) < 0)               /* :234 align to byte */
		return s->err;
	int len = read_bits(s, 16);
	if (len < 0)
		return s->err;
	int nlen = read_bits(s, 16);
	if (nlen < 0)
		return s->err;
	if (len != (nlen ^ 0xFFFF))
		return s->err = ORC_UNCOMPRESSED_BLOCK_LENGTH_MISMATCH;   /* :239-240 */
	/* first unpack saved bits (:259-261), then straight from the input (:263-285) */
	while (len > 0) {
		int b;
		if (s->bitlen >= 8) {
			b = (int)(s->bitbuf & 0xFF);
			s->bitbuf >>= 8;
			s->bitlen -= 8;
		} else {
			if (s->in_pos >= s->in_len)
				return s->err = ORC_UNEXPECTED_END_OF_STREAM;     /* :279-280 */
			b = s->in[s->in_pos++];
		}
		if (s->out_pos >= s->out_cap)
			return s->err = ORC_OUTPUT_OVERFLOW;
		s->out[s->out_pos++] = (uint8_t)b;
		len--;
	}
	return 0;
}

/* ---- Huffman block (Open.java:322-620) ---- */
static int huffman_block(St *s, int dynamic) {
	int16_t ll_tree_buf[574], ll_table_buf[512], d_tree_buf[62], d_table_buf[512];
	const int16_t *ll_tree, *ll_table, *d_tree, *d_table;
	if (!dynamic) {
		ll_tree = FIXED_LL_TREE; ll_table = FIXED_LL_TABLE;
		d_tree = FIXED_D_TREE; d_table = FIXED_D_TABLE;
	} else {
		static const int ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};  /* :794-795 */
		int hlit = read_bits(s, 5);  if (hlit < 0) return s->err;
		int hdist = read_bits(s, 5); if (hdist < 0) return s->err;
		int hclen = read_bits(s, 4); if (hclen < 0) return s->err;
		int num_ll = hlit + 257, num_d = hdist + 1, num_cl = hclen + 4;      /* :336-340 */
		uint8_t cl_len[19];
		memset(cl_len, 0, sizeof cl_len);
		for (int i = 0; i < num_cl; i++) {
			int v = read_bits(s, 3);
			if (v < 0) return s->err;
			cl_len[ORDER[i]] = (uint8_t)v;
		}
		int16_t cl_tree[36];
		int tl;
		int e = code_lengths_to_tree(cl_len, 19, cl_tree, &tl);              /* :344 */
		if (e) return s->err = e;

		uint8_t lens[288 + 32];
		int total = num_ll + num_d;
		int run_val = -1;
		for (int i = 0; i < total; ) {                                       /* :347-379 */
			int sym = decode_symbol_slow(s, cl_tree);
			if (sym < 0) return s->err;
			if (sym < 16) {
				run_val = sym;
				lens[i++] = (uint8_t)sym;
			} else {
				int run_len;
				if (sym == 16) {
					if (run_val == -1)
						return s->err = ORC_NO_PREVIOUS_CODE_LENGTH_TO_COPY;  /* :359-361 (before the extra bits) */
					int x = read_bits(s, 2); if (x < 0) return s->err;
					run_len = x + 3;
				} else if (sym == 17) {
					run_val = 0;
					int x = read_bits(s, 3); if (x < 0) return s->err;
					run_len = x + 3;
				} else {
					run_val = 0;
					int x = read_bits(s, 7); if (x < 0) return s->err;
					run_len = x + 11;
				}
				for (; run_len > 0; run_len--, i++) {
					if (i >= total)
						return s->err = ORC_CODE_LENGTH_CODE_OVER_FULL;       /* :374-375 */
					lens[i] = (uint8_t)run_val;
				}
			}
		}
		if (lens[256] == 0)
			return s->err = ORC_END_OF_BLOCK_CODE_ZERO_LENGTH;               /* :383-384 */
		e = code_lengths_to_tree(lens, num_ll, ll_tree_buf, &tl);            /* :385 */
		if (e) return s->err = e;
		code_tree_to_table(ll_tree_buf, ll_table_buf);
		ll_tree = ll_tree_buf; ll_table = ll_table_buf;

		uint8_t dlen[32];
		memset(dlen, 0, sizeof dlen);
		memcpy(dlen, lens + num_ll, (size_t)num_d);
		if (num_d == 1 && dlen[0] == 0) {                                    /* :398-401 */
			d_tree = NULL; d_table = NULL;
		} else {
			int one = 0, other = 0;
			for (int i = 0; i < num_d; i++) {
				if (dlen[i] == 1) one++;
				else if (dlen[i] > 1) other++;
			}
			int nd = num_d;
			if (one == 1 && other == 0) {                                    /* :421-425 */
				nd = 32;
				dlen[31] = 1;
			}
			e = code_lengths_to_tree(dlen, nd, d_tree_buf, &tl);             /* :426 */
			if (e) return s->err = e;
			code_tree_to_table(d_tree_buf, d_table_buf);
			d_tree = d_tree_buf; d_table = d_table_buf;
		}
	}

	for (;;) {                                                               /* :446-618 */
		int sym = decode_symbol(s, ll_tree, ll_table);
		if (sym < 0) return s->err;
		if (sym < 256) {
			if (s->out_pos >= s->out_cap)
				return s->err = ORC_OUTPUT_OVERFLOW;
			s->out[s->out_pos++] = (uint8_t)sym;
			continue;
		}
		if (sym == 256)
			return 0;
		if (sym - 257 >= 29)
			return s->err = ORC_RESERVED_LENGTH_SYMBOL;                      /* :513-517, :659 */
		int temp = RUN_LENGTH_TABLE[sym - 257];
		int x = read_bits(s, temp & 7);
		if (x < 0) return s->err;
		int run = (temp >> 3) + x;
		if (d_tree == NULL)
			return s->err = ORC_LENGTH_ENCOUNTERED_WITH_EMPTY_DISTANCE_CODE; /* :526-527, :578-579 */
		int dsym = decode_symbol(s, d_tree, d_table);
		if (dsym < 0) return s->err;
		if (dsym >= 30)
			return s->err = ORC_RESERVED_DISTANCE_SYMBOL;                    /* :546-551, :674 */
		int dt = DISTANCE_TABLE[dsym];
		x = read_bits(s, dt & 0xF);
		if (x < 0) return s->err;
		size_t dist = (size_t)((dt >> 4) + x);
		size_t dict_len = s->out_pos < DICTIONARY_LENGTH ? s->out_pos : DICTIONARY_LENGTH;
		if (dist > dict_len)
			return s->err = ORC_COPY_FROM_BEFORE_DICTIONARY_START;           /* :592-593 */
		if (s->out_pos + (size_t)run > s->out_cap) {
			/* deliver what fits, like a short caller buffer would (:604-616), then report overflow */
			while (s->out_pos < s->out_cap) { s->out[s->out_pos] = s->out[s->out_pos - dist]; s->out_pos++; }
			return s->err = ORC_OUTPUT_OVERFLOW;
		}
		uint8_t *dst = s->out + s->out_pos;
		const uint8_t *src = dst - dist;
		for (int i = 0; i < run; i++)                                        /* :596-603 byte-serial, overlap replicates */
			dst[i] = src[i];
		s->out_pos += (size_t)run;
	}
}

static int inflate_impl(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap,
                        size_t *out_len, size_t *in_consumed, int use_table) {
	init_tables();
	St s;
	memset(&s, 0, sizeof s);
	s.in = in; s.in_len = in_len;
	s.out = out; s.out_cap = out_cap;
	s.use_table = use_table;
	int last = 0;
	while (!last) {                                                          /* Open.java:83-110 */
		int b = read_bits(&s, 1);
		if (b < 0) break;
		last = b;
		int type = read_bits(&s, 2);
		if (type < 0) break;
		if (type == 0) { if (stored_block(&s)) break; }
		else if (type == 1) { if (huffman_block(&s, 0)) break; }
		else if (type == 2) { if (huffman_block(&s, 1)) break; }
		else { s.err = ORC_RESERVED_BLOCK_TYPE; break; }                     /* :96 */
	}
	if (out_len) *out_len = s.out_pos;
	if (in_consumed) *in_consumed = s.in_pos - (size_t)(s.bitlen / 8);       /* Open.java:113-124 */
	return s.err;
}

int oracle_inflate(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap,
                   size_t *out_len, size_t *in_consumed) {
	return inflate_impl(in, in_len, out, out_cap, out_len, in_consumed, 1);
}

int oracle_inflate_slow(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap,
                        size_t *out_len, size_t *in_consumed) {
	return inflate_impl(in, in_len, out, out_cap, out_len, in_consumed, 0);
}
