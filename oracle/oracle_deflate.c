/*
 * oracle_deflate.c -- CPU restatement of the reference encoder (TEST INFRASTRUCTURE ONLY, see oracle.h).
 *
 * Follows nayuki/DEFLATE-library-Java:
 *   comp/Lz77Huffman.java        decide/compressTo :42-288, presets :298-305,
 *                                calcHuffmanCodeLengths (package-merge) :309-335,
 *                                codeLengthsToCodes :372-391, static codes :394-410
 *   comp/Uncompressed.java       :19-48
 *   comp/MultiStrategy.java      :31-57
 *   comp/BinarySplit.java        :30-98 (oracle_deflate_split)
 *   DeflaterOutputStream.java    framing :76-137, BitOut :141-171
 *
 * Encoder byte-level parity is UNPINNED (the reference has no golden compressed bytes and no JVM exists
 * here); this restatement is validated by round trips through oracle_inflate and zlib and by the
 * cross-check rows of SURVEY.md Appendix F, and by byte equality with tests/ref_model.py (the same Java restated
 * a second time, in plain Python, with none of this file's data structures).
 *
 * The match search is the reference's greedy "longest run, ties to the smallest distance" rule
 * (Lz77Huffman.java:68-84).  search=1 runs the literal brute-force scan; search=0 enumerates the same
 * candidates through exhaustive 3-byte hash chains, which yields bit-identical tokens: a distance can
 * only win with run >= 3 (shorter runs become literals, :85), and run >= 3 <=> equal 3-byte keys.
 * Unlike the reference, each block is compressed once (the reference runs compressTo twice, once on a
 * counting sink :46-52 and once for real, DeflaterOutputStream.java:123-124); the bytes are the same.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

/* ---------- bit sink (DeflaterOutputStream.java:141-171) ---------- */
typedef struct {
	uint8_t *out;
	size_t cap, pos;
	uint64_t bitbuf;
	int bitlen;
	int counting;       /* CountingBitOutputStream.java:19-31 */
	uint64_t count;
	int overflow;
} BitOut;

static void put_byte(BitOut *o, uint8_t b) {
	if (o->pos < o->cap) o->out[o->pos] = b; else o->overflow = 1;
	o->pos++;
}

static void write_bits(BitOut *o, uint32_t value, int n) {
	if (o->counting) { o->count += (uint64_t)n; return; }
	if (n > 64 - o->bitlen) {
		for (; o->bitlen >= 8; o->bitlen -= 8, o->bitbuf >>= 8)
			put_byte(o, (uint8_t)o->bitbuf);
	}
	o->bitbuf |= (uint64_t)value << o->bitlen;
	o->bitlen += n;
}

static int bit_position(const BitOut *o) {
	return o->counting ? (int)(o->count % 8) : o->bitlen % 8;
}

static void finish_bits(BitOut *o) {
	write_bits(o, 0, (8 - bit_position(o)) % 8);
	for (; o->bitlen >= 8; o->bitlen -= 8, o->bitbuf >>= 8)
		put_byte(o, (uint8_t)o->bitbuf);
}

/* ---------- package-merge (Lz77Huffman.java:309-335) ---------- */
typedef struct Node {
	int64_t freq;
	int sym;             /* >= 0 leaf, -1 internal */
	struct Node *a, *b;
} Node;

static void count_occ(const Node *nd, uint8_t *hist) {
	if (nd->sym >= 0) hist[nd->sym]++;
	else { count_occ(nd->a, hist); count_occ(nd->b, hist); }
}

/* stable merge sort of node pointers by frequency (Collections.sort is a stable merge sort) */
static void stable_sort(Node **v, Node **tmp, int n) {
	if (n < 2) return;
	int h = n / 2;
	stable_sort(v, tmp, h);
	stable_sort(v + h, tmp, n - h);
	int i = 0, j = h, k = 0;
	while (i < h && j < n) tmp[k++] = (v[j]->freq < v[i]->freq) ? v[j++] : v[i++];
	while (i < h) tmp[k++] = v[i++];
	while (j < n) tmp[k++] = v[j++];
	memcpy(v, tmp, (size_t)n * sizeof(Node *));
}

void oracle_package_merge(const int *hist, int n, int max_len, uint8_t *lens) {
	Node *leaves = (Node *)malloc((size_t)n * sizeof(Node));
	int nl = 0;
	for (int s = 0; s < n; s++)
		if (hist[s] > 0) { leaves[nl].freq = hist[s]; leaves[nl].sym = s; leaves[nl].a = leaves[nl].b = NULL; nl++; }
	memset(lens, 0, (size_t)n);
	/* every level holds at most (prev + nl) / 2 < nl packages */
	Node *arena = (Node *)malloc((size_t)(max_len * (nl + 1) + 1) * sizeof(Node));
	int arena_n = 0;
	Node **nodes = (Node **)malloc((size_t)(2 * nl + 2) * sizeof(Node *));
	Node **tmp = (Node **)malloc((size_t)(2 * nl + 2) * sizeof(Node *));
	int nn = 0;
	for (int i = 0; i < max_len; i++) {
		for (int k = 0; k < nl; k++) nodes[nn++] = &leaves[k];        /* nodes.addAll(leaves): packages first, then leaves */
		stable_sort(nodes, tmp, nn);
		int m = 0;
		for (int j = 0; j + 2 <= nn; j += 2) {
			Node *p = &arena[arena_n++];
			p->freq = nodes[j]->freq + nodes[j + 1]->freq;
			p->sym = -1; p->a = nodes[j]; p->b = nodes[j + 1];
			tmp[m++] = p;
		}
		memcpy(nodes, tmp, (size_t)m * sizeof(Node *));
		nn = m;
	}
	for (int i = 0; i < nl - 1; i++)                                   /* :331-334 */
		count_occ(nodes[i], lens);
	free(tmp); free(nodes); free(arena); free(leaves);
}

/* ---------- canonical codes (Lz77Huffman.java:372-391) ---------- */
static uint32_t reverse_bits(uint32_t v, int n) {
	uint32_t r = 0;
	for (int i = 0; i < n; i++) r |= ((v >> i) & 1u) << (n - 1 - i);
	return r;
}

/* result[sym] = reversed_code << 4 | len.  Returns 0, or -1 if over-/under-full (the reference throws). */
static int lengths_to_codes(const uint8_t *lens, int n, int max_len, uint32_t *result) {
	uint32_t next = 0;
	memset(result, 0, (size_t)n * sizeof(uint32_t));
	for (int len = 1; len <= max_len; len++) {
		next <<= 1;
		for (int s = 0; s < n; s++) {
			if (lens[s] != len) continue;
			if (next >> len) return -1;
			result[s] = reverse_bits(next, len) << 4 | (uint32_t)len;
			next++;
		}
	}
	return next == (1u << max_len) ? 0 : -1;
}

static uint32_t STATIC_LL_CODE[288], STATIC_D_CODE[32];
static int static_ready = 0;
static void init_static(void) {
	if (static_ready) return;
	uint8_t ll[288], dl[32];
	memset(ll, 8, 144); memset(ll + 144, 9, 112); memset(ll + 256, 7, 24); memset(ll + 280, 8, 8);
	memset(dl, 5, 32);
	lengths_to_codes(ll, 288, 9, STATIC_LL_CODE);                     /* :394-403 */
	lengths_to_codes(dl, 32, 5, STATIC_D_CODE);                       /* :405-410 */
	__sync_synchronize();
	static_ready = 1;
}

/* ---------- strategies ---------- */
typedef struct { int dynamic, min_run, max_run, min_dist, max_dist; } Lz77Params;

static const Lz77Params PRESETS[6] = {
	{0, 0, 0, 0, 0}, {1, 0, 0, 0, 0},          /* LITERAL_* :298-299 */
	{0, 3, 258, 1, 1}, {1, 3, 258, 1, 1},      /* RLE_*     :301-302 */
	{0, 3, 258, 1, 32768}, {1, 3, 258, 1, 32768}  /* FULL_*  :304-305 */
};

static int nlz32(uint32_t x) { return x == 0 ? 32 : __builtin_clz(x); }

/* exhaustive chains over the whole input: key = 3 bytes; prev links to the previous position with the same key */
typedef struct {
	const uint8_t *data;
	size_t n;
	int32_t *head;      /* 1<<24 entries, -1 = empty */
	int32_t *prev;      /* per position */
	size_t inserted;    /* positions [0, inserted) are in the chains */
} Chains;

static void chains_insert_upto(Chains *c, size_t upto) {   /* insert positions p < upto that have 3 bytes */
	for (; c->inserted < upto && c->inserted + 2 < c->n; c->inserted++) {
		size_t p = c->inserted;
		uint32_t key = (uint32_t)c->data[p] | (uint32_t)c->data[p + 1] << 8 | (uint32_t)c->data[p + 2] << 16;
		c->prev[p] = c->head[key];
		c->head[key] = (int32_t)p;
	}
	if (c->inserted < upto && c->inserted + 2 >= c->n) c->inserted = upto;
}

/* undo insertions back to `mark` (LIFO), so a pricing pass leaves the chains as it found them */
static void chains_rollback(Chains *c, size_t mark) {
	while (c->inserted > mark) {
		size_t p = --c->inserted;
		if (p + 2 >= c->n) continue;
		uint32_t key = (uint32_t)c->data[p] | (uint32_t)c->data[p + 1] << 8 | (uint32_t)c->data[p + 2] << 16;
		c->head[key] = c->prev[p];
	}
}

/* One Lz77Huffman block (Lz77Huffman.java:55-286).  b = whole input, [start,end) = this block,
 * off = start - historyLen. */
static void lz77_block(const Lz77Params *p, const uint8_t *b, size_t off, size_t start, size_t end,
                       Chains *ch, int brute, BitOut *out, int is_final) {
	size_t data_len = end - start;
	size_t cap = data_len * 4 / 3 + 4 + (data_len * 4 % 3 ? 1 : 0) + 1;
	uint16_t *toks = (uint16_t *)malloc(cap * 2 * sizeof(uint16_t) + 64);   /* generous */
	size_t nt = 0;
	int ll_hist[286], d_hist[30];
	memset(ll_hist, 0, sizeof ll_hist);
	memset(d_hist, 0, sizeof d_hist);

	size_t index = start;
	while (index < end) {
		int best_run = 0, best_dist = 0;
		size_t avail = index - off;
		int dist_end = (int)((size_t)p->max_dist < avail ? (size_t)p->max_dist : avail);
		int max_run = p->max_run;
		if (p->max_run > 0 && p->min_dist <= dist_end) {
			size_t lim = end - index;                  /* runs stop at the block end (:75) */
			int cap_run = (size_t)max_run < lim ? max_run : (int)lim;
			if (brute || ch == NULL) {
				for (int dist = p->min_dist; dist <= dist_end && best_run < max_run; dist++) {   /* :71-84 */
					int run = 0;
					const uint8_t *h = b + index - dist, *d = b + index;
					while (run < cap_run && d[run] == h[run]) run++;   /* direct overlapped compare == wrapped compare */
					if (run > best_run) { best_run = run; best_dist = dist; }
				}
			} else if (cap_run >= 3) {
				chains_insert_upto(ch, index);
				uint32_t key = (uint32_t)b[index] | (uint32_t)b[index + 1] << 8 | (uint32_t)b[index + 2] << 16;
				for (int32_t c = ch->head[key]; c >= 0; c = ch->prev[c]) {
					size_t dist = index - (size_t)c;
					if (dist > (size_t)dist_end) break;                /* chains are ordered by decreasing position */
					if (dist < (size_t)p->min_dist) continue;
					int run = 3;
					const uint8_t *h = b + c, *d = b + index;
					while (run < cap_run && d[run] == h[run]) run++;
					if (run > best_run) { best_run = run; best_dist = (int)dist; if (best_run >= max_run) break; }
				}
			}
		}
		if (best_run == 0 || best_run < p->min_run) {                  /* :85-90 */
			int sym = b[index];
			index++;
			toks[nt++] = (uint16_t)(sym << 4);
			ll_hist[sym]++;
		} else {
			{
				int r = best_run - 3, ne, sym, extra;                  /* :92-110 */
				if (best_run < 11) { ne = 0; sym = r + 257; extra = 0; }
				else if (best_run == 258) { ne = 0; sym = 285; extra = 0; }
				else { ne = 29 - nlz32((uint32_t)r); sym = (ne << 2) + (r >> ne) + 257; extra = r & ((1 << ne) - 1); }
				toks[nt++] = (uint16_t)(sym << 4 | ne);
				ll_hist[sym]++;
				toks[nt++] = (uint16_t)extra;
			}
			{
				int d = best_dist - 1, ne, sym, extra;                 /* :112-126 */
				if (best_dist < 5) { ne = 0; sym = d; extra = 0; }
				else { ne = 30 - nlz32((uint32_t)d); sym = (ne << 1) + (d >> ne); extra = d & ((1 << ne) - 1); }
				toks[nt++] = (uint16_t)(sym << 4 | ne);
				d_hist[sym]++;
				toks[nt++] = (uint16_t)extra;
			}
			index += (size_t)best_run;
		}
	}
	toks[nt++] = (uint16_t)(256 << 4);                                 /* :131-132 */
	ll_hist[256]++;

	write_bits(out, is_final ? 1 : 0, 1);
	write_bits(out, p->dynamic ? 2 : 1, 2);

	uint32_t ll_code_buf[288], d_code_buf[32];
	const uint32_t *ll_code, *d_code;
	if (!p->dynamic) {
		ll_code = STATIC_LL_CODE; d_code = STATIC_D_CODE;
	} else {
		int n_ll = 286, n_d = 30;
		if (data_len == 0) ll_hist[0]++;                               /* :146-147 */
		for (; n_ll > 257 && ll_hist[n_ll - 1] == 0; n_ll--);          /* :148-151 */
		uint8_t ll_len[286], d_len[30];
		oracle_package_merge(ll_hist, n_ll, 15, ll_len);               /* :153 */

		int used = 0;
		for (int i = 0; i < 30; i++) if (d_hist[i] > 0) used++;
		if (used == 1) {                                               /* :161-171 */
			for (int i = 0; i < 30; i++) {
				if (d_hist[i] > 0) {
					if (30 - i > 1) d_hist[i + 1] = 1; else d_hist[i - 1] = 1;
					break;
				}
			}
		}
		for (; n_d > 1 && d_hist[n_d - 1] == 0; n_d--);                /* :172-175 */
		int no_dist = (n_d == 1 && d_hist[0] == 0);
		if (no_dist) d_len[0] = 0;                                     /* :177-179 */
		else oracle_package_merge(d_hist, n_d, 15, d_len);             /* :181 */

		uint8_t lens[316];
		int total = n_ll + n_d;
		memcpy(lens, ll_len, (size_t)n_ll);
		memcpy(lens + n_ll, d_len, (size_t)n_d);

		int cl_syms[316], cl_extra[316], ncl = 0;
		for (int i = 0; i < total; ) {                                 /* :189-223 greedy RLE */
			int val = lens[i];
			if (val == 0) {
				int rl = 1;
				for (; rl < 138 && i + rl < total && lens[i + rl] == 0; rl++);
				if (rl < 3) { cl_syms[ncl] = 0; cl_extra[ncl++] = 0; i++; }
				else if (rl < 11) { cl_syms[ncl] = 17; cl_extra[ncl++] = rl - 3; i += rl; }
				else { cl_syms[ncl] = 18; cl_extra[ncl++] = rl - 11; i += rl; }
				continue;
			}
			if (i > 0) {
				int rl = 0;
				for (; rl < 6 && i + rl < total && lens[i + rl] == lens[i - 1]; rl++);
				if (rl >= 3) { cl_syms[ncl] = 16; cl_extra[ncl++] = rl - 3; i += rl; continue; }
			}
			cl_syms[ncl] = val; cl_extra[ncl++] = 0; i++;
		}
		int cl_hist[19];
		memset(cl_hist, 0, sizeof cl_hist);
		for (int i = 0; i < ncl; i++) cl_hist[cl_syms[i]]++;
		uint8_t cl_len[19];
		oracle_package_merge(cl_hist, 19, 7, cl_len);                  /* :228 */

		static const int ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
		int reordered[19], ncll = 19;
		for (int i = 0; i < 19; i++) reordered[i] = cl_len[ORDER[i]];
		for (; ncll > 4 && reordered[ncll - 1] == 0; ncll--);          /* :230-234 */

		write_bits(out, (uint32_t)(n_ll - 257), 5);                    /* :236-238 */
		write_bits(out, (uint32_t)(n_d - 1), 5);
		write_bits(out, (uint32_t)(ncll - 4), 4);
		for (int i = 0; i < ncll; i++) write_bits(out, (uint32_t)reordered[i], 3);

		uint32_t cl_code[19];
		lengths_to_codes(cl_len, 19, 7, cl_code);                      /* :243 */
		for (int i = 0; i < ncl; i++) {                                /* :244-256 */
			uint32_t pair = cl_code[cl_syms[i]];
			write_bits(out, pair >> 4, (int)(pair & 0xF));
			if (cl_syms[i] >= 16)
				write_bits(out, (uint32_t)cl_extra[i], cl_syms[i] == 16 ? 2 : cl_syms[i] == 17 ? 3 : 7);
		}
		lengths_to_codes(ll_len, n_ll, 15, ll_code_buf);               /* :260 */
		ll_code = ll_code_buf;
		if (no_dist) d_code = NULL;
		else { lengths_to_codes(d_len, n_d, 15, d_code_buf); d_code = d_code_buf; }
	}

	for (size_t i = 0; i < nt; ) {                                     /* :267-285 */
		int pair = toks[i++];
		int sym = pair >> 4, ne = pair & 0xF;
		uint32_t cp = ll_code[sym];
		write_bits(out, cp >> 4, (int)(cp & 0xF));
		if (sym > 256) {
			write_bits(out, toks[i++], ne);
			int dp = toks[i++];
			int dsym = dp >> 4, dne = dp & 0xF;
			uint32_t dc = d_code[dsym];
			write_bits(out, dc >> 4, (int)(dc & 0xF));
			write_bits(out, toks[i++], dne);
		}
	}
	free(toks);
}

/* Uncompressed.java:35-45 */
static void stored_emit(const uint8_t *b, size_t start, size_t end, BitOut *out, int is_final) {
	size_t index = start;
	do {
		size_t n = end - index < 65535 ? end - index : 65535;
		write_bits(out, (is_final && n == end - index) ? 1 : 0, 1);
		write_bits(out, 0, 2);
		write_bits(out, 0, (8 - bit_position(out)) % 8);
		write_bits(out, (uint32_t)n, 16);
		write_bits(out, (uint32_t)n ^ 0xFFFF, 16);
		for (size_t e = index + n; index < e; index++)
			write_bits(out, b[index], 8);
	} while (index < end);
}

/* Uncompressed.java:23-25 */
static int64_t stored_cost(size_t data_len, int i) {
	int64_t nb = (int64_t)((data_len + 65534) / 65535);
	if (nb < 1) nb = 1;
	return (int64_t)data_len * 8 + nb * 40 + ((13 - i) % 8 - 5);
}

size_t oracle_deflate_bound(size_t n, int lookahead) {
	size_t blocks = n / (size_t)(lookahead > 0 ? lookahead : 1) + 2;
	return n * 9 / 8 + n / 2 + blocks * 400 + 64;    /* literal-only static worst case is 9 bits/byte */
}

size_t oracle_deflate(const uint8_t *in, size_t n, const int *strategies, int n_strategies,
                      int lookahead, int history, int search, uint8_t *outbuf, size_t cap) {
	if (n_strategies < 1 || n_strategies > 8 || lookahead < 1 || history < 0 || history > 32768)
		return (size_t)-1;
	for (int i = 0; i < n_strategies; i++)
		if (strategies[i] < 0 || strategies[i] > ORC_STRAT_UNCOMPRESSED) return (size_t)-1;
	init_static();

	int need_chains = 0;
	for (int i = 0; i < n_strategies; i++)
		if (strategies[i] == ORC_STRAT_FULL_STATIC || strategies[i] == ORC_STRAT_FULL_DYNAMIC) need_chains = !search;
	Chains ch, *chp = NULL;
	if (need_chains && n >= 3) {
		ch.data = in; ch.n = n; ch.inserted = 0;
		ch.head = (int32_t *)malloc(sizeof(int32_t) << 24);
		ch.prev = (int32_t *)malloc(sizeof(int32_t) * n);
		memset(ch.head, 0xFF, sizeof(int32_t) << 24);
		chp = &ch;
	}

	BitOut out;
	memset(&out, 0, sizeof out);
	out.out = outbuf; out.cap = cap;

	/* DeflaterOutputStream.write/finish (:76-108): a block is flushed when the lookahead buffer is full and
	 * more data arrives; finish() flushes whatever remains (possibly zero bytes) as the final block. */
	size_t start = 0;
	for (;;) {
		size_t remaining = n - start;
		int is_final = remaining <= (size_t)lookahead;
		size_t end = is_final ? n : start + (size_t)lookahead;
		size_t hist_len = start < (size_t)history ? start : (size_t)history;   /* :129 */
		size_t off = start - hist_len;

		int chosen = strategies[0];
		if (n_strategies > 1) {                                               /* MultiStrategy.java:31-57 */
			int pos = bit_position(&out);
			int64_t best = INT64_MAX;
			for (int k = 0; k < n_strategies; k++) {
				int64_t cost;
				if (strategies[k] == ORC_STRAT_UNCOMPRESSED) cost = stored_cost(end - start, pos);
				else {
					BitOut cnt;
					memset(&cnt, 0, sizeof cnt);
					cnt.counting = 1;
					size_t mark = chp ? chp->inserted : 0;
					lz77_block(&PRESETS[strategies[k]], in, off, start, end, chp, search, &cnt, 0);
					if (chp) chains_rollback(chp, mark);
					cost = (int64_t)cnt.count;
				}
				if (cost < best) { best = cost; chosen = strategies[k]; }      /* strict '<': first listed wins ties */
			}
		}
		if (chosen == ORC_STRAT_UNCOMPRESSED) stored_emit(in, start, end, &out, is_final);
		else lz77_block(&PRESETS[chosen], in, off, start, end, chp, search, &out, is_final);
		if (is_final) break;
		start = end;
	}
	finish_bits(&out);
	if (chp) { free(ch.head); free(ch.prev); }
	return out.overflow ? (size_t)-1 : out.pos;
}


/* ---------- BinarySplit (comp/BinarySplit.java:30-98) over MultiStrategy / a single strategy ---------- */

typedef struct Dec {                /* one Decision object of BinarySplit.decide (:36-98) */
	size_t start, end;
	int64_t cost[8];                /* getBitLengths() */
	int leaf_strategy[8];           /* the sub-strategy MultiStrategy picked per start bit position (:35-44) */
	uint8_t use_split[8];           /* subdecisions[i] is the pair of halves (:62-69) */
	struct Dec *sub[2];
} Dec;

typedef struct {
	const uint8_t *in;
	size_t off;                     /* start of the history available to the enclosing block */
	const int *strategies;
	int n_strategies;
	Chains *ch;
	int search;
	int min_len;
} SplitCtx;

/* substrategy.decide(b, off, historyLen, dataLen): Lz77Huffman prices one pass on a counting sink and reports the
 * same cost for all 8 bit positions (Lz77Huffman.java:46-52), Uncompressed has a cost per position
 * (Uncompressed.java:23-25), MultiStrategy keeps the per-position minimum, the first listed winning ties (:35-44). */
static Dec *leaf_decide(const SplitCtx *c, size_t start, size_t end) {
	Dec *d = (Dec *)calloc(1, sizeof(Dec));
	d->start = start; d->end = end;
	for (int i = 0; i < 8; i++) { d->cost[i] = INT64_MAX; d->leaf_strategy[i] = c->strategies[0]; }
	for (int k = 0; k < c->n_strategies; k++) {
		int st = c->strategies[k];
		int64_t lz = 0;
		if (st != ORC_STRAT_UNCOMPRESSED) {
			BitOut cnt;
			memset(&cnt, 0, sizeof cnt);
			cnt.counting = 1;
			size_t mark = c->ch ? c->ch->inserted : 0;
			lz77_block(&PRESETS[st], c->in, c->off, start, end, c->ch, c->search, &cnt, 0);
			if (c->ch) chains_rollback(c->ch, mark);
			lz = (int64_t)cnt.count;
		}
		for (int i = 0; i < 8; i++) {
			int64_t cost = st == ORC_STRAT_UNCOMPRESSED ? stored_cost(end - start, i) : lz;
			if (cost < d->cost[i]) { d->cost[i] = cost; d->leaf_strategy[i] = st; }
		}
	}
	return d;
}

/* the accumulation of BinarySplit.java:49-53 / :60-64: it restarts at 0 for every i, i.e. the halves are always
 * priced as if the first one started byte-aligned */
static int64_t pair_bits(Dec *const *sp) {
	int64_t bits = 0;
	for (int k = 0; k < 2; k++) bits += sp[k]->cost[(int)(bits % 8)];
	return bits;
}

static void split_decide(const SplitCtx *c, Dec *cur) {                /* BinarySplit.decide(b, off, hist, len, curDec) */
	size_t len = cur->end - cur->start;
	size_t first = (len + 1) / 2, second = len - first;
	size_t mn = first < second ? first : second;
	if (mn <= (size_t)c->min_len) return;                              /* :42 */
	Dec *sp[2] = {leaf_decide(c, cur->start, cur->start + first), leaf_decide(c, cur->start + first, cur->end)};
	int64_t bits = pair_bits(sp);
	int improved = 0;
	for (int i = 0; i < 8; i++) improved |= bits < cur->cost[i];       /* :47-54 */
	if (improved) { split_decide(c, sp[0]); split_decide(c, sp[1]); }  /* :56-59 */
	bits = pair_bits(sp);
	int used = 0;
	for (int i = 0; i < 8; i++)                                        /* :60-69 */
		if (bits < cur->cost[i]) { cur->cost[i] = bits; cur->use_split[i] = 1; used = 1; }
	cur->sub[0] = sp[0]; cur->sub[1] = sp[1];
	(void)used;
}

static void split_emit(const SplitCtx *c, const Dec *d, BitOut *out, int is_final) {   /* compressTo :78-82 */
	int pos = bit_position(out);
	if (d->use_split[pos]) {
		split_emit(c, d->sub[0], out, 0);
		split_emit(c, d->sub[1], out, is_final);
		return;
	}
	int st = d->leaf_strategy[pos];                                    /* MultiStrategy dispatches on the position again (:54) */
	if (st == ORC_STRAT_UNCOMPRESSED) stored_emit(c->in, d->start, d->end, out, is_final);
	else lz77_block(&PRESETS[st], c->in, c->off, d->start, d->end, c->ch, c->search, out, is_final);
}

static void split_free(Dec *d) {
	if (!d) return;
	split_free(d->sub[0]);
	split_free(d->sub[1]);
	free(d);
}

size_t oracle_deflate_split(const uint8_t *in, size_t n, const int *strategies, int n_strategies,
                            int lookahead, int history, int search, int min_block_len,
                            uint8_t *outbuf, size_t cap, size_t *n_blocks) {
	if (n_strategies < 1 || n_strategies > 8 || lookahead < 1 || history < 0 || history > 32768 || min_block_len < 1)
		return (size_t)-1;
	for (int i = 0; i < n_strategies; i++)
		if (strategies[i] < 0 || strategies[i] > ORC_STRAT_UNCOMPRESSED) return (size_t)-1;
	init_static();
	int need_chains = 0;
	for (int i = 0; i < n_strategies; i++)
		if (strategies[i] == ORC_STRAT_FULL_STATIC || strategies[i] == ORC_STRAT_FULL_DYNAMIC) need_chains = !search;
	Chains ch, *chp = NULL;
	if (need_chains && n >= 3) {
		ch.data = in; ch.n = n; ch.inserted = 0;
		ch.head = (int32_t *)malloc(sizeof(int32_t) << 24);
		ch.prev = (int32_t *)malloc(sizeof(int32_t) * n);
		memset(ch.head, 0xFF, sizeof(int32_t) << 24);
		chp = &ch;
	}
	BitOut out;
	memset(&out, 0, sizeof out);
	out.out = outbuf; out.cap = cap;
	size_t start = 0, blocks = 0;
	for (;;) {                                                         /* DeflaterOutputStream framing, as oracle_deflate */
		size_t remaining = n - start;
		int is_final = remaining <= (size_t)lookahead;
		size_t end = is_final ? n : start + (size_t)lookahead;
		size_t hist_len = start < (size_t)history ? start : (size_t)history;
		SplitCtx c = {in, start - hist_len, strategies, n_strategies, chp, search, min_block_len};
		Dec *top = leaf_decide(&c, start, end);                        /* BinarySplit.decide :30-33 */
		split_decide(&c, top);
		size_t before = out.pos;
		(void)before;
		split_emit(&c, top, &out, is_final);
		/* count the DEFLATE blocks this outer block became (leaves actually emitted at their positions are
		 * not recoverable after the fact; count tree leaves reachable through position-0 decisions as a gauge) */
		{
			const Dec *stack[64]; int sp_ = 0; stack[sp_++] = top;
			while (sp_) { const Dec *d = stack[--sp_]; if (d->use_split[0] && sp_ < 62) { stack[sp_++] = d->sub[0]; stack[sp_++] = d->sub[1]; } else blocks++; }
		}
		split_free(top);
		if (is_final) break;
		start = end;
	}
	finish_bits(&out);
	if (chp) { free(ch.head); free(ch.prev); }
	if (n_blocks) *n_blocks = blocks;
	return out.overflow ? (size_t)-1 : out.pos;
}
