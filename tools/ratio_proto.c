/* ratio_proto.c -- CPU model of the GPU match-finder/parse (design tool, not product, not shipped).
 * Evaluates compressed size for: hash bytes, hash bits, chain depth, nice length, lazy on/off.
 * Build: gcc -O2 -o /tmp/ratio_proto tools/ratio_proto.c oracle/oracle_deflate.c oracle/oracle_inflate.c oracle/oracle_misc.c deflate-library-java_b200/csrc/corpus.c
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "../oracle/oracle.h"
void b2d_corpus_text(uint64_t, uint8_t*, size_t);
void b2d_corpus_mixed(uint64_t, uint8_t*, size_t);
void b2d_corpus_random(uint64_t, uint8_t*, size_t);

static int HB = 4, HASH_BITS = 15, DEPTH = 16, NICE = 258, LAZY = 1, MIN3FAR = 4096, GOOD=32;

static inline uint32_t hashf(const uint8_t *p) {
	uint32_t v = p[0] | p[1] << 8 | p[2] << 16 | (HB == 4 ? (uint32_t)p[3] << 24 : 0);
	return (v * 2654435761u) >> (32 - HASH_BITS);
}
static int nlz32(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static const int ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

static long dyn_cost(int *ll, int *dh, size_t data_len) {
	int n_ll = 286, n_d = 30;
	if (data_len == 0) ll[0]++;
	for (; n_ll > 257 && ll[n_ll - 1] == 0; n_ll--);
	uint8_t ll_len[286], d_len[30];
	oracle_package_merge(ll, n_ll, 15, ll_len);
	int used = 0; for (int i = 0; i < 30; i++) if (dh[i] > 0) used++;
	if (used == 1) for (int i = 0; i < 30; i++) if (dh[i] > 0) { if (i < 29) dh[i + 1] = 1; else dh[i - 1] = 1; break; }
	for (; n_d > 1 && dh[n_d - 1] == 0; n_d--);
	int no_dist = (n_d == 1 && dh[0] == 0);
	if (no_dist) d_len[0] = 0; else oracle_package_merge(dh, n_d, 15, d_len);
	uint8_t lens[316]; int total = n_ll + n_d;
	memcpy(lens, ll_len, n_ll); memcpy(lens + n_ll, d_len, n_d);
	int cl_hist[19] = {0}; long extra = 0;
	for (int i = 0; i < total;) {
		int val = lens[i];
		if (val == 0) { int rl = 1; for (; rl < 138 && i + rl < total && lens[i + rl] == 0; rl++);
			if (rl < 3) { cl_hist[0]++; i++; } else if (rl < 11) { cl_hist[17]++; extra += 3; i += rl; } else { cl_hist[18]++; extra += 7; i += rl; } continue; }
		if (i > 0) { int rl = 0; for (; rl < 6 && i + rl < total && lens[i + rl] == lens[i - 1]; rl++);
			if (rl >= 3) { cl_hist[16]++; extra += 2; i += rl; continue; } }
		cl_hist[val]++; i++;
	}
	uint8_t cl_len[19]; oracle_package_merge(cl_hist, 19, 7, cl_len);
	int ncll = 19; for (; ncll > 4 && cl_len[ORDER[ncll - 1]] == 0; ncll--);
	long bits = 3 + 14 + 3 * ncll + extra;
	for (int i = 0; i < 19; i++) bits += (long)cl_hist[i] * cl_len[i];
	for (int s = 0; s < n_ll; s++) { bits += (long)ll[s] * ll_len[s]; if (s >= 265 && s < 285) bits += (long)ll[s] * ((s - 261) / 4); }
	for (int s = 0; s < n_d; s++) { bits += (long)dh[s] * d_len[s]; if (s >= 4) bits += (long)dh[s] * (s / 2 - 1); }
	return bits;
}
static long fixed_cost(const int *ll, const int *dh) {
	long bits = 3;
	for (int s = 0; s < 286; s++) { int l = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8; bits += (long)ll[s] * l; if (s >= 265 && s < 285) bits += (long)ll[s] * ((s - 261) / 4); }
	for (int s = 0; s < 30; s++) { bits += (long)dh[s] * 5; if (s >= 4) bits += (long)dh[s] * (s / 2 - 1); }
	return bits;
}

typedef struct { uint16_t len, dist; } M;

static long compress_chunk(const uint8_t *b, size_t n, size_t block, long *tok_count) {
	uint16_t *head = calloc(1u << HASH_BITS, 2);      /* stores pos+1 low 16 bits? use int32 for the model */
	int32_t *headp = malloc(sizeof(int32_t) << HASH_BITS);
	memset(headp, 0xFF, sizeof(int32_t) << HASH_BITS);
	int32_t *prev = malloc(sizeof(int32_t) * (n + 1));
	M *m = calloc(n + 1, sizeof(M));
	for (size_t p = 0; p + HB <= n; p++) { uint32_t h = hashf(b + p); prev[p] = headp[h]; headp[h] = (int32_t)p; }
	for (size_t p = 0; p < n; p++) {
		size_t bend = (p / block + 1) * block; if (bend > n) bend = n;
		int maxlen = bend - p < 258 ? (int)(bend - p) : 258;
		int best = 0, bdist = 0;
		if (p + HB <= n && maxlen >= 3) {
			int depth = DEPTH;
			for (int32_t c = prev[p]; c >= 0 && depth > 0; c = prev[c], depth--) {
				size_t dist = p - c; if (dist > 32768) break;
				int l = 0; while (l < maxlen && b[c + l] == b[p + l]) l++;
				if (l > best) { best = l; bdist = (int)dist; if (l >= NICE || l >= maxlen) break; }
				if (best >= GOOD && depth > DEPTH / 4) depth = DEPTH/4;
			}
		}
		if (best < 3 || (best == 3 && bdist > MIN3FAR)) best = 0;
		m[p].len = (uint16_t)best; m[p].dist = (uint16_t)(bdist - 1);
	}
	long total_bits = 0;
	for (size_t bs = 0; bs < n || bs == 0; bs += block) {
		size_t be = bs + block < n ? bs + block : n;
		int ll[286] = {0}, dh[30] = {0};
		for (size_t i = bs; i < be;) {
			int len = m[i].len;
			if (len >= 3 && LAZY && i + 1 < be && m[i + 1].len > len) len = 0;
			if (len >= 3) {
				int r = len - 3, sym;
				if (len < 11) sym = 257 + r; else if (len == 258) sym = 285; else { int ne = 29 - nlz32(r); sym = (ne << 2) + (r >> ne) + 257; }
				ll[sym]++;
				int d = m[i].dist, ds; if (d < 4) ds = d; else { int ne = 30 - nlz32(d); ds = (ne << 1) + (d >> ne); }
				dh[ds]++; i += len;
			} else { ll[b[i]]++; i++; }
			(*tok_count)++;
		}
		ll[256]++;
		long fc = fixed_cost(ll, dh);
		long dc = dyn_cost(ll, dh, be - bs);
		long sc = (long)(be - bs) * 8 + 40 * ((be - bs + 65534) / 65535 > 0 ? (be - bs + 65534) / 65535 : 1) + 3;
		long best = dc < fc ? dc : fc; if (sc < best) best = sc;
		total_bits += best;
		if (n == 0) break;
	}
	total_bits += 3 + 7 + 32; /* sync flush (upper bound) */
	free(head); free(headp); free(prev); free(m);
	return (total_bits + 7) / 8;
}

int main(int argc, char **argv) {
	const char *kind = argc > 1 ? argv[1] : "text";
	size_t mib = argc > 2 ? atoi(argv[2]) : 4;
	if (argc > 3) HB = atoi(argv[3]);
	if (argc > 4) DEPTH = atoi(argv[4]);
	if (argc > 5) NICE = atoi(argv[5]);
	if (argc > 6) LAZY = atoi(argv[6]);
	if (argc > 7) MIN3FAR = atoi(argv[7]);
	if (argc > 8) GOOD = atoi(argv[8]);
	int do_ref = argc > 9 ? atoi(argv[9]) : 0;
	size_t n = mib << 20;
	uint8_t *buf = malloc(n);
	if (!strcmp(kind, "text")) b2d_corpus_text(0xDEF1A7E, buf, n);
	else if (!strcmp(kind, "mixed")) b2d_corpus_mixed(0xDEF1A7E, buf, n);
	else b2d_corpus_random(0xDEF1A7E, buf, n);
	long total = 0, ref = 0, toks = 0;
	size_t chunk = 1 << 20;
	for (size_t c = 0; c < n; c += chunk) {
		total += compress_chunk(buf + c, chunk, 65536, &toks);
		if (do_ref) {
			int s = ORC_STRAT_FULL_DYNAMIC;
			size_t cap = oracle_deflate_bound(chunk, 65536);
			uint8_t *o = malloc(cap);
			ref += (long)oracle_deflate(buf + c, chunk, &s, 1, 65536, 32768, 0, o, cap);
			free(o);
		}
	}
	printf("%s %zuMiB HB=%d depth=%d nice=%d lazy=%d far3=%d good=%d: %ld bytes ratio %.4f toks/byte %.3f", kind, mib, HB, DEPTH, NICE, LAZY, MIN3FAR, GOOD, total, (double)n / total, (double)toks / n);
	if (do_ref) printf("  | FULL_DYNAMIC %ld  (ours/ref = %.4f)", ref, (double)total / ref);
	printf("\n");
	return 0;
}
