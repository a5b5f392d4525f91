/*
 * corpus.c -- deterministic synthetic corpora for benchmarks and tests.
 *
 * The reference (nayuki/DEFLATE-library-Java) ships no data files; its tests build inputs from an
 * unseeded java.util.Random (DeflaterOutputStreamTest.java:118).  BASELINE.json's configs name
 * "synthetic text-like", "mixed-entropy", "incompressible random" and "all-zero" inputs; SURVEY.md
 * Appendix D fixes their construction so every run sees the same bytes.  Host-side utility, exported
 * from libb2deflate.so as b2d_corpus_* (include/b2deflate.h); it is data generation, not codec code.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(__GNUC__)
#define B2D_EXPORT __attribute__((visibility("default")))
#else
#define B2D_EXPORT
#endif


static inline uint64_t splitmix(uint64_t *s) {
	uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

B2D_EXPORT void b2d_corpus_random(uint64_t seed, uint8_t *out, size_t n) {
	uint64_t s = seed;
	size_t i = 0;
	while (i < n) {
		uint64_t x = splitmix(&s);
		for (int k = 0; k < 8 && i < n; k++, i++) out[i] = (uint8_t)(x >> (8 * k));
	}
}

#define VOCAB 8192
typedef struct { uint8_t len; char w[10]; } Word;

static void make_vocab(uint64_t seed, Word *v) {
	static const char LETTERS[] = "etaoinshrdlucmfwypvbgkqjxz";
	uint64_t s = seed ^ 0x5EED0001ull;
	for (int i = 0; i < VOCAB; i++) {
		uint64_t x = splitmix(&s);
		v[i].len = (uint8_t)(1 + x % 10);
		for (int k = 0; k < v[i].len; k++) {
			x = splitmix(&s);
			int a = (int)(x % 26), b = (int)((x >> 8) % 26);
			v[i].w[k] = LETTERS[a < b ? a : b];
		}
	}
}

B2D_EXPORT void b2d_corpus_text(uint64_t seed, uint8_t *out, size_t n) {
	Word *v = (Word *)malloc(sizeof(Word) * VOCAB);
	make_vocab(seed, v);
	uint64_t s = seed;
	size_t i = 0;
	while (i < n) {
		uint64_t x = splitmix(&s);
		int e = (int)(x % 14);
		uint64_t idx;
		if (e < 13) idx = (((uint64_t)1 << e) - 1 + ((x >> 8) & (((uint64_t)1 << e) - 1))) % VOCAB;
		else idx = (x >> 8) % VOCAB;
		const Word *w = &v[idx];
		for (int k = 0; k < w->len && i < n; k++) out[i++] = (uint8_t)w->w[k];
		int p = (int)((x >> 40) % 16);
		if (p == 0) { if (i < n) out[i++] = '.'; if (i < n) out[i++] = '\n'; }
		else if (p == 1) { if (i < n) out[i++] = ','; if (i < n) out[i++] = ' '; }
		else { if (i < n) out[i++] = ' '; }
	}
	free(v);
}

B2D_EXPORT void b2d_corpus_mixed(uint64_t seed, uint8_t *out, size_t n) {
	uint64_t s = seed ^ 0xA11CEull;
	uint64_t text_seed = seed;
	uint32_t counter = 0;
	size_t i = 0;
	while (i < n) {
		uint64_t x = splitmix(&s);
		size_t L = (size_t)4096 << (x % 6);
		if (L > n - i) L = n - i;
		int t = (int)((x >> 8) % 5);
		switch (t) {
		case 0:
			b2d_corpus_text(++text_seed, out + i, L);
			break;
		case 1:
			b2d_corpus_random(x, out + i, L);
			break;
		case 2:
			memset(out + i, 0, L);
			break;
		case 3: {
			uint64_t rs = x;
			for (size_t k = 0; k < L; k += 16) {
				uint8_t rec[16];
				memset(rec, 0, sizeof rec);
				uint32_t c = counter++;
				rec[0] = (uint8_t)c; rec[1] = (uint8_t)(c >> 8); rec[2] = (uint8_t)(c >> 16); rec[3] = (uint8_t)(c >> 24);
				uint64_t r = splitmix(&rs);
				rec[6] = (uint8_t)(r % 4);
				rec[7] = 0x80;
				uint64_t r2 = splitmix(&rs);
				rec[8] = (uint8_t)r2; rec[9] = (uint8_t)(r2 >> 8);
				size_t m = L - k < 16 ? L - k : 16;
				memcpy(out + i + k, rec, m);
			}
			break;
		}
		default: {
			if (i < 1024) { memset(out + i, 0, L); break; }
			size_t lim = i < 24576 ? i : 24576;
			size_t D = 1 + (size_t)((x >> 16) % lim);
			for (size_t k = 0; k < L; k++) out[i + k] = out[i + k - D];
			break;
		}
		}
		i += L;
	}
}
