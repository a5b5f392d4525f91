// deflate.cu -- chunked DEFLATE encoder for sm_100a.
//
// Replaces the encode path of the reference (paths relative to src/io/nayuki/deflate/):
//   comp/Lz77Huffman.java:68-130   greedy brute-force match loop   -> chains_kernel + match_kernel + parse_kernel
//   comp/Lz77Huffman.java:66-67,89,109,125,132  histograms          -> parse_kernel (shared-memory atomics)
//   comp/Lz77Huffman.java:143-265,309-335,372-391  package-merge, code-length RLE, header, canonical codes
//                                                                   -> huffman_kernel (one warp per block)
//   comp/Uncompressed.java:19-48, comp/MultiStrategy.java:31-57     -> layout_kernel (cost rule per start bit)
//   comp/Lz77Huffman.java:267-285 + DeflaterOutputStream.java:141-171 bit emission -> emit_kernel
//   DeflaterOutputStream.java:119-137 framing                       -> independent chunks joined by empty stored blocks
//
// Pipeline (all on one stream, no host round trips):
//   chains : one warp per 256 KiB segment of a chunk re-inserts the preceding 32 KiB and links every position to the
//            previous position with the same 13-bit hash of its next 4 bytes (u16 distance per input byte)
//   match  : one CTA per 32 KiB tile; the 64 KiB window and its links are staged in shared memory with 128-bit
//            loads; every thread takes runs of 16 consecutive positions and walks each position's chain
//            (depth-limited), candidates compared 12 bytes at a time from registers -> (len, dist) per position
//   parse  : one warp per parse unit (16 KiB piece or block): what every position would emit is worked out for 32
//            positions at once, the warp follows the greedy / lazy chain through them, the visited positions
//            write their tokens and count their symbols (lit/len + distance histograms)
//   huffman: one warp per block: exact package-merge (limit 15 / 7), code-length RLE, header bits, block costs
//   layout : one thread per chunk picks stored / fixed / dynamic per block and assigns bit offsets;
//            scan_kernel turns chunk sizes into byte offsets
//   emit   : one CTA per block: prefix sum of code lengths, bits OR-ed into a shared staging tile, coalesced out
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"

namespace b2d {

constexpr int HASH_BITS = 13;                // 16 KiB head table per warp: 14 chain warps per SM
constexpr u32 WINDOW = 32768;
constexpr int MAX_MATCH = 258;
constexpr u32 TILE = 32768;                 // positions per match CTA
constexpr int MATCH_THREADS = 1024;
constexpr u32 MATCH_DATA_WORDS = (WINDOW + TILE + 272) / 4;
constexpr size_t MATCH_SMEM = MATCH_DATA_WORDS * 4 + (size_t)(WINDOW + TILE) * 2;
constexpr int HDR_WORDS = 144;

__constant__ u8 CL_ORDER_D[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};   // Lz77Huffman.java:368-369

// match / token entry: [31:24] literal byte at the position, [23:15] length (0 = literal), [14:0] distance - 1
__device__ __forceinline__ u32 tok_len(u32 e) { return (e >> 15) & 0x1FF; }
__device__ __forceinline__ u32 tok_dist(u32 e) { return (e & 0x7FFF) + 1; }

struct BlockRec {
	u32 hist_ll[288];
	u32 hist_d[32];
	u32 code_ll[288];       // bit-reversed code << 4 | length  (Lz77Huffman.codeLengthsToCodes, :372-391)
	u32 code_d[32];
	u32 hdr[HDR_WORDS];     // dynamic header after the 3 block-header bits, LSB-first
	u32 hdr_bits;
	u32 n_tokens;           // without the end-of-block symbol
	u32 cost_fixed;         // whole block in bits, fixed code    (3 + symbols + extra bits)
	u32 cost_dyn;           // whole block in bits, dynamic code  (3 + header + symbols + extra bits)
	u32 mode;               // set by layout: 1 stored, 2 fixed, 3 dynamic
	u32 bfinal;
	u64 out_bit;            // bit offset of the block inside its chunk's output
	u64 tok_bit;            // adaptive splitting, piece records: bit offset of the piece's first token
	// adaptive splitting (split_decide_kernel): for piece (leaf) records
	u32 owner;              // heap index of the node that is emitted as the block holding this piece
	u32 first_last;         // bit 0: first piece of that block, bit 1: last piece
	u32 body_dyn, body_fixed;   // bits of this piece's tokens (without end-of-block) under the block's dynamic / the fixed code
	u32 best_cost;          // any node: cheapest cost in bits of its span (one block, or the best split below it)
};

// Adaptive splitting keeps the records of one block_bytes span as a complete binary tree in heap order: node 0 is the
// whole span, nodes 2h+1 / 2h+2 are the halves of node h, the K = block_bytes / leaf_bytes pieces are the leaves
// K-1 .. 2K-2.  Without splitting K = 1 and the tree is the single record there has always been.
struct Heap {
	u32 K, H;               // leaves and nodes per span (H = 2K - 1)
	u32 block_bytes, leaf_bytes;
	u32 roots_only;         // pieces are only a parse unit: every span is emitted as ONE block (see heap_for)
};
__device__ __host__ __forceinline__ u32 heap_level(u32 h) { u32 l = 0; for (u32 v = h + 1; v > 1; v >>= 1) l++; return l; }
// byte range of node h of span `root`
__device__ __forceinline__ void heap_range(const Heap &hp, u32 root, u32 h, u64 n, u64 &start, u64 &len) {
	const u32 lvl = heap_level(h), idx = h + 1 - (1u << lvl), size = hp.block_bytes >> lvl;
	const u64 s0 = (u64)root * hp.block_bytes + (u64)idx * size;
	start = min(n, s0);
	len = min(n, s0 + size) - start;
}
__device__ __forceinline__ u32 leaf_slot(const Heap &hp, u32 leaf) {       // record index of piece number `leaf`
	return (leaf / hp.K) * hp.H + hp.K - 1 + leaf % hp.K;
}

// ---------------------------------------------------------------- helpers
__device__ __forceinline__ u32 load4_global(const u8 *in, u64 p, u64 n_words) {   // unaligned LE 4-byte read
	const u32 *w = (const u32 *)in;
	u64 i = p >> 2;
	u32 a = i < n_words ? __ldg(w + i) : 0u;
	u32 b = i + 1 < n_words ? __ldg(w + i + 1) : 0u;
	return __funnelshift_r(a, b, (u32)(p & 3) * 8);
}
__device__ __forceinline__ u32 hash4(u32 v, int hb) {
	if (hb == 3) v &= 0xFFFFFFu;
	return (v * 2654435761u) >> (32 - HASH_BITS);
}
__device__ __forceinline__ int len_symbol(int len, int &ne, int &extra) {          // Lz77Huffman.java:93-107
	int r = len - 3;
	if (len < 11) { ne = 0; extra = 0; return r + 257; }
	if (len == 258) { ne = 0; extra = 0; return 285; }
	ne = 29 - __clz(r);
	extra = r & ((1 << ne) - 1);
	return (ne << 2) + (r >> ne) + 257;
}
__device__ __forceinline__ int dist_symbol(int dist, int &ne, int &extra) {        // Lz77Huffman.java:113-123
	int d = dist - 1;
	if (dist < 5) { ne = 0; extra = 0; return d; }
	ne = 30 - __clz(d);
	extra = d & ((1 << ne) - 1);
	return (ne << 1) + (d >> ne);
}
__device__ __forceinline__ int ll_extra_bits(int sym) { return (sym >= 265 && sym < 285) ? (sym - 261) >> 2 : 0; }
__device__ __forceinline__ int d_extra_bits(int sym) { return sym >= 4 ? (sym >> 1) - 1 : 0; }
__device__ __forceinline__ int fixed_ll_len(int sym) { return sym < 144 ? 8 : sym < 256 ? 9 : sym < 280 ? 7 : 8; }

__device__ __forceinline__ u32 warp_sum(u32 v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
	return v;
}

// ---------------------------------------------------------------- K1a: hash chains
// One warp per segment of the input (SEG bytes, never crossing a chunk).  The warp first re-inserts up to 32 KiB
// before its segment so the head table is what a sequential insert would have left, then links every position of
// the segment.  32 positions per step: match.any finds equal hashes inside the step, the shared head table (low 16
// bits of the chunk-relative position per hash) gives the link to earlier steps.
constexpr u32 CHAIN_SEG = 256u << 10;

__global__ void __launch_bounds__(32)
chains_kernel(const u8 *__restrict__ in, u64 n, u32 chunk_bytes, u32 seg_bytes, int hb, u16 *__restrict__ prevdist) {
	extern __shared__ __align__(16) u16 head[];      // 1 << HASH_BITS entries
	const u32 lane = threadIdx.x;
	const u32 segs_per_chunk = (chunk_bytes + seg_bytes - 1) / seg_bytes;
	const u64 cs = (u64)(blockIdx.x / segs_per_chunk) * chunk_bytes;
	const u64 ss = cs + (u64)(blockIdx.x % segs_per_chunk) * seg_bytes;
	const u64 ce = min(n, cs + chunk_bytes);
	if (ss >= ce) return;
	const u64 se = min(ce, ss + seg_bytes);
	const u64 ws = (ss - cs > WINDOW) ? ss - WINDOW : cs;       // warm-up start
	for (u32 i = lane; i < (2u << HASH_BITS) / 16; i += 32) ((uint4 *)head)[i] = make_uint4(0, 0, 0, 0);
	__syncwarp();
	// everything below in 32-bit offsets from `org` (4-byte aligned, <= ws)
	const u64 org = ws & ~(u64)3;
	const u32 *__restrict__ words = (const u32 *)(in + org);
	const u32 n_words = (u32)min((u64)0x7FFFFFFF, ((n - org) + 3) >> 2);
	const u32 o_ws = (u32)(ws - org), o_ss = (u32)(ss - org), o_se = (u32)(se - org);
	const u32 o_ce = (u32)min((u64)0x7FFFFFFF, ce - org);      // (clamped: only compared against offsets of this segment)
	const u32 rel0 = (u32)(org - cs);                            // chunk-relative position of offset 0
	const u32 ws_rel = (u32)(ws - cs);
	u16 *__restrict__ pd = prevdist + org;
	const u32 cmask = hb == 3 ? 0xFFFFFFu : 0xFFFFFFFFu;
	// Input: one coalesced load brings 32 consecutive words = the bytes of FOUR steps; the four bytes of a lane's
	// position are taken from the lanes that hold them with shuffles.  Three more groups are in flight (16 steps of
	// lookahead), so the global-load latency is off the serial chain of head-table updates.
#define LOADG(g) (((g) * 32u + lane) < n_words ? __ldg(words + (g) * 32u + lane) : 0u)
	const u32 g_first = o_ws >> 7;                             // group = 128 bytes; positions before o_ws are skipped
	u32 G0 = LOADG(g_first), G1 = LOADG(g_first + 1), G2 = LOADG(g_first + 2), G3 = LOADG(g_first + 3);
	for (u32 g = g_first; g * 128u < o_se; g++) {
		const u32 G4 = LOADG(g + 4);
		// the four steps' hashes first (independent of the head table), then the serial part of each step
		u32 hs[4];
#pragma unroll
		for (u32 s = 0; s < 4; s++) {
			const u32 w0 = __shfl_sync(FULL_MASK, G0, 8 * s + (lane >> 2));
			u32 w1 = __shfl_sync(FULL_MASK, G0, (8 * s + (lane >> 2) + 1) & 31);
			if (s == 3) { const u32 wn = __shfl_sync(FULL_MASK, G1, 0); if (lane >= 28) w1 = wn; }
			const u32 v = __funnelshift_r(w0, w1, (lane & 3) * 8) & cmask;
			hs[s] = (v * 2654435761u) >> (32 - HASH_BITS);
		}
#pragma unroll
		for (u32 s = 0; s < 4; s++) {
			const u32 o = g * 128u + s * 32u + lane;
			const bool valid = o >= o_ws && o < o_se && o + (u32)hb <= o_ce;
			const u32 h = hs[s];
			const u32 prel = rel0 + o;
			u32 dist = 0;
			if (valid) {                                            // link to the newest position of earlier steps
				u32 d = (prel - head[h]) & 0xFFFFu;
				if (d == 0) d = 65536;
				if (d <= WINDOW && d <= prel - ws_rel) dist = d;
			}
			__syncwarp();
			if (valid) head[h] = (u16)prel;                         // equal hashes in this step: one of them wins ...
			__syncwarp();
			// ... then the step is sorted out exactly, one colliding hash at a time (match.any would loop over every
			// DISTINCT hash of the step, about 30 rounds for a single collision; this loops over the colliding ones)
			u32 lostmask = __ballot_sync(FULL_MASK, valid && head[h] != (u16)prel);
			while (lostmask) {
				const u32 hl = __shfl_sync(FULL_MASK, h, __ffs(lostmask) - 1);
				const bool mine = valid && h == hl;
				const u32 grp = __ballot_sync(FULL_MASK, mine);
				const u32 lower = grp & lanemask_lt();
				if (mine && lower) dist = lane - (31 - __clz(lower));     // same hash earlier in this step
				if (mine && (grp >> lane) == 1u) head[h] = (u16)prel;     // the highest lane of the group stays
				lostmask &= ~grp;
			}
			__syncwarp();
			if (o >= o_ss && o < o_se) pd[o] = (u16)dist;
		}
		G0 = G1; G1 = G2; G2 = G3; G3 = G4;
	}
#undef LOADG
}

// ---------------------------------------------------------------- K1b: match search
struct MatchParams {
	int search, hb, depth, nice;
	int skip_min;      // positions 2 .. L-1 behind the start of a searched match of length L >= skip_min are not searched
	int good_len;      // an inherited match at least this long quarters the chain depth (zlib's good_length idea)
};

__device__ __forceinline__ u32 sm_load4(const u32 *W, u32 o) {
	return __funnelshift_r(W[o >> 2], W[(o >> 2) + 1], (o & 3) * 8);
}
// number of equal bytes of a[0..] and b[0..], at most maxlen; 8 bytes per step (three aligned words per side)
__device__ __forceinline__ int sm_match_len(const u32 *W, u32 a, u32 b, int maxlen) {
	int len = 0;
	while (len < maxlen) {
		const u32 ia = (a + len) >> 2, ib = (b + len) >> 2;
		const u32 sa = ((a + len) & 3) * 8, sb = ((b + len) & 3) * 8;
		const u32 a0 = W[ia], a1 = W[ia + 1], a2 = W[ia + 2];
		const u32 b0 = W[ib], b1 = W[ib + 1], b2 = W[ib + 2];
		const u32 x0 = __funnelshift_r(a0, a1, sa) ^ __funnelshift_r(b0, b1, sb);
		const u32 x1 = __funnelshift_r(a1, a2, sa) ^ __funnelshift_r(b1, b2, sb);
		if (x0) { len += (__ffs(x0) - 1) >> 3; break; }
		if (x1) { len += 4 + ((__ffs(x1) - 1) >> 3); break; }
		len += 8;
	}
	return len < maxlen ? len : maxlen;
}

__global__ void __launch_bounds__(MATCH_THREADS, 1)
match_kernel(const u8 *__restrict__ in, u64 n, u32 chunk_bytes, u32 block_bytes, MatchParams mp,
             const u16 *__restrict__ prevdist, u32 *__restrict__ match) {
	extern __shared__ __align__(16) u32 msm[];
	u32 *W = msm;                                     // window bytes as words, W[0] = byte win_start
	u16 *P = (u16 *)(msm + MATCH_DATA_WORDS);         // links for [win_start, tile_end)
	const u64 ts = (u64)blockIdx.x * TILE;
	const u64 te = min(n, ts + TILE);
	const u64 cs = ts / chunk_bytes * chunk_bytes;
	const u64 ce = min(n, cs + chunk_bytes);
	const u64 win = (ts - cs > WINDOW) ? ts - WINDOW : cs;
	const u64 n16 = (n + 15) >> 4;
	{   // stage data: 128-bit loads, zero past the end of input
		const uint4 *src = (const uint4 *)in + (win >> 4);
		const u64 first16 = win >> 4;
		uint4 *dst = (uint4 *)W;
		for (u32 i = threadIdx.x; i < MATCH_DATA_WORDS / 4; i += MATCH_THREADS)
			dst[i] = (first16 + i < n16) ? __ldg(src + i) : make_uint4(0, 0, 0, 0);
		if (mp.search == B2D_SEARCH_DEFAULT || mp.search == 3) {
			const uint4 *ps = (const uint4 *)(prevdist + win);
			uint4 *pd = (uint4 *)P;
			const u32 cnt16 = (u32)((te - win + 7) >> 3);
			const u64 p16_lim = (n + 7) >> 3;
			for (u32 i = threadIdx.x; i < cnt16; i += MATCH_THREADS)
				pd[i] = ((win >> 3) + i < p16_lim) ? __ldg(ps + i) : make_uint4(0, 0, 0, 0);
		}
	}
	__syncthreads();
	const u32 woff = (u32)(ts - win);                 // offset of the tile inside the window
	// Each thread takes a run of consecutive positions, so that in the default search a match found at p is
	// inherited by p + 1 (same distance, one byte shorter, then extended): positions inside long repeats cost a few
	// instructions instead of a full 258-byte comparison per candidate.  The exact searches (RLE / FULL) do not
	// inherit: they must return the longest match with ties to the smallest distance (Lz77Huffman.java:71-84).
	const u32 n_pos = (u32)(te - ts);
	const bool inherit = mp.search == B2D_SEARCH_DEFAULT;
	constexpr u32 RUN = 16;                           // consecutive positions per thread and round: four 16-byte stores per lane
	const u32 cmask = mp.hb == 3 ? 0xFFFFFFu : 0xFFFFFFFFu;
	for (u32 round = 0; round < TILE / (MATCH_THREADS * RUN); round++) {
	const u32 k0 = (round * MATCH_THREADS + threadIdx.x) * RUN;
	if (k0 >= n_pos) break;
	// positions relative to the chunk start fit 32 bits; the end of the current block is found once per run and moved
	// on when a position reaches it (no 64-bit division per position)
	const u32 rel0 = (u32)(ts + k0 - cs), ce_rel = (u32)(ce - cs);
	u32 be_rel = (rel0 / block_bytes + 1) * block_bytes;
	int prev_len = 0, prev_dist = 0;
	int run_len = 0, since = 1 << 20;                 // length of / positions since the last searched long match
	const int SKIP_MIN = mp.skip_min;
	// The run's own bytes (16 + 11: every position's first 12 bytes) and links (16) are read ONCE, with four 128-bit
	// loads, and kept in registers: a position's bytes are funnel shifts of them, its link the low end of a 64-bit
	// window that slides by a link per position.  (Read per position they were 3 shared-memory loads at 4- and 8-way bank
	// conflicts -- runs of 16 positions put the lanes 16 bytes apart -- and half of the kernel's shared wavefronts.)
	const u32 o0 = woff + k0;                        // multiple of 16
	const uint4 own_wa = *(const uint4 *)(W + (o0 >> 2)), own_wb = *(const uint4 *)(W + (o0 >> 2) + 4);
	const uint4 own_pa = *(const uint4 *)(P + o0), own_pb = *(const uint4 *)(P + o0 + 8);
	const u32 ow[8] = {own_wa.x, own_wa.y, own_wa.z, own_wa.w, own_wb.x, own_wb.y, own_wb.z, own_wb.w};
	const u32 pl[8] = {own_pa.x, own_pa.y, own_pa.z, own_pa.w, own_pb.x, own_pb.y, own_pb.z, own_pb.w};
	uint4 *dst = (uint4 *)(match + ts + k0);        // ts is a multiple of TILE and k0 of RUN: 16-byte aligned; padded past n
#pragma unroll
	for (u32 q = 0; q < RUN / 4; q++) {
	u64 lwin = (u64)pl[2 * q] | (u64)pl[2 * q + 1] << 32;
	u32 r0 = 0, r1 = 0, r2 = 0, r3 = 0;
#pragma unroll 1
	for (u32 r = 0; r < 4; r++, lwin >>= 16) {
		const u32 j = 4 * q + r;
		const u32 k = k0 + j;
		if (k >= n_pos) break;
		const u32 rel = rel0 + j;                                  // p - cs
		const u32 o = o0 + j;
		// the position's first 12 bytes, from the run's registers
		const u32 cur = __funnelshift_r(ow[q], ow[q + 1], r * 8), own1 = __funnelshift_r(ow[q + 1], ow[q + 2], r * 8),
		          own2 = __funnelshift_r(ow[q + 2], ow[q + 3], r * 8);
		if (rel >= be_rel) be_rel += block_bytes;
		const int maxlen = (int)min((u32)MAX_MATCH, min(ce_rel, be_rel) - rel);   // runs stop at the block end (Lz77Huffman.java:75)
		int best_len = 0, best_dist = 0;
		if (mp.search == B2D_SEARCH_RLE) {
			if (rel > 0 && maxlen >= 3) {                          // distance 1 only (RLE_*, Lz77Huffman.java:301-302)
				int len = sm_match_len(W, o - 1, o, maxlen);
				if (len >= 3) { best_len = len; best_dist = 1; }
			}
		} else if (mp.search != B2D_SEARCH_LITERAL && maxlen >= mp.hb && rel + mp.hb <= ce_rel) {
			const u32 max_dist = min(WINDOW, rel);
			int depth = mp.depth;
			best_len = mp.hb - 1;
			if (inherit && prev_len > mp.hb) {                     // same source, shifted by one
				int len = min(prev_len - 1, maxlen);
				// (a match that ended at a mismatch cannot grow when it is shifted; only one cut at 258 can)
				if (len < maxlen && prev_len >= MAX_MATCH) len += sm_match_len(W, o - prev_dist + len, o + len, maxlen - len);
				best_len = len;
				best_dist = prev_dist;
				if (len >= 32) depth = len >= maxlen ? 0 : 1;       // already good: barely look further
				else if (len >= mp.good_len) depth = (depth + 3) >> 2;
			}
			// Positions 2 .. L-1 behind the start of a searched match of length L >= SKIP_MIN are only reached by
			// the parser when it arrives sideways; they keep what they inherit and are not searched.
			// (Position 1 is searched: the parser's lazy step compares it with the match before it.)
			if (inherit && since >= 2 && since < run_len) depth = 0;
			u32 d = (u32)lwin & 0xFFFFu;
			u32 dist = 0;
			u32 end_own = 0;                                       // own bytes at best_len - 3 .. best_len, read when best_len changes
			int end_for = -1;
			while (d != 0 && depth-- > 0) {
				dist += d;
				if (dist > max_dist) break;
				const u32 c = o - dist;
				d = P[c];
				if ((int)dist == best_dist) continue;
				// a candidate can only win if it also matches the byte that would make it longer; below 12 bytes the
				// comparison that follows says so by itself
				if (best_len >= 12) {
					if (end_for != best_len) { end_own = sm_load4(W, o + best_len - 3); end_for = best_len; }
					if (sm_load4(W, c + best_len - 3) != end_own) continue;
				}
				// the candidate's first 12 bytes against the position's, every lane alike: four aligned words, no loop
				const u32 ci = c >> 2, sc = (c & 3) * 8;
				const u32 c0 = W[ci], c1 = W[ci + 1], c2 = W[ci + 2], c3 = W[ci + 3];
				const u32 x0 = __funnelshift_r(c0, c1, sc) ^ cur;
				if (x0 & cmask) continue;                          // not even the hashed bytes
				const u32 x1 = __funnelshift_r(c1, c2, sc) ^ own1, x2 = __funnelshift_r(c2, c3, sc) ^ own2;
				int len = x0 ? 3 : x1 ? 4 + ((__ffs(x1) - 1) >> 3) : x2 ? 8 + ((__ffs(x2) - 1) >> 3) : 12;
				if (len >= 12 && maxlen > 12) len = 12 + sm_match_len(W, c + 12, o + 12, maxlen - 12);
				len = min(len, maxlen);
				if (len > best_len) {                              // strict: ties keep the smaller distance (:80)
					best_len = len;
					best_dist = (int)dist;
					if (len >= mp.nice || len >= maxlen) break;
				}
			}
			if (best_dist == 0) best_len = 0;
		}
		if (best_len >= SKIP_MIN && best_len > run_len - since) { run_len = best_len; since = 0; }
		since++;
		prev_len = best_len;
		prev_dist = best_dist;
		const u32 v = (cur << 24) | ((u32)best_len << 15) | (u32)(best_dist > 0 ? best_dist - 1 : 0);
		if (r == 0) r0 = v; else if (r == 1) r1 = v; else if (r == 2) r2 = v; else r3 = v;
	}
	dst[q] = make_uint4(r0, r1, r2, r3);
	}
	}
}

// ---------------------------------------------------------------- K2: parse + histograms
constexpr int PARSE_WARPS = 4;

__device__ __forceinline__ void hist_token(u32 *hist, u32 tok) {
	int len = (int)tok_len(tok);
	if (len == 0) { atomicAdd(&hist[tok >> 24], 1u); return; }
	int ne, ex;
	atomicAdd(&hist[len_symbol(len, ne, ex)], 1u);
	atomicAdd(&hist[288 + dist_symbol((int)tok_dist(tok), ne, ex)], 1u);
}

__global__ void __launch_bounds__(PARSE_WARPS * 32)
parse_kernel(const u32 *__restrict__ match, u64 n, u32 block_bytes, u32 n_blocks, int lazy,
             u32 *__restrict__ tokens, BlockRec *__restrict__ recs, Heap hp) {
	__shared__ u32 hist_sm[PARSE_WARPS][320];
	const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const u32 g = blockIdx.x * PARSE_WARPS + warp;
	if (g >= n_blocks) return;
	u32 *hist = hist_sm[warp];
	for (int i = lane; i < 320; i += 32) hist[i] = 0;
	__syncwarp();
	const u64 bs = (u64)g * block_bytes, be64 = min(n, bs + block_bytes);
	u32 *__restrict__ tk = tokens + bs;
	// 32-bit offsets relative to `org` (bs rounded down to a window of 32 positions)
	const u64 org = bs & ~(u64)31;
	const u32 *__restrict__ mb = match + org;
	const u32 lim = (u32)min((u64)0x7FFFFFFF, n - org);       // positions that exist, relative to org
	const u32 be = (u32)(be64 - org);
	u32 i = (u32)(bs - org);
	u32 base = 0;
#define LOADM(k) (base + (k) * 32 + lane < lim ? __ldg(mb + base + (k) * 32 + lane) : 0u)
	u32 w0 = LOADM(0), w1 = LOADM(1), w2 = LOADM(2), w3 = LOADM(3);    // four windows of 32 positions in flight
	u32 ntok = 0;
	// The parse is a chain -- the position after a token depends on the token -- but only its LINKS are serial: what a
	// position would emit if the parse came by (a literal, or its match unless the next position's is longer) is
	// known for all 32 positions of a window at once.  So per window: every lane works out its own step, the warp
	// follows the chain through the window with one shuffle per match (literal runs are skipped with a bit scan), and
	// the positions the chain visited write their tokens and count their symbols together.
	while (i < be) {
		u32 o = i - base;
		if (o >= 32) {
			if (o >= 128) {                                // a long match jumped past everything loaded
				base = i & ~31u;
				w0 = LOADM(0); w1 = LOADM(1); w2 = LOADM(2); w3 = LOADM(3);
			} else {
				do {
					base += 32;
					w0 = w1; w1 = w2; w2 = w3;
					w3 = LOADM(3);
				} while (i - base >= 32);
			}
			o = i - base;
		}
		const u32 len = tok_len(w0);
		u32 len_next = __shfl_down_sync(FULL_MASK, len, 1);
		const u32 len_w1 = __shfl_sync(FULL_MASK, tok_len(w1), 0);
		if (lane == 31) len_next = len_w1;
		const u32 p = base + lane;
		// greedy (Lz77Huffman.java:85-130), or defer to the next position when it matches longer (one-step lazy)
		const bool take = len != 0 && !(lazy && p + 1 < be && len_next > len);
		const u32 step = take ? len : 1u;
		const u32 takemask = __ballot_sync(FULL_MASK, take);
		const u32 wlim = min(32u, be - base);
		u32 visited = 0, cur = o;
		while (cur < wlim) {
			const u32 ahead = takemask >> cur;                 // literals up to the next position that takes a match
			const u32 run = min(ahead ? (u32)__ffs(ahead) - 1u : 32u, wlim - cur);
			visited |= (run >= 32 ? 0xFFFFFFFFu : (1u << run) - 1u) << cur;
			cur += run;
			if (cur >= wlim) break;
			visited |= 1u << cur;
			cur += __shfl_sync(FULL_MASK, step, cur);
		}
		i = base + cur;
		if ((visited >> lane) & 1u) {
			const u32 tok = take ? w0 : (w0 & 0xFF000000u);
			tk[ntok + __popc(visited & lanemask_lt())] = tok;
			hist_token(hist, tok);
		}
		ntok += __popc(visited);
	}
	__syncwarp();
	BlockRec *rec = &recs[leaf_slot(hp, g)];
	if (lane == 0) { hist[256] += 1; rec->n_tokens = ntok; }      // end-of-block (Lz77Huffman.java:131-132)
	__syncwarp();
	for (int k = lane; k < 288; k += 32) rec->hist_ll[k] = hist[k];
	rec->hist_d[lane] = hist[288 + lane];
}
#undef LOADM

// ---------------------------------------------------------------- K3: Huffman construction
constexpr int HUFF_WARPS = 2;

struct HuffSmem {
	u32 keys[512];
	u32 leaf_w[288];
	u32 pk_a[288];
	u32 pk_b[288];
	u32 merged[576];
	u32 leafmask[15][18];
	u32 hist[320];
	u32 cnt[16];
	u32 run[16];
	u32 first[16];
	u32 cl_hist[19];
	u32 cl_code[19];
	u32 level_leaves[15];
	u32 hdr[HDR_WORDS];
	u16 leaf_sym[288];
	u8 lens[320];
	u8 cl_len[19];
	u8 cl_sym[320];
	u8 cl_ext[320];
};

// calcHuffmanCodeLengths (Lz77Huffman.java:309-335): exact package-merge.  The reference's stable sort puts, among
// equal weights, packages before leaves and leaves in symbol order; the merged order is rebuilt here from ranks
// (leaf r sits after every package with weight <= its own; package j after every leaf with smaller weight).  A
// leaf's code length is the number of levels in which it lies inside the selected prefix, which equals the
// reference's countOccurrences over the first nLeaves-1 final packages.
template <int SORT_N>
__device__ void package_merge(HuffSmem *s, const u32 *hist, int n, int maxlen, u8 *out_lens, u32 lane) {
	for (int i = lane; i < n; i += 32) out_lens[i] = 0;
	int nl;
	if (SORT_N == 32) {
		u32 key = ((int)lane < n && hist[lane] > 0) ? (hist[lane] << 9 | lane) : 0xFFFFFFFFu;
		u32 rank = 0;
		for (int j = 0; j < 32; j++) rank += __shfl_sync(FULL_MASK, key, j) < key;
		nl = __popc(__ballot_sync(FULL_MASK, key != 0xFFFFFFFFu));
		if (key != 0xFFFFFFFFu) { s->leaf_w[rank] = key >> 9; s->leaf_sym[rank] = (u16)(key & 511); }
	} else {
		int cnt = 0;
		for (int i = lane; i < 512; i += 32) {
			u32 key = (i < n && hist[i] > 0) ? (hist[i] << 9 | (u32)i) : 0xFFFFFFFFu;
			s->keys[i] = key;
			cnt += key != 0xFFFFFFFFu;
		}
		nl = (int)warp_sum((u32)cnt);
		__syncwarp();
		for (int k = 2; k <= 512; k <<= 1) {           // bitonic sort, ascending
			for (int j = k >> 1; j > 0; j >>= 1) {
				for (int t = lane; t < 256; t += 32) {
					int a = ((t & ~(j - 1)) << 1) | (t & (j - 1));
					int b = a | j;
					u32 x = s->keys[a], y = s->keys[b];
					bool up = (a & k) == 0;
					if ((x > y) == up) { s->keys[a] = y; s->keys[b] = x; }
				}
				__syncwarp();
			}
		}
		for (int r = lane; r < nl; r += 32) { s->leaf_w[r] = s->keys[r] >> 9; s->leaf_sym[r] = (u16)(s->keys[r] & 511); }
	}
	__syncwarp();
	if (nl < 2) return;                                // the reference would return all zeros too
	for (int i = lane; i < 15 * 18; i += 32) (&s->leafmask[0][0])[i] = 0;
	__syncwarp();
	u32 *pk = s->pk_a, *pk_new = s->pk_b;
	int npk = 0;
	for (int lvl = 0; lvl < maxlen; lvl++) {
		for (int r = lane; r < nl; r += 32) {          // leaves: after packages with weight <= own
			u32 w = s->leaf_w[r];
			int lo = 0, hi = npk;
			while (lo < hi) { int mid = (lo + hi) >> 1; if (pk[mid] <= w) lo = mid + 1; else hi = mid; }
			int pos = r + lo;
			s->merged[pos] = w;
			atomicOr(&s->leafmask[lvl][pos >> 5], 1u << (pos & 31));
		}
		for (int j = lane; j < npk; j += 32) {         // packages: after leaves with smaller weight
			u32 w = pk[j];
			int lo = 0, hi = nl;
			while (lo < hi) { int mid = (lo + hi) >> 1; if (s->leaf_w[mid] < w) lo = mid + 1; else hi = mid; }
			s->merged[j + lo] = w;
		}
		__syncwarp();
		int nn = (npk + nl) >> 1;
		for (int j = lane; j < nn; j += 32) pk_new[j] = s->merged[2 * j] + s->merged[2 * j + 1];
		__syncwarp();
		u32 *t = pk; pk = pk_new; pk_new = t;
		npk = nn;
	}
	int m = 2 * (nl - 1);
	for (int lvl = maxlen - 1; lvl >= 0; lvl--) {      // selected prefix per level, top down
		u32 c = 0;
		if (lane < 18) {
			u32 w = s->leafmask[lvl][lane];
			int lo = (int)lane * 32;
			if (m >= lo + 32) c = __popc(w);
			else if (m > lo) c = __popc(w & ((1u << (m - lo)) - 1));
		}
		c = warp_sum(c);
		if (lane == 0) s->level_leaves[lvl] = c;
		m = 2 * (m - (int)c);
	}
	__syncwarp();
	for (int r = lane; r < nl; r += 32) {
		int len = 0;
		for (int lvl = 0; lvl < maxlen; lvl++) len += s->level_leaves[lvl] > (u32)r;
		out_lens[s->leaf_sym[r]] = (u8)len;
	}
	__syncwarp();
}

// codeLengthsToCodes (Lz77Huffman.java:372-391): canonical, bit-reversed, packed code << 4 | len
__device__ void assign_codes(HuffSmem *s, const u8 *lens, int n, u32 *codes, u32 lane) {
	if (lane < 16) { s->cnt[lane] = 0; s->run[lane] = 0; }
	__syncwarp();
	for (int i = lane; i < n; i += 32) if (lens[i]) atomicAdd(&s->cnt[lens[i]], 1u);
	__syncwarp();
	if (lane == 0) {
		u32 code = 0, prev = 0;
		for (int l = 1; l <= 15; l++) { code = (code + prev) << 1; s->first[l] = code; prev = s->cnt[l]; }
	}
	__syncwarp();
	for (int base = 0; base < n; base += 32) {
		int i = base + (int)lane;
		int l = i < n ? lens[i] : 0;
		u32 grp = __match_any_sync(FULL_MASK, l);
		u32 r = __popc(grp & lanemask_lt());
		u32 start = s->run[l];
		__syncwarp();
		if (r == 0) s->run[l] = start + __popc(grp);
		__syncwarp();
		if (i < n) codes[i] = l ? ((__brev(s->first[l] + start + r) >> (32 - l)) << 4 | (u32)l) : 0u;
	}
	__syncwarp();
}

struct BitW { u32 *w; u32 pos; };
__device__ __forceinline__ void bw_put(BitW &b, u32 v, int n) {      // single writer (lane 0), n <= 16
	if (n == 0) return;
	u32 i = b.pos >> 5, sh = b.pos & 31;
	b.w[i] |= v << sh;
	if (sh + n > 32) b.w[i + 1] |= v >> (32 - sh);
	b.pos += n;
}

__global__ void __launch_bounds__(HUFF_WARPS * 32)
huffman_kernel(BlockRec *__restrict__ recs, u32 n_items, u32 item_stride, u64 n, Heap hp) {
	__shared__ HuffSmem hs[HUFF_WARPS];
	const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const u32 item = blockIdx.x * HUFF_WARPS + warp;
	if (item >= n_items) return;
	const u32 g = item * item_stride;                   // stride H: only the roots of the spans
	HuffSmem *s = &hs[warp];
	BlockRec *rec = &recs[g];
	u64 bs, dl64;
	heap_range(hp, g / hp.H, g % hp.H, n, bs, dl64);
	const u32 data_len = (u32)dl64;
	if (data_len == 0 && n != 0 && hp.K > 1) return;                       // a node past the end of the input
	for (int i = lane; i < 288; i += 32) s->hist[i] = rec->hist_ll[i];
	s->hist[288 + lane] = rec->hist_d[lane];
	__syncwarp();

	// fixed-code cost (Lz77Huffman *_STATIC): 3 header bits + codes + extra bits
	u32 fixed = 0, extra_bits = 0;
	for (int i = lane; i < 286; i += 32) { fixed += s->hist[i] * fixed_ll_len(i); extra_bits += s->hist[i] * ll_extra_bits(i); }
	if (lane < 30) { fixed += s->hist[288 + lane] * 5; extra_bits += s->hist[288 + lane] * d_extra_bits(lane); }
	fixed = warp_sum(fixed);
	extra_bits = warp_sum(extra_bits);

	// histogram fix-ups (Lz77Huffman.java:145-181)
	if (lane == 0 && data_len == 0) s->hist[0] += 1;                       // :146-147 dummy literal
	__syncwarp();
	int n_ll = 286;
	{
		u32 nz = 0;                                                        // highest used lit/len symbol
		for (int i = lane; i < 286; i += 32) if (s->hist[i]) nz = i + 1;
		for (int o = 16; o > 0; o >>= 1) nz = max(nz, __shfl_xor_sync(FULL_MASK, nz, o));
		n_ll = max(257, (int)nz);                                          // :148-151
	}
	int n_d;
	bool no_dist;
	{
		u32 used = __ballot_sync(FULL_MASK, lane < 30 && s->hist[288 + lane] > 0);
		if (__popc(used) == 1) {                                           // :161-171 give a lone distance code a neighbour
			int i = __ffs(used) - 1;
			if (lane == 0) s->hist[288 + (i < 29 ? i + 1 : i - 1)] = 1;
			used |= 1u << (i < 29 ? i + 1 : i - 1);
		}
		__syncwarp();
		n_d = used ? 32 - __clz(used) : 1;                                 // :172-175
		no_dist = used == 0;                                               // :177-179
	}

	package_merge<512>(s, s->hist, n_ll, 15, s->lens, lane);                // :153
	if (no_dist) { if (lane == 0) s->lens[n_ll] = 0; }
	else package_merge<32>(s, s->hist + 288, n_d, 15, s->lens + n_ll, lane);   // :181
	__syncwarp();
	const int total = n_ll + n_d;

	// code-length RLE (Lz77Huffman.java:189-223), greedy and sequential by definition
	int ncl = 0;
	if (lane < 19) s->cl_hist[lane] = 0;
	__syncwarp();
	if (lane == 0) {
		const u8 *lens = s->lens;
		for (int i = 0; i < total;) {
			int val = lens[i];
			if (val == 0) {
				int rl = 1;
				for (; rl < 138 && i + rl < total && lens[i + rl] == 0; rl++);
				if (rl < 3) { s->cl_sym[ncl] = 0; s->cl_ext[ncl] = 0; i++; }
				else if (rl < 11) { s->cl_sym[ncl] = 17; s->cl_ext[ncl] = (u8)(rl - 3); i += rl; }
				else { s->cl_sym[ncl] = 18; s->cl_ext[ncl] = (u8)(rl - 11); i += rl; }
				ncl++;
				continue;
			}
			if (i > 0) {
				int rl = 0;
				for (; rl < 6 && i + rl < total && lens[i + rl] == lens[i - 1]; rl++);
				if (rl >= 3) { s->cl_sym[ncl] = 16; s->cl_ext[ncl] = (u8)(rl - 3); ncl++; i += rl; continue; }
			}
			s->cl_sym[ncl] = (u8)val; s->cl_ext[ncl] = 0; ncl++; i++;
		}
		for (int k = 0; k < ncl; k++) s->cl_hist[s->cl_sym[k]]++;
	}
	ncl = __shfl_sync(FULL_MASK, ncl, 0);
	__syncwarp();
	package_merge<32>(s, s->cl_hist, 19, 7, s->cl_len, lane);               // :228
	assign_codes(s, s->cl_len, 19, s->cl_code, lane);                       // :243

	// header bits (Lz77Huffman.java:230-258) and the dynamic block cost
	for (int i = lane; i < HDR_WORDS; i += 32) s->hdr[i] = 0;
	__syncwarp();
	u32 hdr_bits = 0;
	if (lane == 0) {
		int ncll = 19;
		for (; ncll > 4 && s->cl_len[CL_ORDER_D[ncll - 1]] == 0; ncll--);
		BitW bw{s->hdr, 0};
		bw_put(bw, (u32)(n_ll - 257), 5);
		bw_put(bw, (u32)(n_d - 1), 5);
		bw_put(bw, (u32)(ncll - 4), 4);
		for (int i = 0; i < ncll; i++) bw_put(bw, s->cl_len[CL_ORDER_D[i]], 3);
		for (int k = 0; k < ncl; k++) {
			int sym = s->cl_sym[k];
			u32 pair = s->cl_code[sym];
			bw_put(bw, pair >> 4, (int)(pair & 15));
			if (sym >= 16) bw_put(bw, s->cl_ext[k], sym == 16 ? 2 : sym == 17 ? 3 : 7);
		}
		hdr_bits = bw.pos;
	}
	hdr_bits = __shfl_sync(FULL_MASK, hdr_bits, 0);
	__syncwarp();
	u32 dyn = 0;                                       // emitted symbols only: the fix-up counts are not written
	for (int i = lane; i < n_ll; i += 32) dyn += rec->hist_ll[i] * s->lens[i];
	if ((int)lane < n_d && !no_dist) dyn += rec->hist_d[lane] * s->lens[n_ll + lane];
	dyn = warp_sum(dyn);

	// canonical codes for the emitter (:260-264)
	assign_codes(s, s->lens, n_ll, rec->code_ll, lane);
	if (!no_dist) assign_codes(s, s->lens + n_ll, n_d, rec->code_d, lane);
	for (int i = lane; i < HDR_WORDS; i += 32) rec->hdr[i] = s->hdr[i];
	if (lane == 0) {
		rec->hdr_bits = hdr_bits;
		rec->cost_fixed = 3 + fixed + extra_bits;
		rec->cost_dyn = 3 + hdr_bits + dyn + extra_bits;
	}
}

// ---------------------------------------------------------------- adaptive block splitting (comp/BinarySplit.java)
// The reference cuts a block in halves top-down and keeps a cut where the two halves, priced by the sub-strategy,
// are cheaper (BinarySplit.java:36-98).  Here every node of the span's binary tree gets its real codes and costs
// (huffman_kernel runs on all of them), and the cheapest partition is found bottom-up: at least as good as the
// top-down rule, which stops at the first level that does not pay.  The pieces were parsed once, at the leaf size;
// a block made of several pieces is emitted by its pieces' CTAs with the block's codes.

// K3a': histograms of the inner nodes = sums of their halves; one warp per span
__global__ void __launch_bounds__(128)
node_hist_kernel(BlockRec *__restrict__ recs, u32 n_roots, u64 n, Heap hp) {
	const u32 lane = threadIdx.x & 31;
	const u32 root = blockIdx.x * 4 + (threadIdx.x >> 5);
	if (root >= n_roots) return;
	BlockRec *base = recs + (size_t)root * hp.H;
	for (int h = (int)hp.K - 2; h >= 0; h--) {
		u64 s0, l0, s1, l1;
		heap_range(hp, root, 2 * h + 1, n, s0, l0);
		heap_range(hp, root, 2 * h + 2, n, s1, l1);
		if (l0 == 0) continue;                                              // the node lies past the end of the input
		BlockRec *a = base + 2 * h + 1, *b = base + 2 * h + 2, *o = base + h;
		for (int k = lane; k < 288; k += 32) o->hist_ll[k] = a->hist_ll[k] + (l1 ? b->hist_ll[k] : 0u);
		o->hist_d[lane] = a->hist_d[lane] + (l1 ? b->hist_d[lane] : 0u);
		__syncwarp();
		if (lane == 0) { o->hist_ll[256] = 1; o->n_tokens = a->n_tokens + (l1 ? b->n_tokens : 0u); }   // one end-of-block
		__syncwarp();
	}
}

__device__ __forceinline__ u64 stored_bits(u64 dl, int i) {                  // Uncompressed.java:23-25
	const u64 nsb = max((u64)1, (dl + 65534) / 65535);
	return (u64)((long long)(dl * 8 + nsb * 40) + (((13 - i) % 8) - 5));
}
__device__ __forceinline__ u64 single_cost(const BlockRec *r, u64 dl, int i, int mode, u32 &m) {
	const u64 sc = stored_bits(dl, i);
	if (mode == B2D_MODE_STORED) { m = 1; return sc; }
	if (mode == B2D_MODE_FIXED) { m = 2; return r->cost_fixed; }
	if (mode == B2D_MODE_DYNAMIC) { m = 3; return r->cost_dyn; }
	u64 best = sc;                                                          // first listed wins ties (MultiStrategy.java:39)
	m = 1;
	if (r->cost_fixed < best) { m = 2; best = r->cost_fixed; }
	if (r->cost_dyn < best) { m = 3; best = r->cost_dyn; }
	return best;
}

// K3a'': cheapest partition of every span, owners of the pieces, and the size of every piece's tokens under its
// block's codes; one warp per span, lane h = node h (at most 31 nodes)
__global__ void __launch_bounds__(128)
split_decide_kernel(BlockRec *__restrict__ recs, u32 n_roots, u64 n, Heap hp, int mode) {
	const u32 lane = threadIdx.x & 31;
	const u32 root = blockIdx.x * 4 + (threadIdx.x >> 5);
	if (root >= n_roots) return;
	BlockRec *base = recs + (size_t)root * hp.H;
	u64 st, dl = 0;
	if (lane < hp.H) heap_range(hp, root, lane, n, st, dl);
	const bool exists = lane < hp.H && dl > 0;
	const bool leaf = lane >= hp.K - 1;
	u32 m;
	const u64 single = exists ? single_cost(base + lane, dl, 0, mode, m) : 0;
	u64 best = single;
	bool split = false;
	for (u32 it = 0; it < 5; it++) {                                        // depth <= 4: values settle bottom-up
		const u64 bl = __shfl_sync(FULL_MASK, best, (2 * lane + 1) & 31), br = __shfl_sync(FULL_MASK, best, (2 * lane + 2) & 31);
		if (exists && !leaf && !hp.roots_only) {
			split = bl + br < single;                                       // a cut has to pay (BinarySplit.java:64)
			best = split ? bl + br : single;
		}
	}
	const u32 splitmask = __ballot_sync(FULL_MASK, split);
	if (exists) base[lane].best_cost = (u32)min(best, (u64)0xFFFFFFFFu);
	// pieces: lane j = piece j of the span
	u64 s0, span_len;
	heap_range(hp, root, 0, n, s0, span_len);
	const u32 n_exist = (u32)((span_len + hp.leaf_bytes - 1) / hp.leaf_bytes);
	u32 owner = 0, lo = 0, cnt = hp.K;
	while ((splitmask >> owner) & 1) {
		cnt >>= 1;
		if (lane < lo + cnt) owner = 2 * owner + 1;
		else { owner = 2 * owner + 2; lo += cnt; }
	}
	const u32 last_piece = min(lo + cnt, n_exist) - 1;
	if (lane < n_exist) {
		BlockRec *L = base + hp.K - 1 + lane;
		L->owner = owner;
		L->first_last = (lane == lo ? 1u : 0u) | (lane == last_piece ? 2u : 0u);
	}
	__syncwarp();
	for (u32 j = 0; j < n_exist; j++) {
		const u32 ow = __shfl_sync(FULL_MASK, owner, j);
		const BlockRec *O = base + ow;
		BlockRec *L = base + hp.K - 1 + j;
		u32 dyn = 0, fixed = 0;
		for (int k = lane; k < 286; k += 32) {
			const u32 c = L->hist_ll[k] - (k == 256 ? 1u : 0u);             // the piece's own end-of-block is not emitted
			dyn += c * ((O->code_ll[k] & 15) + ll_extra_bits(k));
			fixed += c * (fixed_ll_len(k) + ll_extra_bits(k));
		}
		if (lane < 30) {
			const u32 c = L->hist_d[lane];
			dyn += c * ((O->code_d[lane] & 15) + d_extra_bits(lane));
			fixed += c * (5 + d_extra_bits(lane));
		}
		dyn = warp_sum(dyn);
		fixed = warp_sum(fixed);
		if (lane == 0) { L->body_dyn = dyn; L->body_fixed = fixed; }
	}
}

// ---------------------------------------------------------------- K3b: per-chunk layout, K3c: scan
// Per block the cheapest of stored / fixed / dynamic for the block's actual start bit position, first listed
// wins ties (MultiStrategy.java:35-44,54); stored cost per Uncompressed.java:23-25.
__global__ void layout_kernel(BlockRec *__restrict__ recs, u32 n_blocks, u32 n_chunks, u64 n, u32 chunk_bytes,
                              u32 block_bytes, int mode, int is_last, int ref_framing, u64 *__restrict__ chunk_len,
                              u64 *__restrict__ chunk_tail, Heap hp) {
	u32 c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= n_chunks) return;
	const u32 bpc = chunk_bytes / block_bytes;          // block_bytes here = the parse unit (the leaf size when splitting)
	const u32 g0 = c * bpc;
	const u32 g1 = min(n_blocks, g0 + bpc);
	u64 bit = 0;
	if (hp.K == 1) {
		for (u32 g = g0; g < g1; g++) {
			BlockRec *r = &recs[g];
			const u64 bs = (u64)g * block_bytes;
			const u64 dl = min(n, bs + block_bytes) - min(n, bs);
			u32 m;
			const u64 best = single_cost(r, dl, (int)(bit & 7), mode, m);
			r->mode = m;
			r->out_bit = bit;
			r->bfinal = (ref_framing && is_last && g + 1 == n_blocks) ? 1u : 0u;
			bit += best;
		}
	} else {
		u64 blk_start = 0, run = 0, blk_stored = 0;
		u32 m = 0, eob = 0;
		for (u32 g = g0; g < g1; g++) {
			BlockRec *L = &recs[leaf_slot(hp, g)];
			const u32 root = g / hp.K;
			BlockRec *O = &recs[(size_t)root * hp.H + L->owner];
			if (L->first_last & 1) {
				u64 os, ol;
				heap_range(hp, root, L->owner, n, os, ol);
				(void)single_cost(O, ol, (int)(bit & 7), mode, m);
				blk_stored = stored_bits(ol, (int)(bit & 7));
				const u32 pieces = hp.K >> heap_level(L->owner);
				O->mode = m;
				O->out_bit = bit;
				O->bfinal = (ref_framing && is_last && min(g + pieces, n_blocks) == n_blocks) ? 1u : 0u;
				blk_start = bit;
				run = bit + 3 + (m == 3 ? O->hdr_bits : 0u);
				eob = m == 3 ? (O->code_ll[256] & 15) : 7u;
			}
			L->tok_bit = run;
			run += m == 3 ? L->body_dyn : m == 2 ? L->body_fixed : 0u;
			if (L->first_last & 2) bit = m == 1 ? blk_start + blk_stored : run + eob;
		}
	}
	chunk_tail[c] = bit;
	if (!ref_framing) {
		// empty stored block closes the chunk byte-aligned: 3 header bits, pad, 00 00 FF FF
		bit += 3;
		bit = (bit + 7) & ~(u64)7;
		bit += 32;
	} else {
		bit = (bit + 7) & ~(u64)7;                   // BitOut.finish pads with zeros (DeflaterOutputStream.java:164-169)
	}
	chunk_len[c] = bit >> 3;
}

__global__ void __launch_bounds__(1024)
scan_kernel(const u64 *__restrict__ chunk_len, u32 n_chunks, u64 *__restrict__ chunk_off, u64 *__restrict__ total_out,
            u64 *__restrict__ user_chunk_len) {
	__shared__ u64 warp_tot[32];
	__shared__ u64 carry_sh;
	const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) carry_sh = 0;
	__syncthreads();
	for (u32 base = 0; base < n_chunks; base += 1024) {
		u32 i = base + threadIdx.x;
		u64 v = i < n_chunks ? chunk_len[i] : 0;
		if (user_chunk_len && i < n_chunks) user_chunk_len[i] = v;
		u64 x = v;
		for (int o = 1; o < 32; o <<= 1) { u64 y = __shfl_up_sync(FULL_MASK, x, o); if ((int)lane >= o) x += y; }
		if (lane == 31) warp_tot[warp] = x;
		__syncthreads();
		u64 wsum = 0;
		for (u32 w = 0; w < warp; w++) wsum += warp_tot[w];
		u64 carry = carry_sh;
		if (i < n_chunks) chunk_off[i] = carry + wsum + x - v;
		__syncthreads();
		if (threadIdx.x == 1023) carry_sh = carry + wsum + x;
		__syncthreads();
	}
	if (threadIdx.x == 0) { chunk_off[n_chunks] = carry_sh; *total_out = carry_sh; }
}

// ---------------------------------------------------------------- K4: bit emission
constexpr int EMIT_THREADS = 256;
constexpr int EMIT_STAGE_WORDS = EMIT_THREADS * 48 / 32 + 4;

__device__ __forceinline__ void or_bits_global(u32 *out_words, u64 bitpos, u64 v, int nbits) {
	if (nbits == 0) return;
	u64 wi = bitpos >> 5;
	u32 sh = (u32)(bitpos & 31);
	atomicOr(out_words + wi, (u32)(v << sh));
	if (sh + nbits > 32) {
		atomicOr(out_words + wi + 1, (u32)(v >> (32 - sh)));
		if (sh + nbits > 64) atomicOr(out_words + wi + 2, (u32)(v >> (64 - sh)));
	}
}
__device__ __forceinline__ void or_bits_shared(u32 *w, u32 bitpos, u64 v, int nbits) {
	if (nbits == 0) return;
	u32 wi = bitpos >> 5, sh = bitpos & 31;
	atomicOr(w + wi, (u32)(v << sh));
	if (sh + nbits > 32) {
		atomicOr(w + wi + 1, (u32)(v >> (32 - sh)));
		if (sh + nbits > 64) atomicOr(w + wi + 2, (u32)(v >> (64 - sh)));
	}
}

__global__ void __launch_bounds__(EMIT_THREADS)
emit_kernel(const u8 *__restrict__ in, u64 n, u32 chunk_bytes, u32 block_bytes, u32 n_blocks, u32 n_chunks,
            const BlockRec *__restrict__ recs, const u32 *__restrict__ tokens, const u64 *__restrict__ chunk_off,
            const u64 *__restrict__ chunk_tail, int is_last, int ref_framing, u8 *out, Heap hp) {
	__shared__ u32 code_ll[288];
	__shared__ u32 code_d[32];
	__shared__ u32 stage[EMIT_STAGE_WORDS];
	__shared__ u32 warp_tot[EMIT_THREADS / 32];
	// One CTA per parse unit (block_bytes here: the leaf size when splitting).  `r` is the record of the DEFLATE block
	// the unit belongs to -- its own without splitting -- and `piece` the unit's own record: the block's first unit
	// writes the block header (a stored block: everything), every unit its own tokens with the block's codes, the
	// last one the end-of-block symbol.
	const u32 g = blockIdx.x;
	const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const BlockRec *piece = &recs[leaf_slot(hp, g)];
	const BlockRec *r = hp.K == 1 ? piece : &recs[(size_t)(g / hp.K) * hp.H + piece->owner];
	const bool first = hp.K == 1 || (piece->first_last & 1), last = hp.K == 1 || (piece->first_last & 2);
	const u32 c = g / (chunk_bytes / block_bytes);
	u64 bs = (u64)g * block_bytes;
	u64 dl = min(n, bs + block_bytes) - bs;
	u32 *outw = (u32 *)out;
	u64 bit = chunk_off[c] * 8 + r->out_bit;
	const u32 mode = r->mode;

	// the last block of a chunk also writes the chunk trailer: empty stored block, BFINAL only at the stream end
	const bool chunk_tail_blk = (g + 1 == n_blocks) || ((g + 1) % (chunk_bytes / block_bytes) == 0);
	if (chunk_tail_blk && !ref_framing && tid == 0) {
		const u64 tb = chunk_off[c] * 8 + chunk_tail[c];
		const bool fin = is_last && c + 1 == n_chunks;
		if (fin) or_bits_global(outw, tb, 1u, 1);                          // BFINAL; BTYPE = 00 is already zero
		or_bits_global(outw, ((tb + 3 + 7) & ~(u64)7) + 16, 0xFFFFu, 16);   // LEN = 0000, NLEN = FFFF
	}

	if (mode == 1) {                                   // stored (Uncompressed.java:35-45)
		if (!first) return;
		if (hp.K > 1) heap_range(hp, g / hp.K, piece->owner, n, bs, dl);   // the whole block's bytes
		u64 pos = bs, end = bs + dl;
		do {
			u64 nb = min((u64)65535, end - pos);
			bool fin = r->bfinal && nb == end - pos;
			if (tid == 0) or_bits_global(outw, bit, fin ? 1u : 0u, 3);
			bit = (bit + 3 + 7) & ~(u64)7;
			if (tid == 0) or_bits_global(outw, bit, (u32)nb | ((u32)nb ^ 0xFFFFu) << 16, 32);
			bit += 32;
			// payload: whole 32-bit words with plain stores; the (up to 3) bytes before the first and after the last
			// whole word share their word with neighbouring blocks' bits, so they are OR-ed in atomically
			const u64 B = bit >> 3;
			const u8 *src = in + pos;
			const u32 head = (u32)min((u64)((4 - (B & 3)) & 3), nb);
			const u64 nw = (nb - head) >> 2;
			const u32 tail = (u32)(nb - head - (nw << 2));
			if (tid < head) atomicOr(outw + (B >> 2), (u32)src[tid] << (((B & 3) + tid) * 8));
			u32 *dw = outw + ((B + head) >> 2);
			const u8 *s4 = src + head;
			if ((((uintptr_t)s4) & 3) == 0) {
				const u32 *sw = (const u32 *)s4;
				for (u64 k = tid; k < nw; k += EMIT_THREADS) dw[k] = __ldg(sw + k);
			} else {
				const u32 *sw = (const u32 *)((uintptr_t)s4 & ~(uintptr_t)3);
				const u32 sh = (u32)((uintptr_t)s4 & 3) * 8;
				for (u64 k = tid; k < nw; k += EMIT_THREADS) dw[k] = __funnelshift_r(__ldg(sw + k), __ldg(sw + k + 1), sh);
			}
			if (tid < tail) atomicOr(dw + nw, (u32)s4[(nw << 2) + tid] << (tid * 8));
			bit += nb * 8;
			pos += nb;
		} while (pos < end);
		return;
	}

	for (int i = tid; i < 288; i += EMIT_THREADS) code_ll[i] = mode == 3 ? r->code_ll[i]
		: ((__brev((u32)(i < 144 ? 0x30 + i : i < 256 ? 0x190 + i - 144 : i < 280 ? i - 256 : 0xC0 + i - 280)) >> (32 - fixed_ll_len(i))) << 4 | (u32)fixed_ll_len(i));
	if (tid < 32) code_d[tid] = mode == 3 ? r->code_d[tid] : ((__brev(tid) >> 27) << 4 | 5u);
	for (int i = tid; i < EMIT_STAGE_WORDS; i += EMIT_THREADS) stage[i] = 0;
	if (first && tid == 0) or_bits_global(outw, bit, (r->bfinal ? 1u : 0u) | (mode == 3 ? 2u : 1u) << 1, 3);
	bit += 3;
	if (mode == 3) {                                   // dynamic header
		const u32 hb = r->hdr_bits;
		if (first)
			for (u32 w = tid; w * 32 < hb; w += EMIT_THREADS) {
				int nb = (int)min(32u, hb - w * 32);
				or_bits_global(outw, bit + (u64)w * 32, r->hdr[w], nb);
			}
		bit += hb;
	}
	if (hp.K > 1) bit = chunk_off[c] * 8 + piece->tok_bit;    // this piece's tokens start where layout_kernel put them
	__syncthreads();

	const u32 ntok = piece->n_tokens + (last ? 1u : 0u);      // + end-of-block
	const u32 *tk = tokens + bs;
	for (u32 base = 0; base < ntok; base += EMIT_THREADS) {
		const u32 t = base + tid;
		u64 bits = 0;
		int nb = 0;
		if (t < ntok) {
			if (last && t + 1 == ntok) { u32 p = code_ll[256]; bits = p >> 4; nb = p & 15; }
			else {
				u32 e = tk[t];
				int len = (int)tok_len(e);
				if (len == 0) { u32 p = code_ll[e >> 24]; bits = p >> 4; nb = p & 15; }
				else {
					int ne, ex;
					int sym = len_symbol(len, ne, ex);
					u32 p = code_ll[sym];
					bits = p >> 4; nb = p & 15;
					bits |= (u64)ex << nb; nb += ne;
					int dsym = dist_symbol((int)tok_dist(e), ne, ex);
					u32 q = code_d[dsym];
					bits |= (u64)(q >> 4) << nb; nb += q & 15;
					bits |= (u64)ex << nb; nb += ne;
				}
			}
		}
		// exclusive prefix sum of nb over the CTA
		u32 x = (u32)nb;
		for (int o = 1; o < 32; o <<= 1) { u32 y = __shfl_up_sync(FULL_MASK, x, o); if ((int)lane >= o) x += y; }
		if (lane == 31) warp_tot[warp] = x;
		__syncthreads();
		u32 woff = 0, tile_bits = 0;
		for (int w = 0; w < EMIT_THREADS / 32; w++) { u32 v = warp_tot[w]; if (w < (int)warp) woff += v; tile_bits += v; }
		const u32 sh0 = (u32)(bit & 31);
		or_bits_shared(stage, sh0 + woff + x - (u32)nb, bits, nb);
		__syncthreads();
		const u32 nwords = (sh0 + tile_bits + 31) >> 5;
		u32 *dstw = outw + (bit >> 5);
		for (u32 w = tid; w < nwords; w += EMIT_THREADS) {
			u32 v = stage[w];
			stage[w] = 0;
			if (w == 0 || w + 1 == nwords) { if (v) atomicOr(dstw + w, v); }    // words shared with neighbours
			else dstw[w] = v;
		}
		bit += tile_bits;
		__syncthreads();
	}
}

// ---------------------------------------------------------------- block index for the block-parallel decoder
__global__ void block_bits_kernel(const BlockRec *__restrict__ recs, u32 n_roots, u32 *__restrict__ out, Heap hp) {
	const u32 g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= n_roots) return;
	// bit offset, inside its chunk's output, of the (first) block of every block_bytes span
	const BlockRec *base = recs + (size_t)g * hp.H;
	out[g] = (u32)(hp.K == 1 ? base->out_bit : base[base[hp.K - 1].owner].out_bit);
}

// ---------------------------------------------------------------- host side
// Worst case over all modes: a block never costs more than 9 bits per byte (the fixed code is always available to
// the optimiser; a flat 9-bit code bounds the length-limited dynamic code) plus 3 header bits and < 384 bytes of
// dynamic header; stored pieces add 5 bytes per 65535; every chunk adds its 5-byte marker plus padding.
uint64_t deflate_bound_bytes(uint64_t in_len, uint32_t chunk_bytes, uint32_t block_bytes) {
	(void)block_bytes;                                    // sized for the smallest accepted block (4 KiB)
	uint64_t n_blocks = in_len / 4096 + 1;
	uint64_t n_chunks = (in_len + chunk_bytes - 1) / chunk_bytes + 1;
	return in_len + in_len / 8 + n_blocks * 384 + n_chunks * 16 + 1024;
}
// The parse unit.  With adaptive splitting it is the piece the caller asked for.  Without, the default search on
// chunked framing still parses in pieces of 16 KiB and emits every block_bytes span as one block with the summed
// histogram: parse_kernel is one warp per unit, and 64 KiB units are only 1.7 waves of warps per GiB (8.9 ms against
// 6.4 ms for the same bytes in 16 KiB units); the price is a match cut every 16 KiB (< 0.05 % of the output).  The
// reference framing and the reference's own searches keep the block as the unit: their bytes are compared with the
// restated reference encoder's.
static Heap heap_for(const DeflateParams &p, u64 n) {
	Heap hp;
	hp.block_bytes = p.block_bytes;
	hp.leaf_bytes = p.block_bytes;
	hp.roots_only = 0;
	if (p.leaf_bytes && n) hp.leaf_bytes = p.leaf_bytes;
	else if (n && p.framing == 0 && p.search == B2D_SEARCH_DEFAULT && p.mode != B2D_MODE_STORED && p.block_bytes >= 65536 &&
	         p.block_bytes <= 262144 && (p.block_bytes & (p.block_bytes - 1)) == 0) {
		hp.leaf_bytes = 16384;
		hp.roots_only = 1;
	}
	hp.K = hp.block_bytes / hp.leaf_bytes;
	hp.H = 2 * hp.K - 1;
	return hp;
}

size_t deflate_scratch_bytes(uint64_t in_len, const DeflateParams &p) {
	u64 n_blocks = (in_len + p.block_bytes - 1) / p.block_bytes;
	n_blocks *= heap_for(p, in_len ? in_len : 1).H;                             // the spans' binary trees
	u64 n_chunks = (in_len + p.chunk_bytes - 1) / p.chunk_bytes;
	size_t s = 0;
	s += ((in_len * 2 + 255) & ~(u64)255) + 256;          // prevdist u16
	s += ((in_len * 4 + 255) & ~(u64)255) + 256;          // match u32
	s += ((in_len * 4 + 255) & ~(u64)255) + 256;          // tokens u32
	s += ((n_blocks * sizeof(BlockRec) + 255) & ~(u64)255) + 256;
	s += ((n_chunks + 2) * 8 + 255) & ~(u64)255;          // chunk_len
	s += ((n_chunks + 2) * 8 + 255) & ~(u64)255;          // chunk_off
	s += ((n_chunks + 2) * 8 + 255) & ~(u64)255;          // chunk_tail
	return s + sizeof(BlockRec) + 2048;
}

static bool g_attr_set[MAX_DEVICES] = {};

cudaError_t launch_deflate(const uint8_t *d_in, uint64_t n, const DeflateParams &p, uint8_t *d_out,
                           uint64_t out_cap, uint64_t *d_out_len_total, uint64_t *d_chunk_out_len,
                           void *d_scratch, size_t scratch_bytes, cudaStream_t st, uint32_t *d_block_bits,
                           const DeflateAux *aux) {
	cudaError_t e;
	const int slot = current_device_slot();
	if (!g_attr_set[slot]) {
		e = cudaFuncSetAttribute(chains_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 << HASH_BITS);
		if (e != cudaSuccess) return e;
		e = cudaFuncSetAttribute(match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MATCH_SMEM);
		if (e != cudaSuccess) return e;
		g_attr_set[slot] = true;
	}
	const int ref_framing = p.framing == 1;
	const int is_last = p.is_last;
	// parse unit: the block, or the piece of adaptive splitting (then a block_bytes span owns a heap of records)
	const Heap hp = heap_for(p, n);
	const u32 unit = hp.leaf_bytes;
	u32 n_blocks = (u32)((n + unit - 1) / unit);                 // parse units
	u32 n_roots = (u32)((n + p.block_bytes - 1) / p.block_bytes);
	u32 n_chunks = (u32)((n + p.chunk_bytes - 1) / p.chunk_bytes);
	if (ref_framing && n == 0) { n_blocks = 1; n_roots = 1; n_chunks = 1; }   // the reference writes one (empty) final block
	const u32 n_recs = n_roots * hp.H;
	// carve scratch
	u8 *sp = (u8 *)d_scratch;
	auto carve = [&](size_t bytes) { u8 *r = sp; sp += (bytes + 255) & ~(size_t)255; return r; };
	u16 *prevdist = (u16 *)carve(n * 2 + 256);
	u32 *match = (u32 *)carve(n * 4 + 256);
	u32 *tokens = (u32 *)carve(n * 4 + 256);
	BlockRec *recs = (BlockRec *)carve((size_t)n_recs * sizeof(BlockRec) + 256);
	u64 *chunk_len = (u64 *)carve((size_t)(n_chunks + 2) * 8);
	u64 *chunk_off = (u64 *)carve((size_t)(n_chunks + 2) * 8);
	u64 *chunk_tail = (u64 *)carve((size_t)(n_chunks + 2) * 8);
	if ((size_t)(sp - (u8 *)d_scratch) > scratch_bytes) return cudaErrorInvalidValue;

	e = cudaMemsetAsync(d_out, 0, out_cap, st);
	if (e != cudaSuccess) return e;
	if (n_blocks == 0) {
		// empty input: only the closing empty stored block (if this call ends the stream)
		static const u8 fin[5] = {0x01, 0x00, 0x00, 0xFF, 0xFF};
		u64 len = is_last ? 5 : 0;
		if (is_last) { e = cudaMemcpyAsync(d_out, fin, 5, cudaMemcpyHostToDevice, st); if (e != cudaSuccess) return e; }
		return cudaMemcpyAsync(d_out_len_total, &len, 8, cudaMemcpyHostToDevice, st);
	}
	const bool need_search = p.mode != B2D_MODE_STORED;
	MatchParams mp;
	mp.search = p.search;
	mp.hb = p.search == 3 ? 3 : 4;
	mp.depth = p.search == 3 ? 0x7FFFFFFF : p.depth;
	mp.nice = MAX_MATCH;
	mp.skip_min = 6;
	mp.good_len = 1 << 20;
	{   // diagnostics for tuning the default search (they change the parse, never the validity of the stream)
		static const char *e_skip = getenv("B2D_MATCH_SKIP_MIN"), *e_good = getenv("B2D_MATCH_GOOD");
		if (e_skip && atoi(e_skip) >= 3) mp.skip_min = atoi(e_skip);
		if (e_good && atoi(e_good) >= 4) mp.good_len = atoi(e_good);
	}
	if (need_search) {
		// The search stage of chunks [c0, c1) on stream S: links, matches, parse + histograms, codes.  Everything in it is
		// indexed from the start of the range it is given, so a part of the input is just a smaller input.
		auto search_part = [&](u64 a, u64 nk, u32 units_k, u32 roots_k, u32 chunks_k, cudaStream_t S, cudaEvent_t match_done) {
			BlockRec *recs_k = recs + (size_t)(a / p.block_bytes) * hp.H;
			if (nk && (p.search == B2D_SEARCH_DEFAULT || p.search == 3)) {
				const u32 n_segs = chunks_k * ((p.chunk_bytes + CHAIN_SEG - 1) / CHAIN_SEG);
				B2D_LAUNCH(chains_kernel, n_segs, 32, 2 << HASH_BITS, S)(d_in + a, nk, p.chunk_bytes, CHAIN_SEG, mp.hb, prevdist + a);
			}
			const u32 n_tiles = (u32)((nk + TILE - 1) / TILE);
			if (n_tiles) B2D_LAUNCH(match_kernel, n_tiles, MATCH_THREADS, MATCH_SMEM, S)(d_in + a, nk, p.chunk_bytes, unit, mp, prevdist + a, match + a);
			if (match_done) cudaEventRecord(match_done, S);
			B2D_LAUNCH(parse_kernel, (units_k + PARSE_WARPS - 1) / PARSE_WARPS, PARSE_WARPS * 32, 0, S)(
				match + a, nk, unit, units_k, p.lazy, tokens + a, recs_k, hp);
			if (hp.K > 1) B2D_LAUNCH(node_hist_kernel, (roots_k + 3) / 4, 128, 0, S)(recs_k, roots_k, nk, hp);
			if (hp.roots_only) B2D_LAUNCH(huffman_kernel, (roots_k + HUFF_WARPS - 1) / HUFF_WARPS, HUFF_WARPS * 32, 0, S)(recs_k, roots_k, hp.H, nk, hp);
			else B2D_LAUNCH(huffman_kernel, (roots_k * hp.H + HUFF_WARPS - 1) / HUFF_WARPS, HUFF_WARPS * 32, 0, S)(recs_k, roots_k * hp.H, 1, nk, hp);
		};
		// Parts (diagnostic, B2D_DEFLATE_PARTS=k): the search stage can run in k parts on two streams, part i + 1's links
		// being built under part i's parse and Huffman kernels (not under its match_kernel: that one needs 197 KB of an
		// SM's shared memory, and chains CTAs that got there first -- 16 KB and three milliseconds each -- keep it off
		// the SM).  Measured on 1 and 2 GiB: no gain (31.0 against 28.5 ms per GiB with two wave-sized parts, far worse
		// with smaller ones: a chains segment takes its ~3 ms however few of them there are), so one part is the default.
		int n_parts = 1;
		{
			static const char *e_parts = getenv("B2D_DEFLATE_PARTS");
			if (e_parts && aux && p.framing == 0 && p.search == B2D_SEARCH_DEFAULT && atoi(e_parts) > 1 && n_chunks >= (u32)atoi(e_parts))
				n_parts = min(64, atoi(e_parts));
		}
		if (n_parts == 1) {
			search_part(0, n, n_blocks, n_roots, n_chunks, st, nullptr);
		} else {
			if ((e = cudaEventRecord(aux->ev[0], st)) != cudaSuccess) return e;
			if ((e = cudaStreamWaitEvent(aux->stream, aux->ev[0], 0)) != cudaSuccess) return e;
			for (int k = 0; k < n_parts; k++) {
				cudaStream_t S = (k & 1) ? aux->stream : st;
				if (k > 0 && (e = cudaStreamWaitEvent(S, aux->ev[1 + ((k - 1) & 3)], 0)) != cudaSuccess) return e;     // match(k - 1) done
				const u32 c0 = (u32)((u64)n_chunks * k / n_parts), c1 = (u32)((u64)n_chunks * (k + 1) / n_parts);
				const u64 a = (u64)c0 * p.chunk_bytes, nk = min(n, (u64)c1 * p.chunk_bytes) - a;
				search_part(a, nk, (u32)((nk + unit - 1) / unit), (u32)((nk + p.block_bytes - 1) / p.block_bytes), c1 - c0, S,
				            k + 1 < n_parts ? aux->ev[1 + (k & 3)] : nullptr);
			}
			if ((e = cudaEventRecord(aux->ev[5], aux->stream)) != cudaSuccess) return e;
			if ((e = cudaStreamWaitEvent(st, aux->ev[5], 0)) != cudaSuccess) return e;
		}
	}
	if (hp.K > 1) B2D_LAUNCH(split_decide_kernel, (n_roots + 3) / 4, 128, 0, st)(recs, n_roots, n, hp, p.mode);
	B2D_LAUNCH(layout_kernel, (n_chunks + 127) / 128, 128, 0, st)(recs, n_blocks, n_chunks, n, p.chunk_bytes, unit,
	                                                     p.mode, is_last, ref_framing, chunk_len, chunk_tail, hp);
	B2D_LAUNCH(scan_kernel, 1, 1024, 0, st)(chunk_len, n_chunks, chunk_off, d_out_len_total, d_chunk_out_len);
	B2D_LAUNCH(emit_kernel, n_blocks, EMIT_THREADS, 0, st)(d_in, n, p.chunk_bytes, unit, n_blocks, n_chunks, recs,
	                                               tokens, chunk_off, chunk_tail, is_last, ref_framing, d_out, hp);
	if (d_block_bits) B2D_LAUNCH(block_bits_kernel, (n_roots + 255) / 256, 256, 0, st)(recs, n_roots, d_block_bits, hp);
	return cudaGetLastError();
}

}  // namespace b2d
