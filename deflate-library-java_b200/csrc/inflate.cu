// inflate.cu -- batched DEFLATE decoder for sm_100a: one warp per independent member / chunk.
//
// Replaces the decode loops of the reference (paths relative to src/io/nayuki/deflate/):
//   decomp/Open.java:83-110   block loop            -> inflate_member()
//   decomp/Open.java:137-170  bit reader            -> BitIn (64-bit buffer, 32-bit aligned refills, prefetch)
//   decomp/Open.java:227-306  stored block          -> stored_block()  (warp-wide coalesced copy)
//   decomp/Open.java:336-431  dynamic header        -> dynamic_header()
//   decomp/Open.java:705-789  code tree + 9-bit LUT -> build_code(): canonical codes built by the whole warp
//                                                      into a 10-bit (lit/len) / 8-bit (distance) LUT in shared
//                                                      memory, longer codes resolved canonically (no tree walk)
//   decomp/Open.java:438-620  symbol loop + copy    -> decode_tokens(): every lane decodes the same symbol from
//                                                      shared tables (no divergence, no broadcast needed), then the
//                                                      32 lanes copy the back-reference together
// Results (bytes, out_len, consumed input, status) are identical to the reference's; the validation ORDER of
// Open.java is kept (first failing check wins).  Not a translation: no dictionary ring (the output buffer is
// the window), no code tree, no per-block allocation.
#include "common.cuh"
#include "kernels.h"

namespace b2d {

constexpr int LL_TB = 10;                 // lit/len LUT index bits
constexpr int D_TB = 8;                   // distance LUT index bits
constexpr int WARPS_PER_CTA = 4;

// LUT entry: [4:0] total bits (code + extra)  [8:5] code length  [12:9] extra-bit count
//            [31:27] flags (lit/len)  or bit 31 (distance)        [26:16] / [30:16] value
constexpr u32 F_LIT = 1u << 27;
constexpr u32 F_LEN = 1u << 28;
constexpr u32 F_EOB = 1u << 29;
constexpr u32 F_LONG = 1u << 30;          // code longer than the LUT index: canonical slow path
constexpr u32 F_RSVD = 1u << 31;          // lit/len symbols 286/287
constexpr u32 FD_SPECIAL = 1u << 31;      // distance: value 0 = long code, else reserved symbol 30/31

__constant__ u8 CL_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct Canon {                            // canonical-code description for the slow path
	u32 cnt[16];
	u32 run[16];
	u16 first[16];
	u16 offs[16];
};

struct WarpSmem {
	u32 ll_lut[1 << LL_TB];
	u32 d_lut[1 << D_TB];
	Canon ll_canon, d_canon;
	u16 ll_sorted[288];
	u16 d_sorted[32];
	u16 cl_lut[128];
	u8 lens[320];
};

struct BitIn {
	const u32 *words;     // 4-byte aligned base at or before the member's first byte
	u32 n_safe;           // words that may be read (cover the last real byte)
	u32 n_full;           // words that hold only real bytes
	u32 lead8;            // bits to skip in word 0
	u64 total_bits;       // real bits in the member
	u64 buf;
	int cnt;              // bits in buf (may include bits past the end of input; see avail())
	u32 next;             // prefetched word `widx`
	u32 widx;
};

__device__ __forceinline__ u32 load_word(const BitIn &b, u32 i) {
	return i < b.n_safe ? __ldg(b.words + i) : 0u;
}
__device__ __forceinline__ void bit_seek(BitIn &b, u64 byte_pos) {     // byte_pos relative to member start
	u64 a = (b.lead8 >> 3) + byte_pos;
	u32 w = (u32)(a >> 2);
	u32 sh = (u32)(a & 3) * 8;
	b.buf = load_word(b, w) >> sh;
	b.cnt = 32 - (int)sh;
	b.widx = w + 1;
	b.next = load_word(b, b.widx);
}
__device__ __forceinline__ void refill(BitIn &b) {                     // afterwards cnt >= 32
	if (b.cnt < 32) {
		b.buf |= (u64)b.next << b.cnt;
		b.cnt += 32;
		b.widx++;
		b.next = load_word(b, b.widx);
	}
}
__device__ __forceinline__ u64 consumed_bits(const BitIn &b) {
	return (u64)b.widx * 32 - b.lead8 - (u64)b.cnt;
}
__device__ __forceinline__ int avail_bits(const BitIn &b) {            // real bits left, clamped to int
	u64 c = consumed_bits(b);
	if (c >= b.total_bits) return 0;
	u64 a = b.total_bits - c;
	return a > 0x3FFFFFFFull ? 0x3FFFFFFF : (int)a;
}
__device__ __forceinline__ void drop(BitIn &b, int n) { b.buf >>= n; b.cnt -= n; }

// checked read for headers (Open.readBits, Open.java:137-170): n <= 16
__device__ __forceinline__ int getbits(BitIn &b, int n, int &avail, int &err) {
	refill(b);
	if (n > avail) { err = B2D_UNEXPECTED_END_OF_STREAM; return 0; }
	u32 v = (u32)b.buf & ((1u << n) - 1);
	drop(b, n);
	avail -= n;
	return (int)v;
}

__device__ __forceinline__ u32 ll_entry(int sym, int l) {
	if (sym < 256) return F_LIT | (u32)sym << 16 | (u32)l << 5 | (u32)l;
	if (sym == 256) return F_EOB | (u32)l << 5 | (u32)l;
	if (sym > 285) return F_RSVD | (u32)sym << 16 | (u32)l << 5 | (u32)l;      // Open.java:513-517
	int base, eb;
	length_sym_info(sym, base, eb);
	return F_LEN | (u32)base << 16 | (u32)eb << 9 | (u32)l << 5 | (u32)(l + eb);
}
__device__ __forceinline__ u32 d_entry(int sym, int l) {
	if (sym > 29) return FD_SPECIAL | (u32)sym << 16 | (u32)l << 5 | (u32)l;    // Open.java:546-551
	int base, eb;
	dist_sym_info(sym, base, eb);
	return (u32)base << 16 | (u32)eb << 9 | (u32)l << 5 | (u32)(l + eb);
}

// Builds LUT + canonical description for n code lengths in smem.  All 32 lanes participate.
// Error classification equals Open.codeLengthsToCodeTree (Open.java:705-756): fewer than two codes or
// Kraft sum < 1 -> under-full, Kraft sum > 1 -> over-full.
template <int TB, bool IS_DIST>
__device__ int build_code(const u8 *lens, int n, u32 *lut, u16 *sorted, Canon *cn, u32 lane) {
	if (lane < 16) { cn->cnt[lane] = 0; cn->run[lane] = 0; }
	__syncwarp();
	for (int i = lane; i < n; i += 32) {
		int l = lens[i];
		if (l) atomicAdd(&cn->cnt[l], 1u);
	}
	__syncwarp();
	u32 code = 0, kraft = 0, ncodes = 0, off = 0, prev = 0;
	u32 my_first = 0, my_off = 0;
#pragma unroll
	for (int l = 1; l <= 15; l++) {
		u32 c = cn->cnt[l];
		code = (code + prev) << 1;
		if (lane == (u32)l) { my_first = code; my_off = off; }
		off += c;
		ncodes += c;
		kraft += c << (15 - l);
		prev = c;
	}
	if (lane < 16) { cn->first[lane] = (u16)my_first; cn->offs[lane] = (u16)my_off; }
	__syncwarp();
	if (ncodes < 2) return B2D_HUFFMAN_CODE_UNDER_FULL;
	if (kraft > 32768u) return B2D_HUFFMAN_CODE_OVER_FULL;
	if (kraft < 32768u) return B2D_HUFFMAN_CODE_UNDER_FULL;

	for (int base = 0; base < n; base += 32) {
		int i = base + (int)lane;
		int l = i < n ? lens[i] : 0;
		u32 grp = __match_any_sync(FULL_MASK, l);
		u32 r = __popc(grp & lanemask_lt());
		u32 start = cn->run[l];
		__syncwarp();
		if (r == 0) cn->run[l] = start + __popc(grp);
		__syncwarp();
		if (l) {
			u32 rank = start + r;
			u32 c = cn->first[l] + rank;
			sorted[cn->offs[l] + rank] = (u16)i;
			u32 rev = __brev(c) >> (32 - l);
			if (l <= TB) {
				u32 e = IS_DIST ? d_entry(i, l) : ll_entry(i, l);
				for (u32 j = rev; j < (1u << TB); j += 1u << l) lut[j] = e;
			} else {
				lut[rev & ((1u << TB) - 1)] = IS_DIST ? FD_SPECIAL : F_LONG;
			}
		}
	}
	__syncwarp();
	return 0;
}

// canonical decode of a code longer than the LUT index (replaces the residual tree walk, Open.java:488-492)
template <int TB, bool IS_DIST>
__device__ __noinline__ u32 slow_decode(u32 lo, const Canon *cn, const u16 *sorted) {
	u32 rb = __brev(lo);
	for (int l = TB + 1; l <= 15; l++) {
		u32 c = rb >> (32 - l);
		u32 idx = c - cn->first[l];
		if (idx < cn->cnt[l]) {
			int sym = sorted[cn->offs[l] + idx];
			return IS_DIST ? d_entry(sym, l) : ll_entry(sym, l);
		}
	}
	return IS_DIST ? (FD_SPECIAL | 31u << 16 | 15u << 5 | 15u) : (F_RSVD | 15u << 5 | 15u);   // unreachable for complete codes
}

struct Member {
	BitIn in;
	u8 *out;
	u64 cap;
	u64 pos;
	bool no_dist;        // dynamic block with an empty distance code (Open.java:398-401)
	int tables;          // 0 none, 1 fixed tables resident
};

enum { TOK_EOB = 0, TOK_SWITCH = 1000 };

// Copies a back-reference with all 32 lanes.  Overlap (dist < len) replicates the pattern like the
// reference's byte-serial loop (Open.java:596-603): byte k comes from out[pos - dist + k mod dist].
__device__ __forceinline__ void copy_match(u8 *out, u64 pos, int len, int dist, u32 lane) {
	u8 *dst = out + pos;
	const u8 *src = dst - dist;
	__syncwarp();                       // earlier stores of other lanes (literals, previous copies) are visible
	if (dist >= len) {
		for (int k = lane; k < len; k += 32) dst[k] = src[k];
	} else if (dist == 1) {
		u8 v = src[0];
		for (int k = lane; k < len; k += 32) dst[k] = v;
	} else {
		for (int k = lane; k < len; k += 32) dst[k] = src[k % dist];
	}
	__syncwarp();
}

// Decodes symbols of one Huffman block until end-of-block.  CAREFUL=false requires that two whole real
// words remain behind `next` at the top of every iteration, so no read can pass the end of input and the
// end-of-stream checks are skipped; the CAREFUL=true instantiation checks after every field, in the
// reference's order (Open.java:565-593).
template <bool CAREFUL>
__device__ int decode_tokens(Member &m, WarpSmem *sm, u32 lane) {
	BitIn &b = m.in;
	int avail = CAREFUL ? avail_bits(b) : 0;
	for (;;) {
		if (!CAREFUL && b.widx + 2 > b.n_full) return TOK_SWITCH;
		refill(b);
		u32 lo = (u32)b.buf;
		u32 e = sm->ll_lut[lo & ((1u << LL_TB) - 1)];
		if (e & F_LONG) e = slow_decode<LL_TB, false>(lo, &sm->ll_canon, sm->ll_sorted);
		int clen = (e >> 5) & 15;
		if (CAREFUL && clen > avail) return B2D_UNEXPECTED_END_OF_STREAM;
		if (e & F_LIT) {
			if (m.pos >= m.cap) return B2D_ERR_OUTPUT_OVERFLOW;
			if (lane == 0) m.out[m.pos] = (u8)(e >> 16);
			m.pos++;
			drop(b, clen);
			if (CAREFUL) avail -= clen;
			continue;
		}
		if (e & F_EOB) {
			drop(b, clen);
			__syncwarp();
			return TOK_EOB;
		}
		if (e & F_RSVD) return B2D_RESERVED_LENGTH_SYMBOL;
		int tot = e & 31;
		if (CAREFUL && tot > avail) return B2D_UNEXPECTED_END_OF_STREAM;
		int len = (int)((e >> 16) & 0x7FF) + (int)bfe(lo, clen, (e >> 9) & 15);
		drop(b, tot);
		if (CAREFUL) avail -= tot;
		if (m.no_dist) return B2D_LENGTH_ENCOUNTERED_WITH_EMPTY_DISTANCE_CODE;
		refill(b);
		lo = (u32)b.buf;
		u32 d = sm->d_lut[lo & ((1u << D_TB) - 1)];
		if (d & FD_SPECIAL) {
			if (((d >> 16) & 0x7FFF) == 0) d = slow_decode<D_TB, true>(lo, &sm->d_canon, sm->d_sorted);
		}
		int dclen = (d >> 5) & 15;
		if (CAREFUL && dclen > avail) return B2D_UNEXPECTED_END_OF_STREAM;
		if (d & FD_SPECIAL) return B2D_RESERVED_DISTANCE_SYMBOL;
		int dtot = d & 31;
		if (CAREFUL && dtot > avail) return B2D_UNEXPECTED_END_OF_STREAM;
		int dist = (int)(d >> 16) + (int)bfe(lo, dclen, (d >> 9) & 15);
		drop(b, dtot);
		if (CAREFUL) avail -= dtot;
		if ((u64)dist > m.pos) return B2D_COPY_FROM_BEFORE_DICTIONARY_START;     // Open.java:592-593
		if (m.pos + (u64)len > m.cap) {
			int fit = (int)(m.cap - m.pos);                                       // deliver what fits (Open.java:604-616)
			copy_match(m.out, m.pos, fit, dist, lane);
			m.pos = m.cap;
			return B2D_ERR_OUTPUT_OVERFLOW;
		}
		copy_match(m.out, m.pos, len, dist, lane);
		m.pos += len;
	}
}

// Open.UncompressedBlock (Open.java:227-306)
__device__ int stored_block(Member &m, int &avail, u32 lane) {
	BitIn &b = m.in;
	int err = 0;
	getbits(b, b.cnt & 7, avail, err);                  // align to byte (:234); cnt%8 == unread bits of the byte
	int len = getbits(b, 16, avail, err);
	if (err) return err;
	int nlen = getbits(b, 16, avail, err);
	if (err) return err;
	if (len != (nlen ^ 0xFFFF)) return B2D_UNCOMPRESSED_BLOCK_LENGTH_MISMATCH;   // :239-240
	u64 byte_pos = consumed_bits(b) >> 3;
	u64 in_len = b.total_bits >> 3;
	u64 have = in_len - byte_pos;
	u64 n = (u64)len < have ? (u64)len : have;
	int status = (u64)len > have ? B2D_UNEXPECTED_END_OF_STREAM : 0;            // :279-280
	if (m.pos + n > m.cap) { n = m.cap - m.pos; status = B2D_ERR_OUTPUT_OVERFLOW; }
	const u8 *src = (const u8 *)b.words + (b.lead8 >> 3) + byte_pos;
	u8 *dst = m.out + m.pos;
	// vector body when source and destination share 16-byte phase, bytes otherwise
	if ((((uintptr_t)src ^ (uintptr_t)dst) & 15) == 0 && n >= 64) {
		u64 head = (16 - ((uintptr_t)dst & 15)) & 15;
		for (u64 k = lane; k < head; k += 32) dst[k] = src[k];
		u64 nv = (n - head) >> 4;
		const uint4 *s4 = (const uint4 *)(src + head);
		uint4 *d4 = (uint4 *)(dst + head);
		for (u64 k = lane; k < nv; k += 32) d4[k] = __ldg(s4 + k);
		for (u64 k = head + (nv << 4) + lane; k < n; k += 32) dst[k] = src[k];
	} else {
		for (u64 k = lane; k < n; k += 32) dst[k] = src[k];
	}
	__syncwarp();
	m.pos += n;
	if (status) return status;
	bit_seek(b, byte_pos + n);
	return 0;
}

// Open.HuffmanBlock constructor, dynamic branch (Open.java:336-431)
__device__ int dynamic_header(Member &m, WarpSmem *sm, int &avail, u32 lane) {
	BitIn &b = m.in;
	int err = 0;
	int num_ll = getbits(b, 5, avail, err) + 257;
	int num_d = getbits(b, 5, avail, err) + 1;
	int num_cl = getbits(b, 4, avail, err) + 4;
	if (err) return err;
	// 19 code-length-code lengths, lane s keeps the length of symbol s (order :794-795)
	int my_cl = 0;
	for (int i = 0; i < num_cl; i++) {
		int v = getbits(b, 3, avail, err);
		if (err) return err;
		if (lane == CL_ORDER[i]) my_cl = v;
	}
	// code-length code -> 7-bit LUT (Open.java:344); under/over-full by Kraft sum
	{
		u32 kraft = 0, ncodes = 0, code = 0, prev = 0, my_first = 0;
#pragma unroll
		for (int l = 1; l <= 7; l++) {
			u32 c = __popc(__ballot_sync(FULL_MASK, my_cl == l));
			code = (code + prev) << 1;
			if (my_cl == l) my_first = code;
			ncodes += c;
			kraft += c << (7 - l);
			prev = c;
		}
		if (ncodes < 2) return B2D_HUFFMAN_CODE_UNDER_FULL;
		if (kraft > 128u) return B2D_HUFFMAN_CODE_OVER_FULL;
		if (kraft < 128u) return B2D_HUFFMAN_CODE_UNDER_FULL;
		u32 grp = __match_any_sync(FULL_MASK, my_cl);
		u32 rank = __popc(grp & lanemask_lt());
		if (my_cl && lane < 19) {
			u32 c = my_first + rank;
			u32 rev = __brev(c) >> (32 - my_cl);
			u16 e = (u16)(lane << 4 | my_cl);
			for (u32 j = rev; j < 128u; j += 1u << my_cl) sm->cl_lut[j] = e;
		}
		__syncwarp();
	}
	// code lengths with run-length symbols 16/17/18 (Open.java:347-379)
	int total = num_ll + num_d;
	int run_val = -1;
	for (int i = 0; i < total;) {
		refill(b);
		u32 e = sm->cl_lut[(u32)b.buf & 127u];
		int l = e & 15, sym = e >> 4;
		if (l > avail) return B2D_UNEXPECTED_END_OF_STREAM;
		drop(b, l);
		avail -= l;
		if (sym < 16) {
			run_val = sym;
			if (lane == 0) sm->lens[i] = (u8)sym;
			i++;
		} else {
			int run_len;
			if (sym == 16) {
				if (run_val == -1) return B2D_NO_PREVIOUS_CODE_LENGTH_TO_COPY;      // :359-361, before the extra bits
				run_len = getbits(b, 2, avail, err) + 3;
			} else if (sym == 17) {
				run_val = 0;
				run_len = getbits(b, 3, avail, err) + 3;
			} else {
				run_val = 0;
				run_len = getbits(b, 7, avail, err) + 11;
			}
			if (err) return err;
			if (i + run_len > total) return B2D_CODE_LENGTH_CODE_OVER_FULL;         // :374-375
			for (int k = lane; k < run_len; k += 32) sm->lens[i + k] = (u8)run_val;
			i += run_len;
		}
	}
	__syncwarp();
	if (sm->lens[256] == 0) return B2D_END_OF_BLOCK_CODE_ZERO_LENGTH;               // :383-384
	int e = build_code<LL_TB, false>(sm->lens, num_ll, sm->ll_lut, sm->ll_sorted, &sm->ll_canon, lane);   // :385
	if (e) return e;
	// distance code special cases (:396-428)
	u8 *dl = sm->lens + num_ll;
	m.no_dist = false;
	if (num_d == 1 && dl[0] == 0) {
		m.no_dist = true;
	} else {
		int v = (int)lane < num_d ? dl[lane] : 0;
		u32 ones = __popc(__ballot_sync(FULL_MASK, v == 1));
		u32 others = __popc(__ballot_sync(FULL_MASK, v > 1));
		if (ones == 1 && others == 0) {                                             // :421-425 dummy symbol 31
			__syncwarp();
			if ((int)lane >= num_d) dl[lane] = (lane == 31) ? 1 : 0;
			if (lane == 31) dl[31] = 1;
			num_d = 32;
			__syncwarp();
		}
		e = build_code<D_TB, true>(dl, num_d, sm->d_lut, sm->d_sorted, &sm->d_canon, lane);   // :426
		if (e) return e;
	}
	m.tables = 0;
	return 0;
}

// fixed code of Open.java:812-830 (288 lit/len lengths incl. the reserved 286/287, 32 distance lengths)
__device__ void fixed_tables(Member &m, WarpSmem *sm, u32 lane) {
	for (int i = lane; i < 288; i += 32) sm->lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
	sm->lens[288 + lane] = 5;
	__syncwarp();
	build_code<LL_TB, false>(sm->lens, 288, sm->ll_lut, sm->ll_sorted, &sm->ll_canon, lane);
	build_code<D_TB, true>(sm->lens + 288, 32, sm->d_lut, sm->d_sorted, &sm->d_canon, lane);
	m.no_dist = false;
	m.tables = 1;
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
inflate_kernel(const u8 *__restrict__ in, const u64 *__restrict__ in_off, u32 n_members,
               u8 *out, const u64 *__restrict__ out_off,
               u64 *__restrict__ out_len, u64 *__restrict__ in_consumed, int *__restrict__ status, u32 flags) {
	__shared__ WarpSmem smem[WARPS_PER_CTA];
	u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	u32 mi = blockIdx.x * WARPS_PER_CTA + warp;
	if (mi >= n_members) return;
	WarpSmem *sm = &smem[warp];

	Member m;
	u64 i0 = in_off[mi], i1 = in_off[mi + 1];
	u64 o0 = out_off[mi], o1 = out_off[mi + 1];
	const u8 *src = in + i0;
	u32 lead = (u32)((uintptr_t)src & 3);
	u64 in_len = i1 - i0;
	m.in.words = (const u32 *)(src - lead);
	m.in.lead8 = lead * 8;
	m.in.total_bits = in_len * 8;
	m.in.n_safe = (u32)((lead + in_len + 3) >> 2);
	m.in.n_full = (u32)((lead + in_len) >> 2);
	bit_seek(m.in, 0);
	m.out = out + o0;
	m.cap = o1 - o0;
	m.pos = 0;
	m.no_dist = false;
	m.tables = 0;

	int err = 0;
	bool last = false;
	const bool chunk_mode = (flags & B2D_INFLATE_CHUNK_INDEXED) != 0;
	while (!last) {                                                        // Open.read, Open.java:83-110
		int avail = avail_bits(m.in);
		if (chunk_mode && avail == 0 && (m.in.cnt & 7) == 0) break;        // chunk ends on a block boundary
		last = getbits(m.in, 1, avail, err) != 0;
		int type = getbits(m.in, 2, avail, err);
		if (err) break;
		if (type == 0) {
			err = stored_block(m, avail, lane);
			if (err) break;
			continue;
		}
		if (type == 3) { err = B2D_RESERVED_BLOCK_TYPE; break; }          // :96
		if (type == 1) { if (m.tables != 1) fixed_tables(m, sm, lane); }
		else { err = dynamic_header(m, sm, avail, lane); if (err) break; }
		int r = decode_tokens<false>(m, sm, lane);
		if (r == TOK_SWITCH) r = decode_tokens<true>(m, sm, lane);
		if (r != TOK_EOB) { err = r; break; }
	}
	__syncwarp();
	if (lane == 0) {
		out_len[mi] = m.pos;
		in_consumed[mi] = (consumed_bits(m.in) + 7) >> 3;                  // Open.finish, Open.java:113-124
		status[mi] = err;
	}
}

cudaError_t launch_inflate(const u8 *d_in, const u64 *d_in_off, u32 n, u8 *d_out, const u64 *d_out_off,
                           u64 *d_out_len, u64 *d_in_consumed, int *d_status, u32 flags, cudaStream_t st) {
	if (n == 0) return cudaSuccess;
	u32 grid = (n + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
	inflate_kernel<<<grid, WARPS_PER_CTA * 32, 0, st>>>(d_in, d_in_off, n, d_out, d_out_off, d_out_len,
	                                                     d_in_consumed, d_status, flags);
	return cudaGetLastError();
}

}  // namespace b2d
