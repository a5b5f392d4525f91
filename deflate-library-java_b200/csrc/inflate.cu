// inflate.cu -- batched DEFLATE decoder for sm_100a: one warp per independent member / chunk.
//
// Replaces the decode loops of the reference (paths relative to src/io/nayuki/deflate/):
//   decomp/Open.java:83-110   block loop            -> inflate_kernel()
//   decomp/Open.java:137-170  bit reader            -> BitIn (three words in registers, funnel-shift peek)
//   decomp/Open.java:227-306  stored block          -> stored_block()  (warp-wide coalesced copy)
//   decomp/Open.java:336-431  dynamic header        -> dynamic_header()
//   decomp/Open.java:705-789  code tree + 9-bit LUT -> build_code(): canonical codes built by the whole warp
//                                                      into a 10-bit (lit/len) / 8-bit (distance) LUT in shared
//                                                      memory, longer codes resolved canonically (no tree walk)
//   decomp/Open.java:438-620  symbol loop + copy    -> decode_block_fast() (its loop is the PTX block HOT_LOOP) /
//                                                      decode_block_careful(): every lane decodes the same symbol
//                                                      from shared tables (no divergence, no broadcast needed);
//                                                      output is staged in a shared-memory tile, back-references
//                                                      are queued and resolved 32 at a time by resolve_pending()
// Results (bytes, out_len, consumed input, status) are identical to the reference's; the validation ORDER of
// Open.java is kept (first failing check wins).  Not a translation: no dictionary ring (the output buffer plus
// the staging tile are the window), no code tree, no per-block allocation.
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"

namespace b2d {

constexpr int LL_TB = 10;                 // lit/len LUT index bits
constexpr int D_TB = 8;                   // distance LUT index bits
constexpr int WARPS_PER_CTA = 4;
constexpr int CTAS_PER_SM = 7;            // 28 resident warps per SM: 4096 members fit one wave on 148 SMs
constexpr int TILE = 1024;                // bytes of output staged per warp in shared memory
constexpr u32 LIT_GUARD = 80;             // see decode_block_fast (NEXT_SYMBOL)

// lit/len LUT entry:  [31:27] bits this entry consumes   [19:16] kind   [15:0] value
//   K_LIT    value = the literal byte (a byte store takes it from the low bits as it is)
//   K_LEN    value = the run length 3..258: the extra bits of a length symbol are part of the LUT index whenever
//            code + extra bits fit the index, so the common lengths cost no bit arithmetic at all
//   K_LENX   length symbol whose extra bits do not all fit the index: value = base, [22:20] extra-bit count,
//            [26:23] code length (the consumed-bits field already counts code + extra)
//   K_OTHER  value 0 = end of block, 1 = code longer than the LUT index, 286/287 = reserved length symbol
//            (Open.java:513-517); the consumed-bits field is the code length
// The consumed bits sit on top so that `sh += e >> 27` is one LEA.HI, and nothing but the value sits in the low half
// so that `tile offset + (length << 16)`, the queued form of a reference, is one IMAD on the entry itself.
constexpr u32 K_LIT = 1u << 16;
constexpr u32 K_LEN = 1u << 17;
constexpr u32 K_LENX = 1u << 18;
constexpr u32 K_OTHER = 1u << 19;
// distance LUT entry: [4:0] total bits (code + extra)   [7] special   [12:8] code length + 2   [31:16] distance base
// (the two shift counts sit where a wrap-mode funnel shift can take them straight from the entry; the + 2 because the
// symbol loop's bits come shifted left by two, see decode_block_fast).  The value of a special entry sits where the
// base does and is at least 0xFF00: the "distance" of such an entry is beyond every valid one, so the symbol loop's
// test of the distance against the output position (clamped to POS_CLAMP) catches the special entries as well.
constexpr u32 KD_SPECIAL = 1u << 7;       // value V_DLONG = long code, V_DLONG | 30/31 = reserved symbol (Open.java:546-551),
                                          //       0xFFFF = the block has no distance code (Open.java:398-401)
constexpr u32 V_EOB = 0, V_LONG = 1, V_NODIST = 0xFFFF, V_DLONG = 0xFF00;
// Every lit/len entry that is neither a literal nor a resolved length carries V_NOPAIR in its value: taken for a length
// by the symbol loop, it is longer than the tile and leaves through the exit of the pairs that need care, so the loop
// does not test the kind of a non-literal entry at all.
constexpr u32 V_NOPAIR = 0x8000;
__device__ __forceinline__ u32 ll_value(u32 e) { return e & 0x7FFF; }

__constant__ u8 CL_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct Canon {                            // canonical-code description for the slow path
	u32 cnt[16];
	u32 run[16];
	u16 first[16];
	u16 offs[16];
};

struct __align__(16) WarpSmem {
	u8 tile[TILE];                        // output staging: tile[i] <-> global byte tile_g[i]
	union {                               // the code-length code's LUT is dead once the lengths are read, before
		u16 cl_lut[128];                  // build_code() fills ll_sorted
		u16 ll_sorted[288];
	};
	u8 lens[320];                         // code lengths while a block's tables are built; then (its first 128 bytes)
	                                      // the current line of input of the symbol loop, see decode_block_fast
	// host mirror (see mirror_progress): the member's output base and how many of its bytes are mirrored already.
	// Kept here, reachable from the tile pointer, rather than in the Member.
	u8 *m_out;
	u64 m_done;
	// progress reports (b2d_inflate_batch delivering a pinned, uniformly strided output with the copy engine): the mapped
	// host word that receives how many PROGRESS_PIECE-sized pieces of this member are final in device memory
	u32 *m_prog;
};
struct __align__(16) SideSmem {           // slow-path description of the two codes
	Canon ll_canon, d_canon;
	u16 d_sorted[32];
};
// Shared memory of a CTA, by shared-WINDOW address (on sm_100 the static segment of a CTA starts at 0x400, the first
// KiB being the system's; b2d_init runs smem_base_probe_kernel once per device and refuses the device with
// B2D_ERR_CUDA if that ever differs):
//   0x0400  4 x reference queue, 256 B each        -> "queue full" is `(next slot & 0xFF) == 0`
//   0x0800  4 x distance LUT, 1 KiB each
//   0x1800  4 x SideSmem
//   0x2000  4 x lit/len LUT, 4 KiB each
//   0x6000  4 x WarpSmem
// Every LUT is naturally aligned in the window, so an entry's address is (index bits << 2) OR-ed into the table
// address, one LOP3; and warp w's distance LUT sits at a quarter of its lit/len LUT's address, so the hot loop keeps
// one table address in a register, not two.
static_assert(WARPS_PER_CTA == 4 && LL_TB == 10 && D_TB == 8, "shared-memory layout below is written for these");
static_assert(sizeof(SideSmem) * 4 <= 0x800, "SideSmem area");
constexpr u32 SM_WINDOW_BASE = INFLATE_SMEM_WINDOW_BASE;
constexpr int SM_MQ_OFF = 0x400 - 0x400, SM_D_OFF = 0x800 - 0x400, SM_SIDE_OFF = 0x1800 - 0x400, SM_LL_OFF = 0x2000 - 0x400,
              SM_W_OFF = 0x6000 - 0x400;
constexpr int SM_MEMBER_BYTES = 208;      // a warp's Member (decoder state), see inflate_kernel
constexpr int SM_M_OFF = SM_W_OFF + WARPS_PER_CTA * (int)sizeof(WarpSmem);
constexpr int SM_BYTES = SM_M_OFF + WARPS_PER_CTA * SM_MEMBER_BYTES;
static_assert(7 * (SM_BYTES + 1024) <= 233472, "seven CTAs per SM");
struct Sm {                               // one warp's view, in registers
	WarpSmem *w;
	u32 *ll;                              // 1 << LL_TB entries
	u32 *dl;                              // 1 << D_TB entries
	uint2 *mq;                            // pending back-references: x = shared-window address of the tile byte
	                                      // | length << 16, y = distance
	SideSmem *side;
	__device__ __forceinline__ WarpSmem *operator->() const { return w; }
};
__device__ __forceinline__ Sm warp_smem(u8 *raw, u32 warp) {
	Sm s;                                 // (raw sits at SM_WINDOW_BASE: checked once per device by b2d_init, see the probe below)
	s.ll = (u32 *)(raw + SM_LL_OFF) + (warp << LL_TB);
	s.dl = (u32 *)(raw + SM_D_OFF) + (warp << D_TB);
	s.mq = (uint2 *)(raw + SM_MQ_OFF) + warp * 32;
	s.side = (SideSmem *)(raw + SM_SIDE_OFF) + warp;
	s.w = (WarpSmem *)(raw + SM_W_OFF) + warp;
	return s;
}

// Bit reader: three consecutive 32-bit words of the member live in registers (cur, nxt and a prefetched
// third); peek() funnel-shifts 32 bits out of (cur, nxt) at bit `sh`, consuming bits is `sh += n`, and
// advance() slides the window by one word.  Loads are 4-byte aligned and never pass n_safe.
struct BitIn {
	const u32 *words;     // 4-byte aligned base at or before the member's first byte
	u32 n_safe;           // words that may be read (cover the last real byte)
	u32 n_full;           // words that hold only real bytes
	u32 lead8;            // bits to skip in word 0
	u64 total_bits;       // real bits in the member
	u32 cur, nxt, pre;    // words widx, widx + 1, widx + 2
	u32 widx;
	u32 sh;               // bits of `cur` already consumed; >= 32 means advance() is due
};

// words = the 128-byte line of the member's first byte: word numbers are then line-relative (decode_block_fast keeps one
// line of input per lane), and the lead bytes in front of the member are simply skipped bits
__device__ __forceinline__ void bitin_init(BitIn &b, const u8 *src, u64 in_len) {
	const u32 lead = (u32)((uintptr_t)src & 127);
	b.words = (const u32 *)(src - lead);
	b.lead8 = lead * 8;
	b.total_bits = in_len * 8;
	b.n_safe = (u32)((lead + in_len + 3) >> 2);
	b.n_full = (u32)((lead + in_len) >> 2);
}
__device__ __forceinline__ u32 load_word(const BitIn &b, u32 i) {
	return i < b.n_safe ? __ldg(b.words + i) : 0u;
}
__device__ __forceinline__ void bit_seek(BitIn &b, u64 byte_pos) {     // byte_pos relative to member start
	u64 a = (b.lead8 >> 3) + byte_pos;
	b.widx = (u32)(a >> 2);
	b.sh = (u32)(a & 3) * 8;
	b.cur = load_word(b, b.widx);
	b.nxt = load_word(b, b.widx + 1);
	b.pre = load_word(b, b.widx + 2);
}
__device__ __forceinline__ void advance(BitIn &b) {
	b.sh -= 32;
	b.cur = b.nxt;
	b.nxt = b.pre;
	b.widx++;
	b.pre = load_word(b, b.widx + 2);
}
__device__ __forceinline__ void norm(BitIn &b) { if (b.sh >= 32) advance(b); }
__device__ __forceinline__ u32 peek(const BitIn &b) { return __funnelshift_r(b.cur, b.nxt, b.sh); }   // sh < 32
__device__ __forceinline__ u64 consumed_bits(const BitIn &b) {
	return (u64)b.widx * 32 + b.sh - b.lead8;
}
__device__ __forceinline__ int avail_bits(const BitIn &b) {            // real bits left, clamped to int
	u64 c = consumed_bits(b);
	if (c >= b.total_bits) return 0;
	u64 a = b.total_bits - c;
	return a > 0x3FFFFFFFull ? 0x3FFFFFFF : (int)a;
}

// checked read for headers (Open.readBits, Open.java:137-170): n <= 16
__device__ __forceinline__ int getbits(BitIn &b, int n, int &avail, int &err) {
	norm(b);
	if (n > avail) { err = B2D_UNEXPECTED_END_OF_STREAM; return 0; }
	u32 v = peek(b) & ((1u << n) - 1);
	b.sh += n;
	avail -= n;
	return (int)v;
}

// lit/len entry of symbol `sym` with code length l; `idx_hi` = the index bits above the code (they hold the extra
// bits when the whole symbol fits the index); TB = index bits, or 0 for an entry made outside the LUT
__device__ __forceinline__ u32 ll_entry(int sym, int l, u32 idx_hi, int tb) {
	if (sym < 256) return (u32)l << 27 | K_LIT | (u32)sym;
	if (sym == 256) return (u32)l << 27 | K_OTHER | V_NOPAIR | V_EOB;
	if (sym > 285) return (u32)l << 27 | K_OTHER | V_NOPAIR | (u32)sym;
	int base, eb;
	length_sym_info(sym, base, eb);
	if (l + eb <= tb) return (u32)(l + eb) << 27 | K_LEN | (u32)(base + (int)(idx_hi & ((1u << eb) - 1)));
	return (u32)(l + eb) << 27 | (u32)l << 23 | (u32)eb << 20 | K_LENX | V_NOPAIR | (u32)base;
}
// K_LENX -> K_LEN with the extra bits taken from the stream bits `lo` (code at bit 0)
__device__ __forceinline__ u32 lenx_resolve(u32 e, u32 lo) {
	const u32 eb = (e >> 20) & 7, cl = (e >> 23) & 15;
	return (e & 0xF8000000u) | K_LEN | (ll_value(e) + ((lo >> cl) & ((1u << eb) - 1)));
}
constexpr int POS_CLAMP = 1 << 15;        // (every valid distance is <= 32768)
__device__ __forceinline__ u32 d_entry(int sym, int l) {
	if (sym > 29) return KD_SPECIAL | (V_DLONG | (u32)sym) << 16 | (u32)(l + 2) << 8 | (u32)l;
	int base, eb;
	dist_sym_info(sym, base, eb);
	return (u32)base << 16 | (u32)(l + 2) << 8 | (u32)(l + eb);
}
__device__ __forceinline__ u32 d_code_len(u32 e) { return ((e >> 8) & 31) - 2; }
// base + extra bits of a distance entry; lo4 = the stream's bits at the code, shifted left by two (the low two bits are
// anything): extra = (lo4 & ((4 << tot) - 1)) >> (clen + 2), with both shift counts read by wrap-mode funnel shifts
// from the entry itself
__device__ __forceinline__ u32 entry_value(u32 e, u32 lo4) {
	u32 x = lo4 & ~__funnelshift_l(0u, 0xFFFFFFFCu, e);
	return (e >> 16) + __funnelshift_r(x, 0u, e >> 8);
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
				if (IS_DIST) {
					const u32 e = d_entry(i, l);
					for (u32 j = rev; j < (1u << TB); j += 1u << l) lut[j] = e;
				} else {
					for (u32 j = rev; j < (1u << TB); j += 1u << l) lut[j] = ll_entry(i, l, j >> l, TB);
				}
			} else {
				lut[rev & ((1u << TB) - 1)] = IS_DIST ? (KD_SPECIAL | V_DLONG << 16 | 2u << 8) : ((u32)TB << 27 | K_OTHER | V_NOPAIR | V_LONG);
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
			if (IS_DIST) return d_entry(sym, l);
			const u32 e = ll_entry(sym, l, 0, 0);
			return (e & K_LENX) ? lenx_resolve(e, lo) : e;
		}
	}
	return IS_DIST ? d_entry(31, 15) : (15u << 27 | K_OTHER | V_NOPAIR | 287u);   // unreachable for complete codes
}

// Per-member decoder state.  Output goes through a TILE-byte staging tile in shared memory: tile[i] holds the
// byte of global address tile_g[i] (tile_g is 16-byte aligned), indices [tstart, tpos) are valid and not yet
// flushed.  Literals are stored into the tile as they are decoded; back-references are only QUEUED (mq[], up to
// 32) and resolved together by resolve_pending(), so the loads of all references whose source is already in
// global memory are in flight together instead of one L2 round trip per match, and references into the tile
// itself are served from shared memory.
struct Member {
	BitIn in;
	u8 *out;
	u64 cap;
	u8 *tile_g;          // global address of tile[0]
	u32 tstart, tpos, tlimit;
	int pos_base;        // min(tile_g - out, 1 << 20): output position of tile[0] for the dictionary-start check
	u32 nm;              // queued back-references (warp-uniform)
	int tables;          // 0 none, 1 fixed tables resident
	// block-parallel decode of our own streams (inflate_units_kernel): back-references are not resolved here but
	// appended to a list in global memory; hist_base = output bytes of the chunk that precede this unit
	u64 *glist;
	u32 gcount, gcap, hist_base;
	// host mirror (b2d_inflate_batch with a pinned output buffer): the member's output is also delivered to
	// out + x + mdelta, a mapped host address with the same 128-byte phase, by the decoding warp itself, so no
	// device-to-host copy follows the kernel (see mirror_progress)
	long long mdelta;
	// input streaming (b2d_inflate_batch with a pinned INPUT buffer): the member's compressed bytes are pulled from
	// the mapped host address (device address + hdelta, same 128-byte phase) into device memory by the decoding warp
	// itself, a couple of KiB ahead of the bit reader, so every member of the batch starts at once instead of
	// waiting for a host-to-device copy of the bytes in front of it (see stage_input)
	long long hdelta;
	u64 in_len, staged;  // member bytes; how many of them are in device memory already
	const u32 *landed;   // device word that turns non-zero once the host's own copy of the member's group has landed
	u32 n_full_total;    // in.n_full once everything is staged (until then in.n_full only covers the staged bytes)
#ifdef B2D_PROF
	long long prof[6];
#endif
};
#ifdef B2D_PROF
#define PROF_T0() const long long prof_t0 = clock64()
#define PROF_ADD(m, i) ((m).prof[i] += clock64() - prof_t0)
#else
#define PROF_T0()
#define PROF_ADD(m, i)
#endif

#ifndef B2D_PROF
static_assert(sizeof(Member) <= SM_MEMBER_BYTES, "Member's place in shared memory");
#endif
enum { R_EOB = 0, R_SWITCH = 1000 };

__device__ __forceinline__ u64 out_pos(const Member &m) { return (u64)((long long)(m.tile_g - m.out) + (long long)m.tpos); }

__device__ __forceinline__ void set_tile_origin(Member &m, u64 pos) {
	u8 *a = m.out + pos;
	u32 mis = (u32)((uintptr_t)a & 15);
	m.tile_g = a - mis;
	m.tstart = m.tpos = mis;
	u64 room = m.cap - pos;
	m.tlimit = room >= (u64)(TILE - mis) ? (u32)TILE : mis + (u32)room;
	long long rel0 = (long long)pos - (long long)mis + (long long)m.hist_base;
	m.pos_base = rel0 > POS_CLAMP ? POS_CLAMP : (int)rel0;
}

// shared-memory accesses by 32-bit shared-window address, and plain (L1-cached, coherent) global loads: through a
// generic pointer every access would carry the generic -> shared / -> global conversion
__device__ __forceinline__ u32 lds_u32(u32 a) {
	u32 v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
	return v;
}
// llb >> 2, kept out of the optimiser's sight so that it is recomputed (one shift) where it is used instead of being
// carried, and spilled, as a second loop invariant
__device__ __forceinline__ u32 quarter(u32 a) {
	u32 r;
	asm volatile("shr.u32 %0, %1, 2;" : "=r"(r) : "r"(a));
	return r;
}
__device__ __forceinline__ void sts_u32(u32 a, u32 v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u8(u32 a, u32 v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_v2(u32 a, u32 x, u32 y) {
	asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}

__device__ __forceinline__ u32 lds_u8(u32 a) {
	u32 v;
	asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
	return v;
}
__device__ __forceinline__ u32 ldg_u8(const u8 *p) {
	u32 v;
	asm volatile("ld.global.u8 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ u32 ldg_u32(const u32 *p) {
	u32 v;
	asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

// Materialises the queued back-references into the tile.  Copies replicate the pattern when dist < len exactly
// like the reference's byte-serial loop (Open.java:596-603): byte k comes from position pos - dist + (k mod dist).
// (All state by value: a by-reference Member would be forced into local memory by the call.)
__device__ __noinline__ void resolve_pending(u8 *tile, const uint2 *mq, u8 *tile_g, int ts, u32 nm, u32 lane, u32 qbase) {
	__syncwarp();                                   // literal and queue stores of lane 0 are visible
	const bool have = lane < nm;
	const uint2 q = mq[lane];
	const int off = (int)((q.x & 0xFFFFu) - qbase), len = (int)(q.x >> 16), dist = (int)q.y;
	const int s = off - dist;                        // tile index of the source start (may be negative)
	const bool far = have && (s + len <= ts);        // source lies wholly in global memory (already flushed)
	// (1) short references whose source is final already -- it lies wholly in global memory (far), or wholly in the tile
	// in front of the first destination of this batch (near, but independent of every queued reference): each lane
	// gathers its own with aligned 32-bit loads (only words that hold a needed byte are touched), shifts the up to 16
	// bytes into place and stores them byte by byte; all loads are issued before the first store
	const int first_dst = __shfl_sync(FULL_MASK, off, 0);
#ifdef B2D_NO_NEARIND
	const bool near_ind = false && first_dst;
#else
	const bool near_ind = have && !far && len <= 16 && s >= ts && s + len <= first_dst;
#endif
	{
		const bool mine_g = far && len <= 16, mine = mine_g || near_ind;
		if (__any_sync(FULL_MASK, mine)) {
			const u32 *wp = (const u32 *)(tile_g + (s & ~3));                 // (tile_g is 16-byte aligned, like the tile)
			const u32 ws = (u32)__cvta_generic_to_shared(tile) + (u32)(s & ~3);
			const u32 lead = (u32)s & 3;
			const int nw = mine ? (int)((lead + (u32)len + 3) >> 2) : 0;     // 1..5 words
			u32 w[5];
#pragma unroll
			for (int k = 0; k < 5; k++) w[k] = k < nw ? (mine_g ? ldg_u32(wp + k) : lds_u32(ws + 4 * k)) : 0u;
			u32 v[4];
#pragma unroll
			for (int k = 0; k < 4; k++) v[k] = __funnelshift_r(w[k], w[k + 1], lead * 8);
			const u32 dst = (u32)__cvta_generic_to_shared(tile) + (u32)off;
#pragma unroll
			for (int k = 0; k < 16; k++)
				if (k < len && mine) sts_u8(dst + k, v[k >> 2] >> ((k & 3) * 8));
		}
	}
	// (2) far, long: the warp copies one reference at a time, 32 bytes per step
	const u32 t_s = (u32)__cvta_generic_to_shared(tile);
	for (u32 mask = __ballot_sync(FULL_MASK, far && len > 16); mask; mask &= mask - 1) {
		const int j = __ffs(mask) - 1;
		const int o = __shfl_sync(FULL_MASK, off, j), l = __shfl_sync(FULL_MASK, len, j), si = __shfl_sync(FULL_MASK, s, j);
		const u8 *gs = tile_g + si;
		for (int k = lane; k < l; k += 160) {            // up to five loads in flight per lane, then the stores
			u32 v[5];
#pragma unroll
			for (int j = 0; j < 5; j++) v[j] = k + 32 * j < l ? ldg_u8(gs + k + 32 * j) : 0u;
#pragma unroll
			for (int j = 0; j < 5; j++) if (k + 32 * j < l) sts_u8(t_s + o + k + 32 * j, v[j]);
		}
	}
	__syncwarp();
	// (3) the rest: the source overlaps the tile and may be another queued reference's destination; strictly in stream
	// order, from shared memory
	for (u32 mask = __ballot_sync(FULL_MASK, have && !far && !near_ind); mask; mask &= mask - 1) {
		const int j = __ffs(mask) - 1;
		const int o = __shfl_sync(FULL_MASK, off, j), l = __shfl_sync(FULL_MASK, len, j), d = __shfl_sync(FULL_MASK, dist, j);
		const int si = o - d;
		if (si >= ts) {
			if (d >= l) {
				for (int k = lane; k < l; k += 32) sts_u8(t_s + o + k, lds_u8(t_s + si + k));
			} else if (d == 1) {
				const u32 v = lds_u8(t_s + si);
				for (int k = lane; k < l; k += 32) sts_u8(t_s + o + k, v);
			} else {
				for (int k = lane; k < l; k += 32) sts_u8(t_s + o + k, lds_u8(t_s + si + k % d));
			}
		} else {                                     // source straddles the flushed / staged boundary
			for (int k = lane; k < l; k += 32) {
				const int idx = si + (d < l ? k % d : k);
				sts_u8(t_s + o + k, idx < ts ? ldg_u8(tile_g + idx) : lds_u8(t_s + idx));
			}
		}
		__syncwarp();
	}
}

// Record mode (block-parallel decode of our own streams, units of a foreign stream: m.glist set): the queued references
// are not materialised -- their sources may lie in a part of the output another warp is still producing -- but
// appended to the unit's list in global memory, one record per lane (a coalesced 256-byte store), as
// (position inside the unit | length << 24 | distance << 40).  The symbol loop is the same in both modes.
__device__ __noinline__ void drain_records(u64 *glist, u32 gcount, u32 gcap, const uint2 *mq, u32 nm, long long tile_off, u32 lane) {
	__syncwarp();                                   // queue stores of the symbol loop are visible
	if (lane < nm && gcount + lane < gcap) {
		const uint2 q = mq[lane];
		glist[gcount + lane] = (u64)((long long)(q.x & 0xFFFFu) + tile_off) | (u64)(q.x >> 16) << 24 | (u64)q.y << 40;
	}
	__syncwarp();
}
__device__ __forceinline__ void resolve(Member &m, const Sm &sm, u32 lane) {
	PROF_T0();
	if (m.glist) {
		// tile address -> position inside the unit: - (window address of tile[0]) + (unit offset of tile[0])
		drain_records(m.glist, m.gcount, m.gcap, sm.mq, m.nm,
		              (long long)(m.tile_g - m.out) - (long long)(u32)__cvta_generic_to_shared(sm->tile), lane);
		m.gcount += m.nm;                           // (beyond gcap nothing was stored; the kernels report the overflow)
	} else {
		resolve_pending(sm->tile, sm.mq, m.tile_g, (int)m.tstart, m.nm, lane, (u32)__cvta_generic_to_shared(sm->tile));
	}
	m.nm = 0;
	PROF_ADD(m, 3);
}

// Host mirror.  Stores from an SM reach pinned host memory at PCIe speed only as whole 128-byte lines (measured with
// tools/pcie_probe.cu: 46-49 GB/s against 29 GB/s for tile-shaped pieces at byte alignment, where every piece ends in
// partial lines the host has to merge).  So the mirror does not follow the tile flushes: whenever MIRROR_STEP more
// bytes of the member are final in device memory, the warp copies them (L2 hits) to the host address as whole lines,
// 64 bytes per lane in flight; the unaligned head and tail of the member are the only partial lines.
constexpr u64 MIRROR_STEP = 4096;           // (1 KiB .. 32 KiB measured: no difference end to end)
__device__ __forceinline__ void mirror_copy(const u8 *dev, long long mdelta, u64 n, u32 lane) {
	__syncwarp();                                    // the bytes were stored by other lanes of this warp
	u8 *h = (u8 *)dev + mdelta;
	u64 head = (16 - ((uintptr_t)dev & 15)) & 15;
	if (head > n) head = n;
	for (u64 k = lane; k < head; k += 32) h[k] = dev[k];
	const uint4 *s4 = (const uint4 *)(dev + head);
	uint4 *d4 = (uint4 *)(h + head);
	const u64 nv = (n - head) >> 4;
	u64 i = lane;
	for (; i + 32 < nv; i += 64) {
		const uint4 a = __ldcg(s4 + i), b = __ldcg(s4 + i + 32);     // L2: keep L1 for the decoder
		d4[i] = a; d4[i + 32] = b;
	}
	for (; i < nv; i += 32) d4[i] = __ldcg(s4 + i);
	for (u64 k = head + (nv << 4) + lane; k < n; k += 32) h[k] = dev[k];
}
// Progress reports instead of a mirror (mdelta == MDELTA_PROGRESS): the output stays in device memory and the HOST
// moves it, piece by piece across all members of the batch, with strided copy-engine transfers while the decode goes
// on (api.cu, inflate_host).  A warp tells the host how many whole pieces of its member are final: its stores are
// fenced to system scope first, then one lane writes the count to the member's word in mapped host memory.
constexpr long long MDELTA_PROGRESS = 1;       // (a real mirror delta is a multiple of 128)
__device__ __forceinline__ void report_progress(WarpSmem *w, const u8 *end, bool final, u32 lane) {
	const u32 pieces = final ? 0x7FFFFFFFu : (u32)((u64)(end - w->m_out) >> INFLATE_PROGRESS_SHIFT);
	if (pieces <= (u32)w->m_done) return;
	__threadfence_system();
	__syncwarp();
	if (lane == 0) { *(volatile u32 *)w->m_prog = pieces; w->m_done = pieces; }
	__syncwarp();
}
// end = device address just past the member's bytes that are final in device memory
__device__ __forceinline__ void mirror_progress(WarpSmem *w, const u8 *end, long long mdelta, bool final, u32 lane) {
	if (mdelta == MDELTA_PROGRESS) { report_progress(w, end, final, lane); return; }
	const u8 *base = w->m_out;
	const u64 done = w->m_done, pos = (u64)(end - base);
	u64 upto = pos;
	if (!final) {
		upto = pos - ((uintptr_t)end & 127);                         // whole lines only
		if (upto < done + MIRROR_STEP || upto > pos) return;        // (upto > pos: the subtraction wrapped)
	}
	if (upto > done) mirror_copy(base + done, mdelta, upto - done, lane);
	__syncwarp();
	if (lane == 0) w->m_done = upto;
	__syncwarp();
}

// Writes tile[lo, hi) to global memory (16-byte vectors for the aligned body); with a host mirror (mdelta != 0; `tile`
// is then the tile of a WarpSmem) the mirror is moved on as well.
__device__ __noinline__ void store_tile(const u8 *tile, u8 *g, u32 lo, u32 hi, u32 lane, long long mdelta) {
	__syncwarp();
	const u32 a = (lo + 15) & ~15u, b = hi & ~15u;
	if (a >= b) {
		for (u32 k = lo + lane; k < hi; k += 32) g[k] = tile[k];
	} else {
		if (lo + lane < a) g[lo + lane] = tile[lo + lane];
		for (u32 v = (a >> 4) + lane; v < (b >> 4); v += 32) ((uint4 *)g)[v] = ((const uint4 *)tile)[v];
		if (b + lane < hi) g[b + lane] = tile[b + lane];
	}
	__syncwarp();
	if (mdelta) mirror_progress((WarpSmem *)tile, g + hi, mdelta, false, lane);
}

// Resolves what is queued, writes the staged bytes out and re-bases the tile at the current output position.
__device__ __forceinline__ void flush_tile(Member &m, const Sm &sm, u32 lane) {
	if (m.nm) resolve(m, sm, lane);
	PROF_T0();
	store_tile(sm->tile, m.tile_g, m.tstart, m.tpos, lane, m.mdelta);
	set_tile_origin(m, out_pos(m));
}

// Input streaming.  SM loads reach pinned host memory at the PCIe rate when enough of them are in flight
// (tools/pcie_probe.cu: 51 GB/s with 4096 warps fetching 2 KiB at a time, 45-50 GB/s with decode-sized pauses between
// the fetches and a device-to-host copy running against them), so the decoder can pull its own input: every warp
// copies whole 128-byte lines of its member from the host address to the same place in device memory -- four 16-byte
// loads per lane in flight, then the stores -- and the bit reader only ever sees lines that are complete (in.n_full
// covers the staged bytes; the fast loop leaves with R_SWITCH when it runs out and the caller stages more).  The
// lines at a member's ends are shared with its neighbours, who write the same bytes.
constexpr u32 STAGE_AHEAD = 2048;
__device__ __noinline__ void stage_input(Member &m, u64 upto, u32 lane) {
	if (upto > m.in_len) upto = m.in_len;
	if (upto <= m.staged) return;
	// Behind the warps' backs the host copies the whole blob with the copy engine, group of members after group, and
	// raises a flag per group: from then on the member is simply there.  The warps only pull what they need before that
	// (nothing for the first groups, about half of the input for the last).
	if (m.landed && __shfl_sync(FULL_MASK, *(volatile const u32 *)m.landed, 0) != 0) {
		m.staged = m.in_len;
		m.in.n_full = m.n_full_total;
		return;
	}
	const uintptr_t base = (uintptr_t)m.in.words + (m.in.lead8 >> 3);
	const uintptr_t lo = (base + m.staged) & ~(uintptr_t)127, hi = (base + upto + 127) & ~(uintptr_t)127;
	for (uintptr_t p = lo + lane * 16; p < hi; p += 2048) {
		uint4 v[4];
#pragma unroll
		for (int k = 0; k < 4; k++)
			if (p + k * 512 < hi) v[k] = __ldcv((const uint4 *)(p + k * 512 + m.hdelta));
#pragma unroll
		for (int k = 0; k < 4; k++)
			if (p + k * 512 < hi) *(uint4 *)(p + k * 512) = v[k];
	}
	// the bit reader loads through the non-coherent path (ld.global.nc), which is not ordered behind these stores: they
	// have to be performed before any lane reads the lines (none of them has been read, so none is cached, before)
	__threadfence();
	__syncwarp();
	u64 st = hi - base;
	if (st >= m.in_len) { m.staged = m.in_len; m.in.n_full = m.n_full_total; }
	else { m.staged = st; m.in.n_full = (u32)(((m.in.lead8 >> 3) + st) >> 2); }
}

// Decodes symbols of one Huffman block until end-of-block.  Two implementations with identical results:
//  * decode_block_fast: requires that the three buffered words (cur, nxt, pre) hold only real input at every symbol
//    start (a whole symbol is at most 48 bits, read from bit offset <= 31), so no read can pass the end of input and
//    the end-of-stream checks are skipped.  Its inner loop contains nothing but the two common symbols -- a literal,
//    and a length/distance pair whose reference exists and fits the tile -- with the next LUT lookup issued at the
//    end of the current symbol; everything else (window nearly at the end of input, tile full, reference queue full,
//    rare symbol kinds, split or invalid references) leaves the inner loop with an event code and is handled outside
//    it, which keeps the register allocation of the hot loop free of the calls' save/restore traffic.
//  * decode_block_careful: checks after every field, in the reference's order (Open.java:565-593); used for the
//    last <= 12 bytes of a member and when the member's output slot is nearly full.
// The hot loop is written in PTX.  Every lane of the warp decodes the same symbol, so every branch of the loop is
// uniform -- which the compiler cannot know (the LUT entries come from shared memory): it fences each of them with
// convergence barriers (BSSY / BSYNC / BREAK), duplicates induction variables across the nested exits and spent 15
// instructions on a literal, 40 on a length/distance pair and 24 on a window refill; as written here they are 10, 27
// and 12.  What was measured on the way (tools/loop_bench.py: the loop on synthetic tables, one warp; inflate_probe.py:
// config 2 with one warp per SM and with all 28):
//  * a warp on its own pays for PREDICATES, not instructions: a test and the branch (or predicated instruction) that
//    uses it are 15 cycles apart, so a literal with two branches takes 93 cycles against the 45 of its dependency chain
//    (LDS 29 + four ALU steps) and a pair with six of them 261;
//  * looking the next entry up before the branches (speculatively, behind a literal) brings a lone warp's literal to
//    60 cycles, merged exits and hoisted tests bring a pair to ~215 -- and with 28 warps per SM all of that LOSES: there
//    the time follows the instruction count (13.4 ms per GiB with this loop, 13.8 - 15.9 with the faster-alone ones),
//    so the loop is the one with the fewest instructions;
//  * ptxas "structures" a reducible loop even when every branch is .uni (a BSSY at the top of every trip, all ways back
//    merged into one BSYNC + BRA): a second, never-taken way into the loop makes it irreducible and is left alone.
// The block leaves with ev = EV_BOUNDARY (at a symbol boundary; the window may need a word), EV_PAIR (a pair that
// needs care, or an entry that is neither literal nor length: E's and the distance entry's bits are skipped and tp is
// advanced by the length) or EV_QFULL (the pair took the last slot of the queue).
#define A_CUR "%0"
#define A_NXT "%1"
#define A_PRE "%2"
#define A_SH "%3"
#define A_W "%4"
#define A_TP "%5"
#define A_QP "%6"
#define A_E "%7"
#define A_LEN "%8"
#define A_D "%9"
#define A_LO2 "%10"
#define A_EV "%11"
#define A_BUF "%12"
#define A_WSTOP1 "%13"
#define A_TGUARD "%14"
#define A_POSOFF "%15"
#define A_LLB "%16"
#define A_DLB "%17"
// One word further: the new word goes into the register that held `cur` (dead now), and the loop goes on in the next
// phase, where the three window registers have moved on by one role -- no register is moved on a refill.
#define A_REFILL_INTO(R)                                                                \
	"add.u32 " A_W ", " A_W ", 4;\n\t"                                                  \
	"sub.u32 " A_SH ", " A_SH ", 32;\n\t"                                               \
	"ld.shared.u32 " R ", [" A_W "];\n\t"
// One phase of the loop: C, N, R = the registers that are cur, nxt, pre in it; P, P1, P2 = this phase's label suffix,
// the next one's and the one after that
#define HOT_PHASE(P, P1, P2, C, N, R)                                                   \
		"L_LOOKUP" P ":\n\t"                                                            \
		"shf.r.wrap.b32 t, " C ", " N ", " A_SH ";\n\t"                                 \
		"lop3.b32 t, t, 0xFFC, " A_LLB ", 0xEA;\n\t"                                    \
		"ld.shared.u32 " A_E ", [t];\n\t"                                               \
		/* E = the entry of the symbol at SH (< 32) */                                  \
		"shr.u32 t, " A_E ", 27;\n\t"                                                   \
		"add.u32 " A_SH ", " A_SH ", t;\n"             /* behind the entry's bits (<= 51) */ \
		"L_MID" P ":\n\t"                                                               \
		"and.b32 t, " A_E ", 0x10000;\n\t"                                              \
		"setp.ne.u32 pl, t, 0;\n\t"                                                     \
		"@!pl bra.uni L_NOTLIT" P ";\n\t"                                               \
		/* ---- literal: every lane stores the same byte (one broadcast write) */       \
		"st.shared.u8 [" A_TP "], " A_E ";\n\t"                                         \
		"add.u32 " A_TP ", " A_TP ", 1;\n\t"                                            \
		"setp.lt.u32 p, " A_SH ", 32;\n\t"                                              \
		"@p bra.uni L_LOOKUP" P ";\n\t"                                                 \
		/* it ended in the next word (at most one) */                                   \
		"setp.eq.u32 p, " A_W ", " A_WSTOP1 ";\n\t"                                     \
		"@p bra.uni L_XB" P ";\n\t"                                                     \
		A_REFILL_INTO(C)                                                                \
		"setp.gt.s32 p, " A_TP ", " A_TGUARD ";\n\t"                                    \
		"@!p bra.uni L_LOOKUP" P1 ";\n\t"                                               \
		"bra.uni L_XB" P1 ";\n"                                                         \
		"L_NOTLIT" P ":\n\t"                                                            \
		/* ---- length/distance pair.  A test and the branch (or predicated instruction) that uses it are   \
		   15 cycles apart for a lone warp: every test is issued as early as its inputs allow, its user     \
		   late, with independent work in between. */                                   \
		"and.b32 t, " A_SH ", 32;\n\t"                                                  \
		"setp.ne.u32 p32, t, 0;\n\t"                                                    \
		/* queued form: tile address | length << 16 (the kind bits shift out) */        \
		"shl.b32 t, " A_E ", 16;\n\t"                                                   \
		"add.u32 qx, " A_TP ", t;\n\t"                                                  \
		"add.s32 dmax, " A_POSOFF ", " A_TP ";\n\t"                                     \
		/* the distance code, wherever it starts */                                     \
		"@p32 shf.r.wrap.b32 " A_LO2 ", " N ", " R ", " A_SH ";\n\t"                    \
		"@!p32 shf.r.wrap.b32 " A_LO2 ", " C ", " N ", " A_SH ";\n\t"                   \
		"lop3.b32 t, " A_LO2 ", 0x3FC, " A_DLB ", 0xEA;\n\t"                            \
		"ld.shared.u32 " A_D ", [t];\n\t"                                               \
		/* while it is on its way: does the reference fit the tile with the literal guard kept (tp is      \
		   advanced: the handlers take it back), is this the last free slot of the queue */ \
		"shr.u32 t, qx, 16;\n\t"                       /* (the length: one LEA.HI with the add) */ \
		"add.u32 " A_TP ", " A_TP ", t;\n\t"                                            \
		"add.u32 " A_QP ", " A_QP ", 8;\n\t"                                            \
		"and.b32 t, " A_QP ", 0xFF;\n\t"                                                \
		"setp.ne.u32 pok, t, 0;\n\t"                                                    \
		"setp.le.and.s32 pok, " A_TP ", " A_TGUARD ", pok;\n\t"                         \
		"and.b32 t, " A_D ", 31;\n\t"                                                   \
		"add.u32 " A_SH ", " A_SH ", t;\n\t"                                            \
		"setp.lt.u32 plt, " A_SH ", 32;\n\t"                                            \
		/* dist = base + extra bits (entry_value) */                                    \
		"shf.l.wrap.b32 x, 0, 0xFFFFFFFC, " A_D ";\n\t"                                 \
		"lop3.b32 x, " A_LO2 ", x, 0, 0x30;\n\t"                                        \
		"shr.u32 t, " A_D ", 8;\n\t"                                                    \
		"shf.r.wrap.b32 x, x, 0, t;\n\t"                                                \
		"shr.u32 t, " A_D ", 16;\n\t"                                                   \
		"add.u32 dist, t, x;\n\t"                                                       \
		/* one exit for three reasons (the stub sorts them out).  The source must exist                  \
		   (Open.java:592-593); a special entry (long code, reserved symbol, no distance code) has a       \
		   "distance" beyond every valid one, and an entry that is no length at all (V_NOPAIR) a "length"   \
		   beyond the tile */                                                           \
		"setp.le.and.s32 pok, dist, dmax, pok;\n\t"                                     \
		"@!pok bra.uni L_XP" P ";\n\t"                                                  \
		"st.shared.v2.u32 [" A_QP "+-8], {qx, dist};\n\t" /* same value from every lane: one broadcast write */ \
		"@plt bra.uni L_LOOKUP" P ";\n\t"                                               \
		/* it ended in the next word, or in the one behind that (the pair has kept the literal guard itself) */ \
		"setp.eq.u32 p, " A_W ", " A_WSTOP1 ";\n\t"                                     \
		"@p bra.uni L_XB" P ";\n\t"                                                     \
		A_REFILL_INTO(C)                                                                \
		"setp.lt.u32 p, " A_SH ", 32;\n\t"                                              \
		"@p bra.uni L_LOOKUP" P1 ";\n\t"                                                \
		"setp.eq.u32 p, " A_W ", " A_WSTOP1 ";\n\t"                                     \
		"@p bra.uni L_XB" P1 ";\n\t"                                                    \
		A_REFILL_INTO(N)                                                                \
		"bra.uni L_LOOKUP" P2 ";\n"
#define HOT_LOOP()                                                                      \
	asm volatile("{\n\t"                                                                \
		".reg .b32 t, x, dist, qx, dmax;\n\t"                                      \
		".reg .pred p, pl, plt, p32, pok;\n\t"                                          \
		/* (never taken: a second way into the loop makes it irreducible, which keeps ptxas from        \
		   "structuring" it -- a BSSY at the top of every trip and all ways back merged into one         \
		   BSYNC + BRA -- in spite of the .uni on every branch) */                      \
		"setp.eq.u32 p, " A_LLB ", 0;\n\t"                                              \
		"@p bra.uni L_MID_0;\n"                                                         \
		/* three copies of the loop, the window registers one role further in each */   \
		HOT_PHASE("_0", "_1", "_2", A_CUR, A_NXT, A_PRE)                                \
		HOT_PHASE("_1", "_2", "_0", A_NXT, A_PRE, A_CUR)                                \
		HOT_PHASE("_2", "_0", "_1", A_PRE, A_CUR, A_NXT)                                \
		/* exits: the window back in the roles the caller knows */                      \
		"L_XB_1:\n\t"                                                                   \
		"mov.b32 t, " A_CUR ";\n\tmov.b32 " A_CUR ", " A_NXT ";\n\tmov.b32 " A_NXT ", " A_PRE ";\n\tmov.b32 " A_PRE ", t;\n\t" \
		"bra.uni L_XB_0;\n"                                                             \
		"L_XB_2:\n\t"                                                                   \
		"mov.b32 t, " A_PRE ";\n\tmov.b32 " A_PRE ", " A_NXT ";\n\tmov.b32 " A_NXT ", " A_CUR ";\n\tmov.b32 " A_CUR ", t;\n" \
		"L_XB_0:\n\t"                                                                   \
		"mov.u32 " A_EV ", 2;\n\t"                                                      \
		"bra.uni L_END;\n"                                                              \
		"L_XP_1:\n\t"                                                                   \
		"mov.b32 t, " A_CUR ";\n\tmov.b32 " A_CUR ", " A_NXT ";\n\tmov.b32 " A_NXT ", " A_PRE ";\n\tmov.b32 " A_PRE ", t;\n\t" \
		"bra.uni L_XP_0;\n"                                                             \
		"L_XP_2:\n\t"                                                                   \
		"mov.b32 t, " A_PRE ";\n\tmov.b32 " A_PRE ", " A_NXT ";\n\tmov.b32 " A_NXT ", " A_CUR ";\n\tmov.b32 " A_CUR ", t;\n" \
		"L_XP_0:\n\t"                                  /* which of the three was it? */ \
		"setp.le.s32 p, " A_TP ", " A_TGUARD ";\n\t"                                    \
		"setp.le.and.s32 p, dist, dmax, p;\n\t"                                         \
		"@p bra.uni L_X_QFULL;\n\t"                                                     \
		"sub.u32 " A_QP ", " A_QP ", 8;\n\t"                                            \
		"shr.u32 " A_LEN ", qx, 16;\n\t"                                                \
		"mov.u32 " A_EV ", 4;\n\t"                                                      \
		"bra.uni L_END;\n"                                                              \
		"L_X_QFULL:\n\t"                               /* only the queue: the pair is done, it took the last slot */ \
		"st.shared.v2.u32 [" A_QP "+-8], {qx, dist};\n\t"                               \
		"mov.u32 " A_EV ", 5;\n"                                                        \
		"L_END:\n\t"                                                                    \
		"}"                                                                             \
		: "+r"(cur), "+r"(nxt), "+r"(pre), "+r"(sh), "+r"(wa), "+r"(tp), "+r"(qp), "+r"(e), "=r"(len), "=r"(d), "=r"(lo2), "=r"(ev) \
		: "r"(0), "r"(wa_stop1), "r"(tguard), "r"(pos_off), "r"(llb), "r"(dlb)         \
		: "memory")
enum { EV_BOUNDARY = 2, EV_PAIR = 4, EV_QFULL = 5 };

__device__ __noinline__ int decode_block_fast(Member &m, const Sm &sm, const u32 lane) {
	BitIn &b = m.in;
	// The window holds words of the stream SHIFTED LEFT BY TWO BITS (word k = stream word k << 2 | word k - 1 >> 30):
	// 32 bits taken from it at a stream position are the stream's bits there times four with two stale bits below,
	// i.e. a LUT entry's byte offset after one mask -- the lookup's address is funnel shift + LOP3.  The two bits in
	// front of the member's first word are never looked at.
	u32 sh = b.sh;                                   // sh < 32, widx + 3 <= n_full
	u32 cur = __funnelshift_l(b.widx ? load_word(b, b.widx - 1) : 0u, b.cur, 2);
	u32 nxt = __funnelshift_l(b.cur, b.nxt, 2), pre = __funnelshift_l(b.nxt, b.pre, 2);
	// The output position is kept as the shared-window ADDRESS of the next tile byte (tp = tile_s + tpos), so a
	// literal store needs no address arithmetic; limits and the queued references are in the same terms.
	const u32 tile_s = (u32)__cvta_generic_to_shared(sm->tile);
	const u32 mq_s = (u32)__cvta_generic_to_shared(sm.mq);
	const u32 llb = (u32)__cvta_generic_to_shared(sm.ll);       // 4 KiB-aligned; the distance LUT sits at llb >> 2:
	                                                             // entry address = (bits << 2) masked, OR-ed into the table's
	u32 tp = tile_s + m.tpos;
	u32 tend = tile_s + m.tlimit;                    // end of the usable tile
	int tguard = (int)tend - (int)LIT_GUARD;         // tp <= tguard at every word boundary (see below)
	int pos_off = m.pos_base - (int)tile_s;          // output position of the byte at tp = pos_off + tp
	u32 qp = mq_s + m.nm * 8;                        // next free slot of the reference queue (256 B, 256-aligned:
	                                                 // full when the next slot's address wraps to 0 mod 256)
	// Input: the 128-byte line of the member that holds the word in `pre` sits, shifted, in shared memory (in the code
	// lengths' place: they are dead once the tables are built), the line behind it is on its way in registers, loaded a
	// whole line ahead with one coalesced request: a refill is one shared-memory load from an address that moves on by
	// four -- no global load, no address arithmetic, no memory latency.  (A shuffle from a line held across the lanes
	// costs two instructions more: ptxas guards it with UMOV + BRA.DIV.)  b.words is 128-byte aligned (bitin_init), so
	// word numbers are line-relative as they are.  The three buffered words must stay real input, so only full words
	// (all of them staged) are ever fetched: wstop, the next word number that needs a decision, is the nearer of the
	// next line's first word and the first word that is not full; there the loop is left with the window untouched
	// (the checked path advances by itself).
	const u32 line_s = (u32)__cvta_generic_to_shared(sm->lens);
	u32 line_first = (b.widx + 2) & ~31u;            // number of the line's first word
	u32 wa = line_s + (b.widx + 2 - line_first) * 4;  // shared-window address of the word in `pre`
	u32 wstop, bufn, bufn_lo;
	// (the next line is kept as it was loaded, this lane's word and the one in front of it, and only shifted when its turn
	// comes: shifting it at once would wait for the loads -- 6 % of the kernel when that was tried)
#define LOAD_RAW(i) ((i) >= (b.lead8 >> 5) && (i) < b.n_full ? __ldg(b.words + (i)) : 0u)
#define LOAD_NEXT(first) do { bufn_lo = (first) + lane ? LOAD_RAW((first) + lane - 1) : 0u; bufn = LOAD_RAW((first) + lane); } while (0)
#define NEXT_LINE() do { __syncwarp(); sts_u32(line_s + lane * 4, __funnelshift_l(bufn_lo, bufn, 2)); __syncwarp(); } while (0)
#define WORD_NO() (line_first + ((wa - line_s) >> 2))
	LOAD_NEXT(line_first);
	NEXT_LINE();
	LOAD_NEXT(line_first + 32);
	wstop = min(line_first + 32, b.n_full);
	u32 wa_stop1 = line_s + (wstop - 1 - line_first) * 4;    // address of the last word the loop may move on from
	const u32 dlb = llb >> 2;
	// before resolve() / flush_tile(): the tile's state; before a return: the bit reader's as well (its own, unshifted words)
#define SAVE_TILE() do { m.tpos = tp - tile_s; m.nm = (qp - mq_s) >> 3; } while (0)
#define SAVE_STATE() do { b.widx = WORD_NO() - 2; b.sh = sh; b.cur = load_word(b, b.widx); b.nxt = load_word(b, b.widx + 1); \
                          b.pre = load_word(b, b.widx + 2); SAVE_TILE(); } while (0)
#define LOAD_TILE() do { tp = tile_s + m.tpos; tend = tile_s + m.tlimit; qp = mq_s + m.nm * 8; \
                         pos_off = m.pos_base - (int)tile_s; tguard = (int)tend - (int)LIT_GUARD; } while (0)
	u32 e = 0;
	for (;;) {
		// ---- the hot loop: literals, and length/distance pairs whose reference exists and fits the tile.  Literals are
		// stored without a capacity check: fewer than LIT_GUARD symbols can start before the next word boundary is
		// crossed (<= 79 bits at >= 1 bit each), so room for that many is secured once per word (tp <= tguard).
		u32 ev, d, lo2, len;
#ifdef B2D_PROF
		{ const long long t0 = clock64();
		HOT_LOOP();
		m.prof[5] += clock64() - t0; m.prof[1] += 1; if (ev == EV_PAIR) m.prof[4] += 1; }
#else
		HOT_LOOP();
#endif
		if (ev == EV_PAIR && !(e & K_LEN)) {                 // ---- not a pair at all but a rare entry: decoded here
			sh -= d & 31;                                    // (what the loop did with the "distance" behind it)
			tp -= len;
			sh -= e >> 27;                                   // nothing of it is consumed yet
			const u32 lo = __funnelshift_r(cur, nxt, sh) >> 2;
			if (e & K_LENX) e = lenx_resolve(e, lo);
			else if (ll_value(e) == V_LONG) e = slow_decode<LL_TB, false>(lo, &sm.side->ll_canon, sm->ll_sorted);
			if (e & K_OTHER) {
				if (ll_value(e) == V_EOB) { sh += e >> 27; SAVE_STATE(); return R_EOB; }
				SAVE_STATE();
				return B2D_RESERVED_LENGTH_SYMBOL;
			}
			sh += e >> 27;                                   // (<= 51)
			if (e & K_LIT) {                                 // a literal with a long code
				sts_u8(tp, e);
				tp++;
				ev = EV_BOUNDARY;
			} else {                                         // a length: the distance lookup of the loop, then as below
				len = e & 0xFFFF;
				lo2 = (sh & 32) ? __funnelshift_r(nxt, pre, sh) : __funnelshift_r(cur, nxt, sh);
				d = lds_u32(quarter(llb) | (lo2 & ((4u << D_TB) - 4)));
				sh += d & 31;
				tp += len;
				ev = EV_PAIR;
			}
		}
		if (ev == EV_PAIR) {                                 // ---- a pair that needs care
			tp -= len;                                       // the loop had advanced it already
			if (d & KD_SPECIAL) {
				sh -= d & 31;                                // (what the loop added for the special entry)
				u32 v = d >> 16;
				if (v == V_DLONG) {
					d = slow_decode<D_TB, true>(lo2 >> 2, &sm.side->d_canon, sm.side->d_sorted);
					v = (d & KD_SPECIAL) ? d >> 16 : 0;
				}
				if (v) {
					SAVE_STATE();
					return v == V_NODIST ? B2D_LENGTH_ENCOUNTERED_WITH_EMPTY_DISTANCE_CODE : B2D_RESERVED_DISTANCE_SYMBOL;
				}
				sh += d & 31;
			}
			const u32 dist = entry_value(d, lo2);
			if ((int)dist > pos_off + (int)tp) {             // Open.java:592-593 (the distance bits stay unconsumed)
				sh -= d & 31;
				SAVE_STATE();
				return B2D_COPY_FROM_BEFORE_DICTIONARY_START;
			}
			// A reference that does not fit the tile is split (same distance); when the member's slot is full, what
			// fits is still delivered (Open.java:604-616) before the overflow is reported.
			for (;;) {
				const u32 fit = tend - tp;
				const u32 take = len < fit ? len : fit;
				if (take) {
					sts_v2(qp, tp | take << 16, dist);
					qp += 8;
					tp += take;
					len -= take;
				}
				if (len == 0 && (qp & 0xFF) != 0) break;
				SAVE_TILE();
				if (len == 0) { resolve(m, sm, lane); qp = mq_s; break; }
				flush_tile(m, sm, lane);
				LOAD_TILE();
				if (tp >= tend) { SAVE_STATE(); return B2D_ERR_OUTPUT_OVERFLOW; }
			}
		}
		if (ev == EV_QFULL) {                                // ---- the reference queue is full
			SAVE_TILE();
			resolve(m, sm, lane);
			qp = mq_s;
		}
		// ---- back to a symbol boundary of the hot loop: window (with the change of line), literal guard
		while (sh >= 32) {
			if (wa == wa_stop1) {
				if (wstop >= b.n_full) { SAVE_STATE(); return R_SWITCH; }
				NEXT_LINE();
				line_first = wstop;
				LOAD_NEXT(line_first + 32);
				wstop = min(line_first + 32, b.n_full);
				wa_stop1 = line_s + (wstop - 1 - line_first) * 4;
				wa = line_s - 4;
			}
			wa += 4;
			sh -= 32; cur = nxt; nxt = pre;
			pre = lds_u32(wa);
		}
		if ((int)tp > tguard) {
			SAVE_TILE();
			flush_tile(m, sm, lane);
			LOAD_TILE();
			if ((int)tp > tguard) { SAVE_STATE(); return R_SWITCH; }         // the member's slot is nearly full: checked path
		}
	}
#undef SAVE_STATE
#undef SAVE_TILE
#undef LOAD_RAW
#undef LOAD_TILE
#undef LOAD_NEXT
#undef NEXT_LINE
#undef WORD_NO
}

#ifdef B2D_LOOPBENCH
// Development aid (tools/variants.py ... loopbench=-DB2D_LOOPBENCH): the symbol loop on synthetic tables, one warp.
// mode 0: literals of 5 bits; 1: length/distance pairs of 7 + 5 bits; 2: both, chosen by the stream's bits
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, CTAS_PER_SM) loop_bench_kernel(u32 *out, int mode, int rounds) {
	__shared__ __align__(1024) u8 smem_raw[SM_BYTES];
	const u32 lane = threadIdx.x & 31;
	if (threadIdx.x >= 32) return;
	const Sm sm = warp_smem(smem_raw, 0);
	for (u32 i = lane; i < (1u << LL_TB); i += 32) {
		const bool lit = mode == 0 || (mode == 2 && (__popc(i & 0x1F) & 1));
		sm.ll[i] = lit ? (5u << 27 | K_LIT | 0x41u) : (7u << 27 | K_LEN | 5u);
	}
	for (u32 i = lane; i < (1u << D_TB); i += 32) sm.dl[i] = 4u << 16 | 7u << 8 | 5u;
	__syncwarp();
	const u32 tile_s = (u32)__cvta_generic_to_shared(sm->tile), mq_s = (u32)__cvta_generic_to_shared(sm.mq);
	const u32 llb = (u32)__cvta_generic_to_shared(sm.ll);
	const u32 line_s = (u32)__cvta_generic_to_shared(sm->lens), wa_stop1 = line_s + 31 * 4, dlb = llb >> 2;
	sts_u32(line_s + lane * 4, (lane + 1) * 0x9E3779B9u);
	__syncwarp();
	u32 cur = 0x12345678u, nxt = 0x9ABCDEF1u, pre = 0x0F1E2D3Cu, sh = 0, wa = line_s, words = 0;
	u32 tp = tile_s + 100, qp = mq_s, e = 0;
	const int tguard = (int)tile_s + 900, pos_off = 1 << 20;
	long long bytes = 0, events = 0;
	const long long t0 = clock64();
	for (int r = 0; r < rounds; r++) {
		u32 ev, d, lo2, len;
		const u32 wa0 = wa;
		HOT_LOOP();
		words += (wa - wa0) >> 2;
		events++;
		if (ev == EV_QFULL) qp = mq_s;
		else if (ev != EV_BOUNDARY && ev != EV_PAIR) break;  // (a pair leaving with EV_PAIR ran into the guard: counted as done)
		while (sh >= 32) {                                   // (the same line over and over)
			if (wa + 4 == line_s + 128) wa = line_s - 4;
			wa += 4; words++; sh -= 32; cur = nxt; nxt = pre; pre = lds_u32(wa);
		}
		if ((int)tp > tguard) { bytes += tp - (tile_s + 100); tp = tile_s + 100; }
	}
	const long long t1 = clock64();
	bytes += tp - (tile_s + 100);
	const long long bits = (long long)words * 32 + sh;
	if (lane == 0) { out[0] = (u32)(t1 - t0); out[1] = (u32)bytes; out[2] = (u32)bits; out[3] = (u32)events; }
}
cudaError_t run_loop_bench(int mode, int rounds, uint32_t *out4) {
	u32 *d = nullptr;
	cudaError_t e = cudaMalloc(&d, 16);
	if (e != cudaSuccess) return e;
	for (int k = 0; k < 2; k++) loop_bench_kernel<<<1, WARPS_PER_CTA * 32>>>(d, mode, rounds);
	e = cudaMemcpy(out4, d, 16, cudaMemcpyDeviceToHost);
	cudaFree(d);
	return e;
}
#endif

__device__ __noinline__ int decode_block_careful(Member &m, const Sm &sm, const u32 lane) {
	BitIn &b = m.in;
	u32 cur = b.cur, nxt = b.nxt, pre = b.pre, widx = b.widx, sh = b.sh;
	u32 tpos = m.tpos, tlimit = m.tlimit, nm = m.nm;
	int pos_base = m.pos_base;
	int tguard = (int)tlimit - (int)LIT_GUARD;      // FAST keeps tpos <= tguard at every word boundary
	constexpr bool CAREFUL = true;
	int avail = avail_bits(b);
	const u32 fast_last = b.n_full - 3;              // FAST is only entered with n_full >= 3
	const u32 n_safe = b.n_safe;
	const u32 *const words = b.words;
	const u32 *const ll = sm.ll;
	const u32 *const dl = sm.dl;
	u8 *const tile = sm->tile;
	const u32 tile_s = (u32)__cvta_generic_to_shared(sm->tile);
	int ret;
#define SAVE_STATE() do { b.cur = cur; b.nxt = nxt; b.pre = pre; b.widx = widx; b.sh = sh; m.tpos = tpos; m.nm = nm; } while (0)
#define LOAD_TILE() do { tpos = m.tpos; tlimit = m.tlimit; nm = m.nm; pos_base = m.pos_base; tguard = (int)tlimit - (int)LIT_GUARD; } while (0)
	for (;;) {
		if (sh >= 32) {
			if (CAREFUL) {
				sh -= 32; cur = nxt; nxt = pre; widx++;
				pre = (widx + 2 < n_safe) ? __ldg(words + widx + 2) : 0u;
				if (sh >= 32) {                              // a long symbol crossed two words (rare)
					sh -= 32; cur = nxt; nxt = pre; widx++;
					pre = (widx + 2 < n_safe) ? __ldg(words + widx + 2) : 0u;
				}
			} else {
				// FAST: the three buffered words must stay real input, so the word about to be loaded (widx + 3) has to
				// be a full one; otherwise leave with the window untouched (the checked path advances by itself)
				if (widx >= fast_last) { ret = R_SWITCH; break; }
				sh -= 32; cur = nxt; nxt = pre; widx++;
				pre = __ldg(words + widx + 2);
				if (sh >= 32) {
					if (widx >= fast_last) { ret = R_SWITCH; break; }
					sh -= 32; cur = nxt; nxt = pre; widx++;
					pre = __ldg(words + widx + 2);
				}
				// FAST literals are stored without a capacity check: fewer than LIT_GUARD symbols can start before the
				// next word boundary is crossed (<= 79 bits at >= 1 bit each), so room for that many is secured here.
				if ((int)tpos > tguard) {
					SAVE_STATE();
					flush_tile(m, sm, lane);
					LOAD_TILE();
					if ((int)tpos > tguard) { ret = R_SWITCH; break; }       // the member's slot is nearly full: checked path
				}
			}
		}
		const u32 lo = __funnelshift_r(cur, nxt, sh);
		u32 e = ll[lo & ((1u << LL_TB) - 1)];
	dispatch:
		if (e & K_LIT) {
			if (CAREFUL && (int)(e >> 27) > avail) { ret = B2D_UNEXPECTED_END_OF_STREAM; break; }
			if (CAREFUL && tpos >= tlimit) {
				SAVE_STATE();
				flush_tile(m, sm, lane);
				LOAD_TILE();
				if (tpos >= tlimit) { ret = B2D_ERR_OUTPUT_OVERFLOW; break; }
			}
			tile[tpos] = (u8)e;                              // every lane stores the same byte: one broadcast write, no predicate
			tpos++;
			sh += e >> 27;
			if (CAREFUL) avail -= e >> 27;
			continue;
		}
		if (!(e & K_LEN)) {                                  // rare kinds
			if (e & K_LENX) { e = lenx_resolve(e, lo); goto dispatch; }
			const u32 v = ll_value(e);
			if (v == V_LONG) { e = slow_decode<LL_TB, false>(lo, &sm.side->ll_canon, sm->ll_sorted); goto dispatch; }
			if (CAREFUL && (int)(e >> 27) > avail) { ret = B2D_UNEXPECTED_END_OF_STREAM; break; }
			if (v == V_EOB) { sh += e >> 27; ret = R_EOB; break; }
			ret = B2D_RESERVED_LENGTH_SYMBOL;
			break;
		}
		// code and extra bits of the length are both missing-checked by the consumed-bits field: either is the same
		// UNEXPECTED_END_OF_STREAM (Open.java:565-577)
		if (CAREFUL) {
			if ((int)(e >> 27) > avail) { ret = B2D_UNEXPECTED_END_OF_STREAM; break; }
			avail -= e >> 27;
		}
		u32 len = e & 0xFFFF;
		sh += e >> 27;                                       // <= 51: the distance may start in nxt
		const u32 lo2 = (sh & 32) ? __funnelshift_r(nxt, pre, sh) : __funnelshift_r(cur, nxt, sh);
		u32 d = dl[lo2 & ((1u << D_TB) - 1)];
	dispatch_d:
		if (d & KD_SPECIAL) {
			const u32 v = d >> 16;
			if (v == V_NODIST) { ret = B2D_LENGTH_ENCOUNTERED_WITH_EMPTY_DISTANCE_CODE; break; }
			if (v == V_DLONG) { d = slow_decode<D_TB, true>(lo2, &sm.side->d_canon, sm.side->d_sorted); goto dispatch_d; }
			if (CAREFUL && (int)d_code_len(d) > avail) { ret = B2D_UNEXPECTED_END_OF_STREAM; break; }
			ret = B2D_RESERVED_DISTANCE_SYMBOL;
			break;
		}
		if (CAREFUL) {
			if ((int)d_code_len(d) > avail || (int)(d & 31) > avail) { ret = B2D_UNEXPECTED_END_OF_STREAM; break; }
			avail -= d & 31;
		}
		const u32 dist = entry_value(d, lo2 << 2);
		sh += d & 31;
		// common case: the source exists (Open.java:592-593) and the whole reference fits the tile
		if ((int)dist <= pos_base + (int)tpos && (int)(tpos + len) <= (CAREFUL ? (int)tlimit : tguard)) {
			sm.mq[nm] = make_uint2(tile_s + tpos + (len << 16), dist);   // same value from every lane: one broadcast write
			tpos += len;
			if (++nm == 32) {
				SAVE_STATE();
				resolve(m, sm, lane);
				nm = 0;
			}
			continue;
		}
		if ((int)dist > pos_base + (int)tpos) { ret = B2D_COPY_FROM_BEFORE_DICTIONARY_START; break; }   // Open.java:592-593
		// A reference that does not fit the tile is split (same distance); when the member's slot is full, what
		// fits is still delivered (Open.java:604-616) before the overflow is reported.
		ret = 0;
		for (;;) {
			const u32 fit = tlimit - tpos;
			const u32 take = len < fit ? len : fit;
			if (take) {
				if (lane == 0) sm.mq[nm] = make_uint2((tile_s + tpos) | take << 16, dist);
				nm++;
				tpos += take;
				len -= take;
			}
			if (len == 0 && nm < 32) break;
			SAVE_STATE();
			if (len == 0) { resolve(m, sm, lane); nm = 0; break; }
			flush_tile(m, sm, lane);
			LOAD_TILE();
			if (tpos >= tlimit) { ret = B2D_ERR_OUTPUT_OVERFLOW; break; }
		}
		if (ret) break;
		if (!CAREFUL && (int)tpos > tguard) {                // re-establish the literal guard of the FAST path
			SAVE_STATE();
			flush_tile(m, sm, lane);
			LOAD_TILE();
			if ((int)tpos > tguard) { ret = R_SWITCH; break; }
		}
	}
	SAVE_STATE();
#undef SAVE_STATE
#undef LOAD_TILE
	return ret;
}

// The fast decoder over input that may not be staged yet: when it runs out of staged words (R_SWITCH with input left
// on the host) more is pulled and it goes on.  It returns R_SWITCH only with the whole member in device memory, which
// is what the checked decoder expects.  The unit decoders call the symbol loop through this thin function too: ptxas
// allocates registers across calls, and called straight from THEIR kernel bodies with dozens of live values the loop's
// clone keeps its table address and window limit in local memory (7 LDL per trip instead of 1).  (The plain member
// decoder is the exception: there the direct call gives the clean loop, the wrapper one LDL on the refill path that
// costs 4 % -- tests/test_sass_hot_loop.py checks every clone.)
template <bool STREAM>
__device__ __noinline__ int decode_block_streamed(Member &m, const Sm &sm, const u32 lane) {
	for (;;) {
		if (m.tpos + LIT_GUARD > m.tlimit) flush_tile(m, sm, lane);
		const int r = (m.in.widx + 3 <= m.in.n_full && m.tpos + LIT_GUARD <= m.tlimit) ? decode_block_fast(m, sm, lane) : (int)R_SWITCH;
		if (!STREAM) return r;
		if (r != R_SWITCH || m.staged >= m.in_len) return r;
		const u64 before = m.staged;
		stage_input(m, ((consumed_bits(m.in) + 7) >> 3) + STAGE_AHEAD, lane);       // out of staged input: pull more
		if (m.staged == before) stage_input(m, m.in_len, lane);   // not short of input (the slot is nearly full): the
		                                                           // checked path takes over, with all input at hand
		while (m.in.sh >= 32) advance(m.in);                       // the fast loop left its window untouched
	}
}

// warp-wide global -> global copy: vector body when source and destination share 16-byte (or 4-byte) phase
__device__ __noinline__ void copy_global(u8 *dst, const u8 *src, u64 n, u32 lane) {
	if ((((uintptr_t)src ^ (uintptr_t)dst) & 15) == 0 && n >= 64) {
		u64 head = (16 - ((uintptr_t)dst & 15)) & 15;
		for (u64 k = lane; k < head; k += 32) dst[k] = src[k];
		u64 nv = (n - head) >> 4;
		const uint4 *s4 = (const uint4 *)(src + head);
		uint4 *d4 = (uint4 *)(dst + head);
		for (u64 k = lane; k < nv; k += 32) d4[k] = __ldg(s4 + k);
		for (u64 k = head + (nv << 4) + lane; k < n; k += 32) dst[k] = src[k];
	} else if ((((uintptr_t)src ^ (uintptr_t)dst) & 3) == 0 && n >= 64) {
		u64 head = (4 - ((uintptr_t)dst & 3)) & 3;
		for (u64 k = lane; k < head; k += 32) dst[k] = src[k];
		u64 nv = (n - head) >> 2;
		const u32 *s4 = (const u32 *)(src + head);
		u32 *d4 = (u32 *)(dst + head);
		for (u64 k = lane; k < nv; k += 32) d4[k] = __ldg(s4 + k);
		for (u64 k = head + (nv << 2) + lane; k < n; k += 32) dst[k] = src[k];
	} else {
		for (u64 k = lane; k < n; k += 32) dst[k] = src[k];
	}
}

// Open.UncompressedBlock (Open.java:227-306)
template <bool STREAM>
__device__ int stored_block(Member &m, const Sm &sm, int &avail, u32 lane) {
	BitIn &b = m.in;
	int err = 0;
	norm(b);
	getbits(b, (8 - (b.sh & 7)) & 7, avail, err);       // align to byte (:234)
	int len = getbits(b, 16, avail, err);
	if (err) return err;
	int nlen = getbits(b, 16, avail, err);
	if (err) return err;
	if (len != (nlen ^ 0xFFFF)) return B2D_UNCOMPRESSED_BLOCK_LENGTH_MISMATCH;   // :239-240
	flush_tile(m, sm, lane);                            // the payload goes global -> global, past the tile
	u64 pos = out_pos(m);
	u64 byte_pos = consumed_bits(b) >> 3;
	if (STREAM && m.hdelta) stage_input(m, byte_pos + (u64)len + 64, lane);
	u64 in_len = b.total_bits >> 3;
	u64 have = in_len - byte_pos;
	u64 n = (u64)len < have ? (u64)len : have;
	int status = (u64)len > have ? B2D_UNEXPECTED_END_OF_STREAM : 0;            // :279-280
	if (pos + n > m.cap) { n = m.cap - pos; status = B2D_ERR_OUTPUT_OVERFLOW; }
	const u8 *src = (const u8 *)b.words + (b.lead8 >> 3) + byte_pos;
	u8 *dst = m.out + pos;
	copy_global(dst, src, n, lane);
	if (m.mdelta) mirror_progress(sm.w, dst + n, m.mdelta, false, lane);
	__syncwarp();
	set_tile_origin(m, pos + n);
	if (status) return status;
	bit_seek(b, byte_pos + n);
	return 0;
}

// Open.HuffmanBlock constructor, dynamic branch (Open.java:336-431)
__device__ int dynamic_header(Member &m, const Sm &sm, int &avail, u32 lane) {
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
		norm(b);
		u32 e = sm->cl_lut[peek(b) & 127u];
		int l = e & 15, sym = e >> 4;
		if (l > avail) return B2D_UNEXPECTED_END_OF_STREAM;
		b.sh += l;
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
	int e = build_code<LL_TB, false>(sm->lens, num_ll, sm.ll, sm->ll_sorted, &sm.side->ll_canon, lane);   // :385
	if (e) return e;
	// distance code special cases (:396-428)
	u8 *dl = sm->lens + num_ll;
	if (num_d == 1 && dl[0] == 0) {
		// no distance code: any length symbol is an error (:526-527,578-579), reported by the distance lookup
		for (int i = lane; i < (1 << D_TB); i += 32) sm.dl[i] = KD_SPECIAL | V_NODIST << 16 | 2u << 8;
		__syncwarp();
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
		e = build_code<D_TB, true>(dl, num_d, sm.dl, sm.side->d_sorted, &sm.side->d_canon, lane);   // :426
		if (e) return e;
	}
	m.tables = 0;
	return 0;
}

// fixed code of Open.java:812-830 (288 lit/len lengths incl. the reserved 286/287, 32 distance lengths)
__device__ void fixed_tables(Member &m, const Sm &sm, u32 lane) {
	for (int i = lane; i < 288; i += 32) sm->lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
	sm->lens[288 + lane] = 5;
	__syncwarp();
	build_code<LL_TB, false>(sm->lens, 288, sm.ll, sm->ll_sorted, &sm.side->ll_canon, lane);
	build_code<D_TB, true>(sm->lens + 288, 32, sm.dl, sm.side->d_sorted, &sm.side->d_canon, lane);
	m.tables = 1;
}

// STREAM: the host entry point's variant for a pinned input buffer (the warps pull their input, see stage_input); the
// device-resident path runs the instantiation without any of it.
template <bool STREAM>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, CTAS_PER_SM)
inflate_kernel(const u8 *__restrict__ in, const u64 *__restrict__ in_off, const u64 *__restrict__ in_end, u32 n_members,
               u8 *out, const u64 *__restrict__ out_off,
               u64 *__restrict__ out_len, u64 *__restrict__ in_consumed, int *__restrict__ status, u32 flags,
               long long mdelta, u32 *progress, const u8 *in_host, const u32 *landed, u32 group_size) {
	__shared__ __align__(1024) u8 smem_raw[SM_BYTES];
	u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	u32 mi = blockIdx.x * WARPS_PER_CTA + warp;
	if (mi >= n_members) return;
	const Sm sm = warp_smem(smem_raw, warp);

	// The decoder's state lives in shared memory, one copy per warp (every lane reads and writes the same values): as
	// a local variable it is a copy per LANE in local memory, 6 KB per warp that 28 warps push through what the
	// shared-memory carve-out leaves of L1 -- every hand-over between the symbol loop and its handlers then waited
	// for L2.
	Member &m = *((Member *)(smem_raw + SM_M_OFF) + warp);
	u64 i0 = in_off[mi], i1 = in_end ? in_end[mi] : in_off[mi + 1];     // in_end: members need not be back to back
	u64 o0 = out_off[mi], o1 = out_off[mi + 1];
	const u8 *src = in + i0;
	u64 in_len = i1 > i0 ? i1 - i0 : 0;
	bitin_init(m.in, src, in_len);
	m.n_full_total = m.in.n_full;
	m.in_len = in_len;
	m.hdelta = STREAM && in_host ? (long long)(in_host - in) : 0;
	m.staged = in_len;
	m.landed = STREAM && landed ? landed + mi / group_size : nullptr;
	// (the first fetch differs from member to member: members of a batch consume their input at about the same rate, and
	// fetches that all fall due together would queue up behind each other on the PCIe link, stalling every warp)
	if (STREAM && m.hdelta) { m.staged = 0; m.in.n_full = 0; stage_input(m, STAGE_AHEAD / 2 + ((mi * 2654435761u) >> 27) * 128u, lane); }
	bit_seek(m.in, 0);
	m.out = out + o0;
	m.cap = o1 - o0;
	m.nm = 0;
	m.glist = nullptr;
	m.gcount = m.gcap = m.hist_base = 0;
	m.mdelta = mdelta;
#ifdef B2D_PROF
	for (int i = 1; i < 6; i++) m.prof[i] = 0;
	m.prof[0] = clock64();
#endif
	if (lane == 0) { sm->m_out = m.out; sm->m_done = 0; sm->m_prog = progress ? progress + mi : nullptr; }
	__syncwarp();
	set_tile_origin(m, 0);
	m.tables = 0;

	int err = 0;
	bool last = false;
	const bool chunk_mode = (flags & B2D_INFLATE_CHUNK_INDEXED) != 0;
	while (!last) {                                                        // Open.read, Open.java:83-110
		if (STREAM && m.hdelta) stage_input(m, ((consumed_bits(m.in) + 7) >> 3) + 1024, lane);   // a block header is < 600 bytes
		norm(m.in);
		int avail = avail_bits(m.in);
		if (chunk_mode && avail == 0 && (m.in.sh & 7) == 0) break;         // chunk ends on a block boundary
		last = getbits(m.in, 1, avail, err) != 0;
		int type = getbits(m.in, 2, avail, err);
		if (err) break;
		if (type == 0) {
			err = stored_block<STREAM>(m, sm, avail, lane);
			if (err) break;
			continue;
		}
		if (type == 3) { err = B2D_RESERVED_BLOCK_TYPE; break; }          // :96
		{ PROF_T0();
		if (type == 1) { if (m.tables != 1) fixed_tables(m, sm, lane); }
		else { err = dynamic_header(m, sm, avail, lane); if (err) break; }
		}
		norm(m.in);
		PROF_T0();
		#ifdef B2D_FORCE_CAREFUL
		int r = R_SWITCH;
#else
		int r;
		if (STREAM) r = decode_block_streamed<true>(m, sm, lane);
		else {             // (called straight from here, this kernel's clone of the loop is the one without any local-memory access)
			if (m.tpos + LIT_GUARD > m.tlimit) flush_tile(m, sm, lane);
			r = (m.in.widx + 3 <= m.in.n_full && m.tpos + LIT_GUARD <= m.tlimit) ? decode_block_fast(m, sm, lane) : (int)R_SWITCH;
		}
#endif
		PROF_ADD(m, 2);
		if (r == R_SWITCH) { r = decode_block_careful(m, sm, lane); }
		if (r != R_EOB) { err = r; break; }
	}
	flush_tile(m, sm, lane);                                               // also resolves what is pending
#ifdef B2D_PROF
	if (mi == 7 && lane == 0) printf("PROF n=%u total %lld entries %lld fast %lld resolve %lld rare %lld inloop %lld\n", n_members, clock64() - m.prof[0], m.prof[1], m.prof[2], m.prof[3], m.prof[4], m.prof[5]);
#endif
	if (m.mdelta) mirror_progress(sm.w, m.out + out_pos(m), m.mdelta, true, lane);
	if (lane == 0) {
		out_len[mi] = out_pos(m);
		in_consumed[mi] = (consumed_bits(m.in) + 7) >> 3;                  // Open.finish, Open.java:113-124
		status[mi] = err;
	}
}

// ---------------------------------------------------------------- block-parallel decode of our own streams
// A stream made by b2d_deflate_chunks comes with the bit offset of every DEFLATE block inside its chunk.  The Huffman
// decode -- the serial, expensive part -- then runs with one warp per BLOCK (16 times the units of the chunk index)
// without resolving any back-reference (a block's references reach up to 32 KiB into the previous block, which
// another warp is still producing): phase A writes the literals to their final places and lists the references;
// phase B, one warp per chunk, replays the lists in stream order, 32 references and a staged region of output at
// a time, with the same machinery the member decoder uses (resolve_pending).
constexpr int UNIT_TILE = 4096;                    // bytes of output staged per warp in phase B

__global__ void __launch_bounds__(WARPS_PER_CTA * 32, CTAS_PER_SM)
inflate_units_kernel(const u8 *__restrict__ in, const u64 *__restrict__ chunk_in_off, const u32 *__restrict__ block_bits,
                     u32 n_units, u32 bpc, u32 chunk_bytes, u32 block_bytes, u64 out_total, u8 *out,
                     u64 *__restrict__ glist, u32 gcap, u32 *__restrict__ gcount, int *__restrict__ ustatus) {
	__shared__ __align__(1024) u8 smem_raw[SM_BYTES];
	const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const u32 u = blockIdx.x * WARPS_PER_CTA + warp;
	if (u >= n_units) return;
	const Sm sm = warp_smem(smem_raw, warp);
	const u32 c = u / bpc, bi = u % bpc;
	const u64 pos0 = (u64)c * chunk_bytes + (u64)bi * block_bytes;
	if (pos0 >= out_total) { if (lane == 0) { gcount[u] = 0; ustatus[u] = 0; } return; }
	const u64 expect = min((u64)block_bytes, out_total - pos0);

	Member &m = *((Member *)(smem_raw + SM_M_OFF) + warp);       // (in shared memory: see inflate_kernel)
	const u64 i0 = chunk_in_off[c], i1 = chunk_in_off[c + 1];
	const u8 *src = in + i0;
	const u64 in_len = i1 - i0;
	bitin_init(m.in, src, in_len);
	const u32 bit0 = block_bits[u];
	bit_seek(m.in, bit0 >> 3);
	m.in.sh += bit0 & 7;
	m.out = out + pos0;
	m.cap = expect;
	m.nm = 0;
	m.glist = glist + (u64)u * gcap;
	m.gcount = 0;
	m.gcap = gcap;
	m.hist_base = bi * block_bytes;
	m.mdelta = 0;
	m.hdelta = 0;
	m.landed = nullptr;
	m.in_len = m.staged = in_len;
	m.n_full_total = m.in.n_full;
	set_tile_origin(m, 0);
	m.tables = 0;

	int err = 0;
	while (out_pos(m) < expect) {
		norm(m.in);
		int avail = avail_bits(m.in);
		(void)getbits(m.in, 1, avail, err);                                 // BFINAL is only set on the stream's closing marker
		int type = getbits(m.in, 2, avail, err);
		if (err) break;
		if (type == 0) {
			err = stored_block<false>(m, sm, avail, lane);
			if (err) break;
			continue;
		}
		if (type == 3) { err = B2D_RESERVED_BLOCK_TYPE; break; }
		if (type == 1) { if (m.tables != 1) fixed_tables(m, sm, lane); }
		else { err = dynamic_header(m, sm, avail, lane); if (err) break; }
		norm(m.in);
		int r = decode_block_streamed<false>(m, sm, lane);  // (through the same thin wrapper as the member decoder: see there)
		if (r == R_SWITCH) r = decode_block_careful(m, sm, lane);
		if (r != R_EOB) { err = r; break; }
	}
	flush_tile(m, sm, lane);
	if (err == 0 && (out_pos(m) != expect || m.gcount > m.gcap)) err = B2D_ERR_OUTPUT_OVERFLOW;
	if (lane == 0) { gcount[u] = min(m.gcount, m.gcap); ustatus[u] = err; }
}

struct ResolveSmem {
	__align__(16) u8 tile[UNIT_TILE];
	uint2 mq[32];
};
struct ReplaySmem {
	__align__(16) u8 tile[2][UNIT_TILE];  // the region of the batch being resolved, and the next batch's on its way
	uint2 mq[32];
};
// a batch of phase B: the references k .. k + cnt - 1 of a unit's list and the region of output they fall into
struct ReplayBatch {
	u32 first, mis, cnt, hi;              // staged: tile[mis, hi) <-> tile_g[mis, hi), tile_g 16-byte aligned
	u8 *tile_g;
	u32 head, tail;                       // this lane's byte of the region's unaligned ends (stored once the loads are back)
};
__device__ __forceinline__ ReplayBatch replay_batch(u64 rec, u32 k, u32 n, u8 *ubase, u32 lane) {
	ReplayBatch r;
	const u32 pos = (u32)(rec & 0xFFFFFF), len = (u32)(rec >> 24) & 0xFFFF;
	r.first = __shfl_sync(FULL_MASK, pos, 0);
	u8 *g0 = ubase + r.first;
	r.mis = (u32)((uintptr_t)g0 & 15);
	// as many references as fit the staged region (a single one always does: <= 258 bytes)
	const bool fits = k + lane < n && (pos + len - r.first) + r.mis <= (u32)UNIT_TILE;
	const u32 okmask = __ballot_sync(FULL_MASK, fits);
	r.cnt = okmask == 0xFFFFFFFFu ? 32u : (u32)(__ffs(~okmask) - 1);
	const u32 last_end = __shfl_sync(FULL_MASK, pos + len, r.cnt - 1);
	r.hi = r.mis + (last_end - r.first);
	r.tile_g = g0 - r.mis;
	r.head = r.tail = 0;
	return r;
}
// Stages a batch's region (literals are final, reference bytes are holes that get filled by the replay): the aligned
// middle with asynchronous 16-byte copies (one commit group), the unaligned ends into registers.
__device__ __forceinline__ void replay_stage(u8 *tile, ReplayBatch &r, u32 lane) {
	const u32 a = (r.mis + 15) & ~15u, b = r.hi & ~15u;
	if (a >= b) {                                   // fewer than 31 bytes: one per lane
		if (r.mis + lane < r.hi) r.head = r.tile_g[r.mis + lane];
	} else {
		if (r.mis + lane < a) r.head = r.tile_g[r.mis + lane];
		if (b + lane < r.hi) r.tail = r.tile_g[b + lane];
		const u32 t_s = (u32)__cvta_generic_to_shared(tile);
		for (u32 v = (a >> 4) + lane; v < (b >> 4); v += 32)
			asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(t_s + v * 16), "l"(r.tile_g + (size_t)v * 16) : "memory");
	}
	asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void replay_stage_ends(u8 *tile, const ReplayBatch &r, u32 lane) {
	const u32 a = (r.mis + 15) & ~15u, b = r.hi & ~15u;
	if (a >= b) {
		if (r.mis + lane < r.hi) tile[r.mis + lane] = (u8)r.head;
	} else {
		if (r.mis + lane < a) tile[r.mis + lane] = (u8)r.head;
		if (b + lane < r.hi) tile[b + lane] = (u8)r.tail;
	}
}

// Phase B: the lists are replayed in stream order, a batch of up to 32 references and the region of output they fall
// into at a time.  A batch costs dependent trips to global memory -- its records, its region, its far sources, the
// store -- and a chunk has thousands of batches on one warp, so the next batch's records (two ahead) and region (one
// ahead, into a second tile) are fetched while the current batch is resolved and stored.  Regions of consecutive
// batches are disjoint: nothing the current batch writes is part of what is staged ahead.
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
resolve_units_kernel(u8 *out, const u64 *__restrict__ glist, u32 gcap, const u32 *__restrict__ gcount,
                     const int *__restrict__ ustatus, u32 n_chunks, u32 bpc, u32 chunk_bytes, u32 block_bytes, u64 out_total,
                     int *__restrict__ cstatus) {
	__shared__ ReplaySmem smem[WARPS_PER_CTA];
	const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const u32 c = blockIdx.x * WARPS_PER_CTA + warp;
	if (c >= n_chunks) return;
	ReplaySmem *sm = &smem[warp];
	int err = 0;
	for (u32 bi = 0; bi < bpc && err == 0; bi++) {
		const u32 u = c * bpc + bi;
		const u64 pos0 = (u64)c * chunk_bytes + (u64)bi * block_bytes;
		if (pos0 >= out_total) break;
		err = ustatus[u];
		if (err) break;                                  // the caller re-decodes this chunk serially for the exact outcome
		u8 *ubase = out + pos0;
		const u64 *list = glist + (u64)u * gcap;
		const u32 n = gcount[u];
		if (n == 0) continue;
		u32 k = 0, buf = 0;
		u64 rec = lane < n ? list[lane] : 0;
		__syncwarp();                                    // (the previous block's last store has read its tile)
		ReplayBatch cur = replay_batch(rec, 0, n, ubase, lane);
		replay_stage(sm->tile[0], cur, lane);
		u64 rec_n = cur.cnt + lane < n ? list[cur.cnt + lane] : 0;
		while (k < n) {
			const u32 kn = k + cur.cnt;
			ReplayBatch nxt = cur;
			u64 rec_nn = 0;
			if (kn < n) {                                // the next batch: its region into the other tile, the records behind it
				nxt = replay_batch(rec_n, kn, n, ubase, lane);
				replay_stage(sm->tile[buf ^ 1], nxt, lane);
				rec_nn = kn + nxt.cnt + lane < n ? list[kn + nxt.cnt + lane] : 0;
				asm volatile("cp.async.wait_group 1;" ::: "memory");     // the current batch's region has landed
			} else {
				asm volatile("cp.async.wait_group 0;" ::: "memory");
			}
			replay_stage_ends(sm->tile[buf], cur, lane);
			const u32 pos = (u32)(rec & 0xFFFFFF), len = (u32)(rec >> 24) & 0xFFFF, dist = (u32)(rec >> 40);
			sm->mq[lane] = make_uint2((cur.mis + (pos - cur.first)) | len << 16, dist);
			__syncwarp();
			resolve_pending(sm->tile[buf], sm->mq, cur.tile_g, (int)cur.mis, cur.cnt, lane, 0);
			store_tile(sm->tile[buf], cur.tile_g, cur.mis, cur.hi, lane, 0);
			k = kn; rec = rec_n; rec_n = rec_nn; cur = nxt; buf ^= 1;
		}
	}
	if (lane == 0) cstatus[c] = err;
}

// ---------------------------------------------------------------- ONE foreign stream, decoded in parallel
// A DEFLATE stream from any producer (system gzip, zlib, the reference's own encoder) has no index: the decoder of
// Open.java -- and the member decoder above, one warp per stream -- reads it front to back, a few tens of MB/s on one
// warp.  Here the stream is decoded speculatively instead (the idea of pugz / rapidgzip, rebuilt for the GPU):
//   find     the compressed bytes are cut into segments; one warp per segment looks for the first bit offset in it
//            where a dynamic-Huffman block starts: every lane tests a different offset (block type, HLIT/HDIST in
//            range, code-length code complete), and the survivors get the full header check of Open.java:336-431
//            (dynamic_header: both codes complete, end-of-block present).
//   decode   one warp per found start decodes block after block until it stands exactly at another found start (or
//            has decoded the final block).  It cannot know the 32 KiB in front of it, so -- like phase A of the
//            block-parallel decoder -- literals go to their places in the unit's own buffer and back-references are
//            only listed.
//   replay   the lists are replayed per unit, twice, on two byte planes that begin with a synthetic window: after
//            that every byte of a unit is either known (plane H zero, plane L the byte) or is "byte j of the unknown
//            window" (plane H = 0x80 | j >> 8, plane L = j & 0xFF).  Copies of copies sort themselves out for free.
//   chain    one CTA walks the units in stream order -- a unit must end where the next one starts, otherwise the
//            start was a false positive and its unit is simply never visited -- sums up the output offsets and
//            resolves each unit's LAST 32 KiB against the window its predecessor left: the only serial step, 32 KiB
//            per unit.
//   finish   every unit resolves all its bytes against its predecessor's window and writes them to the output.
// Anything unusual (no block start found in a long stretch, a unit that outgrows its buffer, a reference reaching
// before the start of the stream, a decode error) makes the result say so, and the host entry point decodes the stream
// again with the sequential decoder, whose status and delivered bytes are the reference's.
constexpr u64 NO_START = ~0ull;
constexpr u32 STREAM_END = 0xFFFFFFFFu;
constexpr u32 STREAM_WINDOW = 32768;

struct StreamUnit {
	u64 start_bit;       // where a block starts inside the segment (NO_START: none found)
	u64 end_bit;         // where the unit stopped (a block boundary)
	u32 out_len;         // bytes it produced
	u32 n_refs;          // back-references it listed
	u32 next;            // segment whose start it stopped at; STREAM_END: it decoded the final block
	int status;
};
struct StreamResult {
	u64 out_len, in_consumed;
	u64 status;          // 0, or nonzero: decode sequentially for the exact outcome
	u64 crc;             // filled by the caller's checksum kernel
	u32 n_live, pad;
};

__device__ __forceinline__ void stream_bitin(BitIn &b, const u8 *in, u64 in_len, u64 bit) {
	bitin_init(b, in, in_len);
	bit_seek(b, bit >> 3);
	b.sh += (u32)(bit & 7);
}

// Second stage of the block-start filter, one candidate per LANE (the first stage leaves a few hundred candidates per
// segment; checking them one at a time with the whole warp -- dynamic_header builds its decode tables -- was most of the
// kernel): the lane decodes the code-length sequence of its candidate with a canonical decoder of its own and checks
// what Open.java:336-431 / :705-756 check -- the run lengths fit, end-of-block has a code, the literal/length code
// is complete, the distance code is complete or one of the two special cases.  `bit` is the block's first header bit.
__device__ bool header_plausible(const u32 *__restrict__ words, u64 n_words, u64 bit, u64 total_bits) {
	u64 wi = bit >> 5;
	u64 buf = 0;
	int cnt = 0;
	u64 pos = bit;                                   // bits consumed so far (absolute)
	auto refill = [&]() {
		while (cnt <= 32) { buf |= (u64)(wi < n_words ? __ldg(words + wi) : 0u) << cnt; wi++; cnt += 32; }
	};
	refill();
	{ const int skip = (int)(bit & 31); buf >>= skip; cnt -= skip; }
	auto take = [&](int nb) { refill(); const u32 v = (u32)buf & ((1u << nb) - 1); buf >>= nb; cnt -= nb; pos += nb; return v; };
	take(3);
	const u32 n_ll = take(5) + 257, n_d = take(5) + 1, n_cl = take(4) + 4;
	u8 cl[19];
#pragma unroll
	for (int i = 0; i < 19; i++) cl[i] = 0;
	for (u32 i = 0; i < n_cl; i++) cl[CL_ORDER[i]] = (u8)take(3);
	// canonical code of the code-length alphabet: symbols by (length, symbol)
	u8 count[8], sorted[19];
#pragma unroll
	for (int l = 0; l < 8; l++) count[l] = 0;
	for (int i = 0; i < 19; i++) count[cl[i]]++;
	{
		u8 offs[8];
		offs[1] = 0;
		for (int l = 1; l < 7; l++) offs[l + 1] = offs[l] + count[l];
		for (int i = 0; i < 19; i++) if (cl[i]) sorted[offs[cl[i]]++] = (u8)i;
	}
	const u32 total = n_ll + n_d;
	u32 kraft_ll = 0, codes_ll = 0, kraft_d = 0, codes_d = 0, ones_d = 0, eob = 0;
	int prev = -1;
	for (u32 i = 0; i < total;) {
		if (pos + 16 > total_bits) return false;
		refill();
		int code = 0, first = 0, index = 0, sym = -1;
		for (int l = 1; l <= 7; l++) {                  // one bit at a time, the code's first bit first
			code |= (int)(buf & 1);
			buf >>= 1; cnt--; pos++;
			const int c = count[l];
			if (code - c < first) { sym = sorted[index + (code - first)]; break; }
			index += c;
			first = (first + c) << 1;
			code <<= 1;
		}
		if (sym < 0) return false;
		u32 rep = 1;
		int len = sym;
		if (sym == 16) { if (prev < 0) return false; len = prev; rep = 3 + take(2); }
		else if (sym == 17) { len = 0; rep = 3 + take(3); }
		else if (sym == 18) { len = 0; rep = 11 + take(7); }
		if (i + rep > total) return false;
		prev = len;
		for (u32 r = 0; r < rep; r++, i++) {
			if (len == 0) continue;
			if (i < n_ll) { kraft_ll += 32768u >> len; codes_ll++; if (i == 256) eob = 1; }
			else { kraft_d += 32768u >> len; codes_d++; ones_d += len == 1; }
		}
	}
	if (!eob || codes_ll < 2 || kraft_ll != 32768u) return false;
	if (codes_d == 0) return n_d == 1;                                  // no distance code at all (Open.java:398-401)
	if (codes_d == 1) return ones_d == 1;                               // one code of length 1 (:421-425)
	return kraft_d == 32768u;
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32, CTAS_PER_SM)
stream_find_kernel(const u8 *__restrict__ in, u64 in_len, u32 seg_bytes, u32 n_seg, u32 n_units, StreamUnit *__restrict__ units,
                   u32 *__restrict__ pool) {
	__shared__ __align__(1024) u8 smem_raw[SM_BYTES];
	const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const u32 k = blockIdx.x * WARPS_PER_CTA + warp;
	if (k >= n_units) return;
	if (k == 0 && lane == 0) *pool = 0;
	if (k >= n_seg) {                                  // a spare unit (see stream_units_kernel): nobody starts here
		if (lane == 0) {
			StreamUnit u;
			u.start_bit = NO_START; u.end_bit = 0; u.out_len = 0; u.n_refs = 0; u.next = STREAM_END; u.status = 0;
			units[k] = u;
		}
		return;
	}
	const Sm sm = warp_smem(smem_raw, warp);
	u64 found = k == 0 ? 0 : NO_START;
	const u64 total_bits = in_len * 8;
	const u32 *__restrict__ words = (const u32 *)in;
	const u64 n_words = (in_len + 3) >> 2;
	const u64 b0 = (u64)k * seg_bytes * 8, b1 = min(total_bits, b0 + (u64)seg_bytes * 8);
	// Stage 1, a bit offset per lane: block type, HLIT / HDIST in range, code-length code complete.  Its survivors (one
	// offset in a few hundred) are queued until there are 32 of them, stage 2 (header_plausible) then checks one per
	// lane, and what survives that -- in practice only true block starts -- is confirmed with the decoder's own header
	// routine, lowest offset first.
	u64 my_cand = NO_START;                            // the queued candidate of this lane
	u32 n_cand = 0;
	auto flush_candidates = [&]() {
		const bool good = my_cand != NO_START && header_plausible(words, n_words, my_cand, total_bits);
		u32 mask = __ballot_sync(FULL_MASK, good);
		while (mask && found == NO_START) {             // (queued in ascending order: lane 0 holds the lowest offset)
			const u64 c = __shfl_sync(FULL_MASK, my_cand, __ffs(mask) - 1);
			mask &= mask - 1;
			Member m;
			stream_bitin(m.in, in, in_len, c);
			m.tables = 0;
			int avail = avail_bits(m.in), err = 0;
			(void)getbits(m.in, 1, avail, err);
			const int type = getbits(m.in, 2, avail, err);
			if (err == 0 && type == 2 && dynamic_header(m, sm, avail, lane) == 0) found = c;
			__syncwarp();
		}
		my_cand = NO_START;
		n_cand = 0;
	};
	for (u64 base = b0; base < b1 && found == NO_START; base += 32) {
		const u64 wi = base >> 5;
		const u32 w0 = wi < n_words ? __ldg(words + wi) : 0u, w1 = wi + 1 < n_words ? __ldg(words + wi + 1) : 0u,
		          w2 = wi + 2 < n_words ? __ldg(words + wi + 2) : 0u, w3 = wi + 3 < n_words ? __ldg(words + wi + 3) : 0u;
		// 96 bits from bit offset base + lane
		const u32 v0 = __funnelshift_r(w0, w1, lane), v1 = __funnelshift_r(w1, w2, lane), v2 = __funnelshift_r(w2, w3, lane);
		const u64 bit = base + lane;
		bool ok = bit < b1 && bit + 96 <= total_bits && ((v0 >> 1) & 3) == 2;       // BTYPE = 10 (Open.java:88-98)
		const u32 hlit = (v0 >> 3) & 31, hdist = (v0 >> 8) & 31, hclen = ((v0 >> 13) & 15) + 4;
		ok = ok && hlit <= 29 && hdist <= 29;
		if (ok) {                                          // the code-length code must be complete (Open.java:344, :705-756)
			const u64 lo = (u64)v0 | (u64)v1 << 32;
			const u64 f = (lo >> 17) | ((u64)v2 << 47);     // bits 17 .. 80
			u32 kraft = 0, cnt = 0;
			for (u32 i = 0; i < hclen; i++) {
				const u32 l = (u32)(f >> (3 * i)) & 7;
				kraft += l ? 128u >> l : 0u;
				cnt += l != 0;
			}
			ok = kraft == 128 && cnt >= 2;
		}
		u32 mask = __ballot_sync(FULL_MASK, ok);
		while (mask) {                                     // queue them, lowest offset into the lowest free lane
			const u32 room = 32 - n_cand, have = (u32)__popc(mask);
			const u32 takes = min(room, have);
			// the r-th set bit of mask goes to lane n_cand + r
			const u32 r = lane - n_cand;                   // which of the new ones this lane would take
			if (lane >= n_cand && r < takes) {
				u32 mm = mask;
				for (u32 q = 0; q < r; q++) mm &= mm - 1;
				my_cand = base + (u32)(__ffs(mm) - 1);
			}
			for (u32 q = 0; q < takes; q++) mask &= mask - 1;
			n_cand += takes;
			if (n_cand == 32) { flush_candidates(); if (found != NO_START) break; }
		}
	}
	if (found == NO_START && n_cand) flush_candidates();
	if (lane == 0) {
		StreamUnit u;
		u.start_bit = found; u.end_bit = 0; u.out_len = 0; u.n_refs = 0; u.next = STREAM_END; u.status = 0;
		units[k] = u;
	}
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32, CTAS_PER_SM)
stream_units_kernel(const u8 *__restrict__ in, u64 in_len, u32 n_seg, u32 n_units, StreamUnit *units, u8 *planeL, u64 stride, u32 cap,
                    u64 *__restrict__ glist, u32 gcap, u32 *__restrict__ pool) {
	__shared__ __align__(1024) u8 smem_raw[SM_BYTES];
	const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	u32 k = blockIdx.x * WARPS_PER_CTA + warp;
	if (k >= n_seg) return;
	const u64 start = units[k].start_bit;
	if (start == NO_START) return;
	const Sm sm = warp_smem(smem_raw, warp);
	Member &m = *((Member *)(smem_raw + SM_M_OFF) + warp);       // (in shared memory: see inflate_kernel)
	stream_bitin(m.in, in, in_len, start);
	m.out = planeL + (u64)k * stride + STREAM_WINDOW;
	m.cap = cap;
	m.nm = 0;
	m.glist = glist + (u64)k * gcap;
	m.gcount = 0;
	m.gcap = gcap;
	m.hist_base = k ? STREAM_WINDOW : 0;               // the first unit knows that nothing precedes it (Open.java:592-593)
	m.mdelta = 0;
	m.hdelta = 0;
	m.landed = nullptr;
	m.in_len = m.staged = in_len;
	m.n_full_total = m.in.n_full;
	set_tile_origin(m, 0);
	m.tables = 0;
	int err = 0;
	u32 next = STREAM_END, j = k + 1;
	bool first = true;
	for (;;) {
		norm(m.in);
		if (!first) {                                  // at a block boundary: does another unit start exactly here?
			const u64 pos = consumed_bits(m.in);
			while (j < n_seg) {
				const u64 sj = units[j].start_bit;
				if (sj != NO_START && sj >= pos) break;
				j++;
			}
			if (j < n_seg && units[j].start_bit == pos) { next = j; break; }
			// Nobody starts here, and the unit's buffer is half full (long stretches of stored or fixed blocks have no
			// dynamic block to restart at): the warp goes on in a SPARE unit -- a fresh buffer, list and unknown
			// window -- and the chain leads from this unit to that one like to any other.
			if (out_pos(m) > cap / 2) {
				flush_tile(m, sm, lane);
				u32 s = 0;
				if (lane == 0) s = n_seg + atomicAdd(pool, 1u);
				s = __shfl_sync(FULL_MASK, s, 0);
				if (s >= n_units) { err = B2D_ERR_OUTPUT_OVERFLOW; break; }
				if (lane == 0) {
					units[k].end_bit = pos;
					units[k].out_len = (u32)out_pos(m);
					units[k].n_refs = min(m.gcount, m.gcap);
					units[k].next = s;
					units[k].status = m.gcount > m.gcap ? B2D_ERR_OUTPUT_OVERFLOW : 0;
					units[s].start_bit = pos;
				}
				__syncwarp();
				k = s;
				m.out = planeL + (u64)k * stride + STREAM_WINDOW;
				m.glist = glist + (u64)k * gcap;
				m.gcount = 0;
				m.hist_base = STREAM_WINDOW;
				set_tile_origin(m, 0);
			}
		}
		first = false;
		int avail = avail_bits(m.in);
		const bool last = getbits(m.in, 1, avail, err) != 0;
		const int type = getbits(m.in, 2, avail, err);
		if (err) break;
		if (type == 0) {
			err = stored_block<false>(m, sm, avail, lane);
			if (err) break;
		} else {
			if (type == 3) { err = B2D_RESERVED_BLOCK_TYPE; break; }
			if (type == 1) { if (m.tables != 1) fixed_tables(m, sm, lane); }
			else { err = dynamic_header(m, sm, avail, lane); if (err) break; }
			norm(m.in);
			int r = decode_block_streamed<false>(m, sm, lane);   // (through the same thin wrapper as the member decoder: see there)
			if (r == R_SWITCH) r = decode_block_careful(m, sm, lane);
			if (r != R_EOB) { err = r; break; }
		}
		if (last) break;                                // next stays STREAM_END
	}
	flush_tile(m, sm, lane);
	if (lane == 0) {
		units[k].end_bit = consumed_bits(m.in);
		units[k].out_len = (u32)out_pos(m);
		units[k].n_refs = min(m.gcount, m.gcap);
		units[k].next = next;
		units[k].status = err == 0 && m.gcount > m.gcap ? B2D_ERR_OUTPUT_OVERFLOW : err;
	}
}

// the two planes of every unit: plane H zero over the unit's bytes, and the synthetic windows in front of both
__global__ void __launch_bounds__(256)
stream_planes_kernel(const StreamUnit *__restrict__ units, u8 *planeL, u8 *planeH, u64 stride) {
	const u32 k = blockIdx.x;
	if (units[k].start_bit == NO_START || units[k].status != 0) return;
	u8 *L = planeL + (u64)k * stride, *H = planeH + (u64)k * stride;
	for (u32 i = threadIdx.x; i < STREAM_WINDOW / 4; i += blockDim.x) {
		const u32 j = 4 * i;                            // window bytes j .. j + 3
		((u32 *)L)[i] = (j & 0xFF) | ((j + 1) & 0xFF) << 8 | ((j + 2) & 0xFF) << 16 | ((j + 3) & 0xFF) << 24;
		((u32 *)H)[i] = (0x80u | j >> 8) * 0x01010101u;
	}
	const u32 n16 = (units[k].out_len + 15) >> 4;
	uint4 *h4 = (uint4 *)(H + STREAM_WINDOW);
	for (u32 i = threadIdx.x; i < n16; i += blockDim.x) h4[i] = make_uint4(0, 0, 0, 0);
}

// replays a unit's reference list on one plane (blockIdx.y: 0 = L, 1 = H), exactly like resolve_units_kernel
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
stream_replay_kernel(const StreamUnit *__restrict__ units, u32 n_seg, u8 *planeL, u8 *planeH, u64 stride,
                     const u64 *__restrict__ glist, u32 gcap) {
	__shared__ ResolveSmem smem[WARPS_PER_CTA];
	const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const u32 k = blockIdx.x * WARPS_PER_CTA + warp;
	if (k >= n_seg) return;
	if (units[k].start_bit == NO_START || units[k].status != 0) return;
	ResolveSmem *sm = &smem[warp];
	u8 *ubase = (blockIdx.y ? planeH : planeL) + (u64)k * stride + STREAM_WINDOW;
	const u64 *list = glist + (u64)k * gcap;
	const u32 n = units[k].n_refs;
	for (u32 q = 0; q < n;) {
		const u64 rec = q + lane < n ? list[q + lane] : 0;
		const u32 pos = (u32)(rec & 0xFFFFFF), len = (u32)(rec >> 24) & 0xFFFF, dist = (u32)(rec >> 40);
		const u32 first = __shfl_sync(FULL_MASK, pos, 0);
		u8 *g0 = ubase + first;
		const u32 mis = (u32)((uintptr_t)g0 & 15);
		const bool fits = q + lane < n && (pos + len - first) + mis <= (u32)UNIT_TILE;
		const u32 okmask = __ballot_sync(FULL_MASK, fits);
		const u32 cnt = okmask == 0xFFFFFFFFu ? 32u : (u32)(__ffs(~okmask) - 1);
		const u32 last_end = __shfl_sync(FULL_MASK, pos + len, cnt - 1);
		const u32 hi = mis + (last_end - first);
		u8 *tile_g = g0 - mis;
		__syncwarp();
		{
			const u32 a = (mis + 15) & ~15u, b = hi & ~15u;
			if (a >= b) {
				for (u32 i = mis + lane; i < hi; i += 32) sm->tile[i] = tile_g[i];
			} else {
				if (mis + lane < a) sm->tile[mis + lane] = tile_g[mis + lane];
				for (u32 v = (a >> 4) + lane; v < (b >> 4); v += 32) ((uint4 *)sm->tile)[v] = ((const uint4 *)tile_g)[v];
				if (b + lane < hi) sm->tile[b + lane] = tile_g[b + lane];
			}
		}
		sm->mq[lane] = make_uint2((mis + (pos - first)) | len << 16, dist);
		resolve_pending(sm->tile, sm->mq, tile_g, (int)mis, cnt, lane, 0);
		store_tile(sm->tile, tile_g, mis, hi, lane, 0);
		q += cnt;
	}
}

// The chain, part 1: the units in stream order and their output offsets (one warp; the units are read 32 at a time and
// followed with shuffles, since a unit's successor is nearly always one of the next few segments).
__global__ void __launch_bounds__(32)
stream_walk_kernel(const StreamUnit *__restrict__ units, u32 n_seg, u32 *__restrict__ live, u64 *__restrict__ off, u64 out_cap,
                   StreamResult *res) {
	const u32 lane = threadIdx.x;
	u32 k = 0, t = 0, base = 0xFFFFFFFFu;
	u64 total = 0, prev_end = 0;
	int status = 0;
	StreamUnit mine;
	mine.start_bit = NO_START; mine.end_bit = 0; mine.out_len = 0; mine.n_refs = 0; mine.next = STREAM_END; mine.status = 0;
	for (;;) {
		if (k < base || k >= base + 32) {                  // (re)load the batch that starts at k
			base = k;
			if (base + lane < n_seg) mine = units[base + lane];
			else mine.start_bit = NO_START;
		}
		const u32 j = k - base;
		const u64 start = __shfl_sync(FULL_MASK, mine.start_bit, j), end = __shfl_sync(FULL_MASK, mine.end_bit, j);
		const u32 n = __shfl_sync(FULL_MASK, mine.out_len, j), next = __shfl_sync(FULL_MASK, mine.next, j);
		const int st = __shfl_sync(FULL_MASK, mine.status, j);
		if (k >= n_seg || start == NO_START || start != prev_end || t >= n_seg) { status = B2D_ERR_BAD_ARGUMENT; break; }   // the chain is broken
		if (st != 0) { status = st; break; }
		if (total + n > out_cap) { status = B2D_ERR_OUTPUT_OVERFLOW; break; }
		if (lane == 0) { live[t] = k; off[t] = total; }
		total += n;
		prev_end = end;
		t++;
		if (next == STREAM_END) break;
		k = next;
	}
	if (lane == 0) {
		off[t] = total;
		res->n_live = t;
		res->out_len = total;
		res->in_consumed = (prev_end + 7) >> 3;            // Open.finish, Open.java:113-124
		res->status = (u64)(long long)status;
		res->crc = 0;
	}
}

// The chain, part 2.  What unit t leaves behind is a MAP of the window it found to the window it leaves: entry i is a
// byte (the unit produced it, or copied a known byte there) or "byte j of the window I found" (0x8000 | j: the unit did
// not reach that far back, or copied it from there).  Maps compose -- (A after B)[i] = A[i] is a reference ? B[A[i]] : A[i]
// -- and composition is associative, so the windows of ALL units come out of a parallel prefix scan over the maps
// (Hillis-Steele, log2(units) rounds, every round all units at once) instead of a walk along the stream.
constexpr u16 MAP_REF = 0x8000;
__global__ void __launch_bounds__(256)
stream_maps_kernel(const StreamUnit *__restrict__ units, const u32 *__restrict__ live, const StreamResult *__restrict__ res,
                   const u8 *__restrict__ planeL, const u8 *__restrict__ planeH, u64 stride, u16 *__restrict__ maps) {
	const u32 t = blockIdx.x;
	if (res->status != 0 || t >= res->n_live) return;
	const u32 k = live[t];
	const u32 n = units[k].out_len;
	const u8 *L = planeL + (u64)k * stride + STREAM_WINDOW, *H = planeH + (u64)k * stride + STREAM_WINDOW;
	u16 *M = maps + (u64)t * STREAM_WINDOW;
	for (u32 i = blockIdx.y * (STREAM_WINDOW / gridDim.y) + threadIdx.x; i < (blockIdx.y + 1) * (STREAM_WINDOW / gridDim.y); i += blockDim.x) {
		u16 e;
		if (i + n >= STREAM_WINDOW) {                      // a byte of this unit
			const u32 idx = i + n - STREAM_WINDOW;
			const u32 l = L[idx], h = H[idx];
			e = (h & 0x80) ? (u16)(MAP_REF | (h & 0x7F) << 8 | l) : (u16)l;
		} else {
			e = (u16)(MAP_REF | (i + n));                  // the window it found, moved on by n
		}
		M[i] = e;
	}
}
__global__ void __launch_bounds__(256)
stream_compose_kernel(const u16 *__restrict__ src, u16 *__restrict__ dst, u32 d, const StreamResult *__restrict__ res) {
	const u32 t = blockIdx.x;
	if (res->status != 0 || t >= res->n_live) return;
	const u16 *A = src + (u64)t * STREAM_WINDOW, *B = t >= d ? src + (u64)(t - d) * STREAM_WINDOW : nullptr;
	u16 *C = dst + (u64)t * STREAM_WINDOW;
	for (u32 i = blockIdx.y * (STREAM_WINDOW / gridDim.y) + threadIdx.x; i < (blockIdx.y + 1) * (STREAM_WINDOW / gridDim.y); i += blockDim.x) {
		u16 e = A[i];
		if (B && (e & MAP_REF)) e = B[e & 0x7FFF];
		C[i] = e;
	}
}

// every live unit: markers resolved against the window its predecessor left, bytes to their place in the output
__global__ void __launch_bounds__(256)
stream_finish_kernel(const StreamUnit *__restrict__ units, const u32 *__restrict__ live, const u64 *__restrict__ off,
                     StreamResult *res, const u8 *__restrict__ planeL, const u8 *__restrict__ planeH, u64 stride,
                     const u16 *__restrict__ maps, u8 *__restrict__ out) {
	const u32 t = blockIdx.x;
	if (res->status != 0 || t >= res->n_live) return;
	const u32 k = live[t];
	const u32 n = units[k].out_len;
	const u64 o = off[t];
	const u8 *L = planeL + (u64)k * stride + STREAM_WINDOW, *H = planeH + (u64)k * stride + STREAM_WINDOW;
	const u16 *W = t ? maps + (u64)(t - 1) * STREAM_WINDOW : nullptr;     // the window in front of the unit (after the scan: bytes)
	bool bad = false;
	for (u32 i = (blockIdx.y * blockDim.x + threadIdx.x) * 4; i < n; i += gridDim.y * blockDim.x * 4) {
		const u32 l4 = *(const u32 *)(L + i), h4 = *(const u32 *)(H + i);      // the unit's buffers are 16-byte aligned, padded
		u32 v4 = l4;
		if (h4 & 0x80808080u) {
#pragma unroll
			for (int b = 0; b < 4; b++) {
				const u32 h = (h4 >> (8 * b)) & 0xFF;
				if (h & 0x80) {
					const u32 mk = (h & 0x7F) << 8 | ((l4 >> (8 * b)) & 0xFF);
					// window byte mk is stream position o - 32768 + mk: it must exist (Open.java:592-593); what is still
					// a reference after the scan points in front of the stream
					const u32 e = W ? W[mk] : MAP_REF;
					if (e & MAP_REF) { bad |= i + b < n; continue; }
					v4 = (v4 & ~(0xFFu << (8 * b))) | e << (8 * b);
				}
			}
		}
		if (i + 4 <= n && ((uintptr_t)(out + o + i) & 3) == 0) *(u32 *)(out + o + i) = v4;
		else for (u32 b = 0; b < 4 && i + b < n; b++) out[o + i + b] = (u8)(v4 >> (8 * b));
	}
	if (bad) atomicMax((unsigned long long *)&res->status, (unsigned long long)B2D_COPY_FROM_BEFORE_DICTIONARY_START);
}

struct StreamPlan {
	u32 seg_bytes, n_seg, n_units, cap, gcap;
	u64 stride;
	size_t o_units, o_planeL, o_planeH, o_glist, o_mapsA, o_mapsB, o_live, o_off, o_pool, total;
};
static StreamPlan stream_plan(uint64_t in_len, uint32_t cap_scale) {
	StreamPlan p;
	static const char *e_seg = getenv("B2D_STREAM_SEGMENT"), *e_cap = getenv("B2D_STREAM_UNIT_CAP");
	p.seg_bytes = e_seg && atoi(e_seg) >= 4096 ? (u32)atoi(e_seg) & ~3u : 32768u;
	p.stride = e_cap && atoll(e_cap) >= 65536 ? ((u64)atoll(e_cap) + STREAM_WINDOW + 255) & ~(u64)255 : (u64)16 * p.seg_bytes;
	p.stride *= cap_scale ? cap_scale : 1;
	p.cap = (u32)min((u64)0xFFFFF0, p.stride - STREAM_WINDOW - 64);
	p.gcap = p.cap / 3 + 8;
	p.n_seg = (u32)max((u64)1, (in_len + p.seg_bytes - 1) / p.seg_bytes);
	p.n_units = p.n_seg + max(16u, p.n_seg / 4);          // + spare units for stretches without a block start
	size_t o = 0;
	auto carve = [&](size_t bytes) { size_t r = o; o += (bytes + 255) & ~(size_t)255; return r; };
	p.o_units = carve((size_t)p.n_units * sizeof(StreamUnit));
	p.o_planeL = carve((size_t)p.n_units * p.stride + 256);
	p.o_planeH = carve((size_t)p.n_units * p.stride + 256);
	p.o_glist = carve((size_t)p.n_units * p.gcap * 8);
	p.o_mapsA = carve((size_t)p.n_units * STREAM_WINDOW * 2);
	p.o_mapsB = carve((size_t)p.n_units * STREAM_WINDOW * 2);
	p.o_live = carve((size_t)p.n_units * 4);
	p.o_off = carve((size_t)(p.n_units + 1) * 8);
	p.o_pool = carve(256);
	p.total = o;
	return p;
}
size_t inflate_stream_scratch_bytes(uint64_t in_len, uint32_t cap_scale) { return stream_plan(in_len, cap_scale).total; }

cudaError_t launch_inflate_stream(const u8 *d_in, u64 in_len, u8 *d_out, u64 out_cap, void *d_result, void *d_scratch, cudaStream_t st,
                                  uint32_t cap_scale) {
	static bool attr_set[MAX_DEVICES] = {};
	const int slot = current_device_slot();
	cudaError_t e;
	if (!attr_set[slot]) {
		if ((e = cudaFuncSetAttribute(stream_find_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return e;
		if ((e = cudaFuncSetAttribute(stream_units_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return e;
		attr_set[slot] = true;
	}
	const StreamPlan p = stream_plan(in_len, cap_scale);
	u8 *sp = (u8 *)d_scratch;
	StreamUnit *units = (StreamUnit *)(sp + p.o_units);
	u8 *planeL = sp + p.o_planeL, *planeH = sp + p.o_planeH;
	u64 *glist = (u64 *)(sp + p.o_glist);
	u16 *mapsA = (u16 *)(sp + p.o_mapsA), *mapsB = (u16 *)(sp + p.o_mapsB);
	u32 *live = (u32 *)(sp + p.o_live);
	u64 *off = (u64 *)(sp + p.o_off);
	StreamResult *res = (StreamResult *)d_result;
	u32 *pool = (u32 *)(sp + p.o_pool);
	const u32 wgrid = (p.n_seg + WARPS_PER_CTA - 1) / WARPS_PER_CTA, ugrid = (p.n_units + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
	B2D_LAUNCH(stream_find_kernel, ugrid, WARPS_PER_CTA * 32, 0, st)(d_in, in_len, p.seg_bytes, p.n_seg, p.n_units, units, pool);
	B2D_LAUNCH(stream_units_kernel, wgrid, WARPS_PER_CTA * 32, 0, st)(d_in, in_len, p.n_seg, p.n_units, units, planeL, p.stride, p.cap, glist, p.gcap, pool);
	B2D_LAUNCH(stream_planes_kernel, p.n_units, 256, 0, st)(units, planeL, planeH, p.stride);
	B2D_LAUNCH(stream_replay_kernel, dim3(ugrid, 2), WARPS_PER_CTA * 32, 0, st)(units, p.n_units, planeL, planeH, p.stride, glist, p.gcap);
	B2D_LAUNCH(stream_walk_kernel, 1, 32, 0, st)(units, p.n_units, live, off, out_cap, res);
	B2D_LAUNCH(stream_maps_kernel, dim3(p.n_units, 4), 256, 0, st)(units, live, res, planeL, planeH, p.stride, mapsA);
	u16 *src = mapsA, *dst = mapsB;
	for (u32 d = 1; d < p.n_units; d <<= 1) {          // (the number of live units is only known on the device: enough rounds for all)
		B2D_LAUNCH(stream_compose_kernel, dim3(p.n_units, 4), 256, 0, st)(src, dst, d, res);
		u16 *tmp = src; src = dst; dst = tmp;
	}
	B2D_LAUNCH(stream_finish_kernel, dim3(p.n_units, 4), 256, 0, st)(units, live, off, res, planeL, planeH, p.stride, src, d_out);
	return cudaGetLastError();
}

size_t inflate_units_scratch_bytes(uint64_t out_total, uint32_t chunk_bytes, uint32_t block_bytes) {
	const u64 n_chunks = (out_total + chunk_bytes - 1) / chunk_bytes;
	const u64 n_units = n_chunks * (chunk_bytes / block_bytes);
	const u64 gcap = block_bytes / 3 + 8;
	return (size_t)(n_units * gcap * 8 + n_units * 8 + 1024);
}

cudaError_t launch_inflate_units(const u8 *d_in, const u64 *d_chunk_in_off, u32 n_chunks, const u32 *d_block_bits,
                                 u32 chunk_bytes, u32 block_bytes, u64 out_total, u8 *d_out, int *d_chunk_status,
                                 void *d_scratch, cudaStream_t st) {
	if (n_chunks == 0) return cudaSuccess;
	static bool attr_set[MAX_DEVICES] = {};
	const int slot = current_device_slot();
	if (!attr_set[slot]) {
		cudaError_t e = cudaFuncSetAttribute(inflate_units_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
		                                     cudaSharedmemCarveoutMaxShared);
		if (e != cudaSuccess) return e;
		attr_set[slot] = true;
	}
	const u32 bpc = chunk_bytes / block_bytes;
	const u32 n_units = n_chunks * bpc;
	const u32 gcap = block_bytes / 3 + 8;
	u64 *glist = (u64 *)d_scratch;
	u32 *gcount = (u32 *)(glist + (u64)n_units * gcap);
	int *ustatus = (int *)(gcount + n_units);
	B2D_LAUNCH(inflate_units_kernel, (n_units + WARPS_PER_CTA - 1) / WARPS_PER_CTA, WARPS_PER_CTA * 32, 0, st)(
		d_in, d_chunk_in_off, d_block_bits, n_units, bpc, chunk_bytes, block_bytes, out_total, d_out, glist, gcap, gcount, ustatus);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return e;
	B2D_LAUNCH(resolve_units_kernel, (n_chunks + WARPS_PER_CTA - 1) / WARPS_PER_CTA, WARPS_PER_CTA * 32, 0, st)(
		d_out, glist, gcap, gcount, ustatus, n_chunks, bpc, chunk_bytes, block_bytes, out_total, d_chunk_status);
	return cudaGetLastError();
}

// The static shared segment must start at SM_WINDOW_BASE for the LUT addressing above; this kernel has the decoder's
// declaration and reports where it landed.
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, CTAS_PER_SM) smem_base_probe_kernel(u32 *out) {
	__shared__ __align__(1024) u8 smem_raw[SM_BYTES];
	smem_raw[threadIdx.x] = (u8)threadIdx.x;          // keep the array alive
	__syncthreads();
	if (threadIdx.x == 0) out[0] = (u32)__cvta_generic_to_shared(smem_raw) + (smem_raw[1] == 1 ? 0u : 1u);
}
cudaError_t probe_inflate_smem_base(uint32_t *base_out, cudaStream_t st) {
	u32 *d = nullptr;
	cudaError_t e = cudaMalloc(&d, 4);
	if (e != cudaSuccess) return e;
	B2D_LAUNCH(smem_base_probe_kernel, 1, WARPS_PER_CTA * 32, 0, st)(d);
	e = cudaGetLastError();
	if (e == cudaSuccess) e = cudaMemcpyAsync(base_out, d, 4, cudaMemcpyDeviceToHost, st);
	if (e == cudaSuccess) e = cudaStreamSynchronize(st);
	cudaFree(d);
	return e;
}

cudaError_t launch_inflate(const u8 *d_in, const u64 *d_in_off, u32 n, u8 *d_out, const u64 *d_out_off,
                           u64 *d_out_len, u64 *d_in_consumed, int *d_status, u32 flags, cudaStream_t st,
                           uint8_t *out_mirror, const u64 *d_in_end, uint32_t *progress, const uint8_t *in_host,
                           const uint32_t *landed, uint32_t group_size) {
	if (n == 0) return cudaSuccess;
	static bool attr_set[MAX_DEVICES] = {};
	const int slot = current_device_slot();
	if (!attr_set[slot]) {
		cudaError_t e = cudaFuncSetAttribute(inflate_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
		if (e == cudaSuccess) e = cudaFuncSetAttribute(inflate_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
		if (e != cudaSuccess) return e;
		attr_set[slot] = true;
	}
	u32 grid = (n + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
	const long long mdelta = progress ? MDELTA_PROGRESS : out_mirror ? (long long)(out_mirror - d_out) : 0ll;
	if (in_host)
		B2D_LAUNCH(inflate_kernel<true>, grid, WARPS_PER_CTA * 32, 0, st)(d_in, d_in_off, d_in_end, n, d_out, d_out_off, d_out_len, d_in_consumed,
		                                                                  d_status, flags, mdelta, progress, in_host, landed, group_size ? group_size : 1u);
	else
		B2D_LAUNCH(inflate_kernel<false>, grid, WARPS_PER_CTA * 32, 0, st)(d_in, d_in_off, d_in_end, n, d_out, d_out_off, d_out_len, d_in_consumed,
		                                                                   d_status, flags, mdelta, progress, nullptr, nullptr, 1u);
	return cudaGetLastError();
}

}  // namespace b2d

#ifdef B2D_LOOPBENCH
extern "C" __attribute__((visibility("default"))) int b2d_loop_bench(int mode, int rounds, uint32_t *out4) {
	return (int)b2d::run_loop_bench(mode, rounds, out4);
}
#endif
