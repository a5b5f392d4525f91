// crc32.cu -- CRC-32 (IEEE 802.3, reflected 0xEDB88320) on sm_100a.
//
// Replaces java.util.zip.CRC32 at the reference's call sites GzipOutputStream.java:25,57,67 and
// GzipInputStream.java:32,72,83 (the reference itself has no CRC code; it uses the JDK's).
// One warp per segment (a gzip member's output, or a 1 MiB chunk of deflate input), in persistent CTAs.  Every lane
// folds its own run of 16-byte vectors with slice-by-4 tables held in shared memory -- a copy per lane, so that no
// lookup conflicts with another -- then the 32 lane CRCs are combined in a log-depth tree with carry-less
// multiplications by x^(8*bytes) mod P -- the crc32_combine identity crc(A||B) = crc(A)*x^(8|B|) + crc(B) -- so no
// byte is read twice.
#include "common.cuh"
#include "kernels.h"

namespace b2d {

constexpr u32 POLY = 0xEDB88320u;
constexpr int CRC_THREADS = 256;

__constant__ u32 X2N[32];          // x^(2^n) mod P, n = 0..31 (reflected; bit 31 = x^0)
static u32 h_x2n[32];
static bool h_ready = false;

__host__ __device__ inline u32 multmodp(u32 a, u32 b) {      // a * b mod P over GF(2)
	u32 p = 0;
#pragma unroll 4
	for (int i = 31; i >= 0; i--) {
		p ^= b & (0u - ((a >> i) & 1u));
		b = (b >> 1) ^ (POLY & (0u - (b & 1u)));
	}
	return p;
}

static void host_init_tables() {
	if (h_ready) return;
	u32 p = 1u << 30;                // x^1
	h_x2n[0] = p;
	for (int n = 1; n < 32; n++) h_x2n[n] = p = multmodp(p, p);
	h_ready = true;
}

static u32 host_x2nmodp(u64 n, unsigned k) {                  // x^(n * 2^k) mod P
	host_init_tables();
	u32 p = 1u << 31;
	while (n) {
		if (n & 1) p = multmodp(h_x2n[k & 31], p);
		n >>= 1;
		k++;
	}
	return p;
}

uint32_t host_crc32_combine(uint32_t a, uint32_t b, uint64_t len_b) {
	return multmodp(host_x2nmodp(len_b, 3), a) ^ b;
}

uint32_t host_crc32_bytes(uint32_t crc, const uint8_t *p, size_t n) {
	u32 c = ~crc;
	for (size_t i = 0; i < n; i++) {
		c ^= p[i];
		for (int k = 0; k < 8; k++) c = (c >> 1) ^ (POLY & (0u - (c & 1u)));
	}
	return ~c;
}

__device__ inline u32 dev_x2nmodp(u64 n, unsigned k) {
	u32 p = 1u << 31;
	while (n) {
		if (n & 1) p = multmodp(X2N[k & 31], p);
		n >>= 1;
		k++;
	}
	return p;
}

// Slice-by-4 tables, one copy per LANE: TR[(k * 256 + v) * 32 + lane].  With one copy per CTA the 32 lanes of a warp hit
// the 32 banks at random, 3.2 wavefronts per lookup (ncu: 67 % of the kernel's shared wavefronts were conflicts, stall
// mio_throttle); here lane l only ever touches bank l.
constexpr int CRC_WARPS = 32;
constexpr int CRC_SMEM = 4 * 256 * 32 * 4;            // 128 KiB
__device__ __forceinline__ u32 tr_at(u32 tr_lane, u32 k, u32 v) {   // tr_lane = shared address of TR + lane * 4
	u32 r;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(tr_lane + ((k * 256 + v) << 7)));
	return r;
}
__device__ __forceinline__ u32 crc_byte(u32 tr_lane, u32 c, u32 byte) {
	return tr_at(tr_lane, 0, (c ^ byte) & 0xFF) ^ (c >> 8);
}
__device__ __forceinline__ u32 crc_word(u32 tr_lane, u32 c, u32 w) {
	c ^= w;
	return tr_at(tr_lane, 3, c & 0xFF) ^ tr_at(tr_lane, 2, (c >> 8) & 0xFF) ^ tr_at(tr_lane, 1, (c >> 16) & 0xFF) ^ tr_at(tr_lane, 0, c >> 24);
}

// One CTA per SM, one WARP per part of a segment at a time (`sub` parts per segment, a power of two: few large
// segments are shared by several warps so that the machine is full): every lane folds its own run of 16-byte vectors,
// the 32 lane CRCs are combined in a five-step tree with multiplications by x^(8 * bytes) mod P, the parts' CRCs by
// the first part's warp.
__global__ void __launch_bounds__(CRC_WARPS * 32, 1)
crc32_kernel(const u8 *__restrict__ data, const u64 *__restrict__ off, const u64 *__restrict__ len,
             u64 total, u64 piece, u32 n_seg, u32 sub, u32 *__restrict__ crc_out) {
	extern __shared__ __align__(16) u32 TR[];
	__shared__ u32 T0[1024];
	__shared__ u32 part_crc[CRC_WARPS];
	__shared__ u64 part_len[CRC_WARPS];
	const u32 t = threadIdx.x, lane = t & 31, warp = t >> 5;
	{   // slice-by-4 tables, then a copy per lane (written bank by bank)
		if (t < 256) {
			u32 c = t;
			for (int k = 0; k < 8; k++) c = (c >> 1) ^ (POLY & (0u - (c & 1u)));
			T0[t] = c;
		}
		__syncthreads();
		if (t < 256) {
			u32 v = T0[t];
			for (u32 k = 1; k < 4; k++) { v = T0[v & 0xFF] ^ (v >> 8); T0[k * 256 + t] = v; }
		}
		__syncthreads();
		for (u32 i = t; i < 32768; i += CRC_WARPS * 32) TR[i] = T0[i >> 5];
		__syncthreads();
	}
	const u32 tr_lane = (u32)__cvta_generic_to_shared(TR) + lane * 4;
	u64 last_L = ~0ull;
	u32 M0 = 0;
	const u64 n_items = (u64)n_seg * sub;
	for (u64 base = (u64)blockIdx.x * CRC_WARPS; base < n_items; base += (u64)gridDim.x * CRC_WARPS) {   // (uniform per CTA)
		const u64 item = base + warp;
		u32 crc = 0;
		u64 my_len = 0;
		const u32 seg = (u32)(item / sub), part_no = (u32)(item % sub);
		if (item < n_items) {
			u64 s_off, s_len;
			if (off) { s_off = off[seg]; s_len = len[seg]; }
			else { s_off = (u64)seg * piece; s_len = total - s_off < piece ? total - s_off : piece; }
			const u64 plen = (((s_len + sub - 1) / sub) + 15) & ~(u64)15;          // bytes per part
			const u64 pa = min(s_len, (u64)part_no * plen), pb = min(s_len, pa + plen);
			my_len = pb - pa;
			const u8 *s = data + s_off + pa, *e = data + s_off + pb;
			const u8 *A = (const u8 *)(((uintptr_t)s + 15) & ~(uintptr_t)15);
			const u8 *B = (const u8 *)((uintptr_t)e & ~(uintptr_t)15);
			if (B <= A) {                                   // tiny: one lane, bytewise (every lane: the same result)
				u32 c = 0xFFFFFFFFu;
				for (const u8 *p = s; p < e; p++) c = crc_byte(tr_lane, c, *p);
				crc = my_len ? ~c : 0u;
			} else {
				const u64 nv = (u64)(B - A) >> 4;                // 16-byte vectors in the aligned middle
				const u64 L = (nv + 31) / 32;
				// pieces are right-aligned: lane l owns vectors [nv - (32 - l) L, nv - (31 - l) L) clipped at 0, so every
				// non-empty piece except the left-most is full and the combine multipliers are uniform per tree level
				long long lo = (long long)nv - (long long)(32 - lane) * (long long)L;
				const long long hi = lo + (long long)L;
				const u32 l0 = 32 - (u32)((nv + L - 1) / L);     // left-most non-empty lane
				if (lo < 0) lo = 0;
				u32 c = 0xFFFFFFFFu;
				if (lane == l0) for (const u8 *p = s; p < A; p++) c = crc_byte(tr_lane, c, *p);   // unaligned head joins the first piece
				if (hi > lo) {
					const uint4 *v = (const uint4 *)A + lo;
					const long long n = hi - lo;
					uint4 w0 = __ldg(v), w1 = n > 1 ? __ldg(v + 1) : make_uint4(0, 0, 0, 0);
					long long k = 0;
					for (; k + 2 <= n; k += 2) {                 // a whole 32-byte sector per trip, the next one on its way
						const uint4 a0 = w0, a1 = w1;
						if (k + 2 < n) w0 = __ldg(v + k + 2);
						if (k + 3 < n) w1 = __ldg(v + k + 3);
						c = crc_word(tr_lane, c, a0.x); c = crc_word(tr_lane, c, a0.y); c = crc_word(tr_lane, c, a0.z); c = crc_word(tr_lane, c, a0.w);
						c = crc_word(tr_lane, c, a1.x); c = crc_word(tr_lane, c, a1.y); c = crc_word(tr_lane, c, a1.z); c = crc_word(tr_lane, c, a1.w);
					}
					if (k < n) {
						c = crc_word(tr_lane, c, w0.x); c = crc_word(tr_lane, c, w0.y); c = crc_word(tr_lane, c, w0.z); c = crc_word(tr_lane, c, w0.w);
					}
				}
				u32 part = (lane >= l0) ? ~c : 0u;              // crc of an empty piece is 0
				if (L != last_L) { M0 = dev_x2nmodp(L * 16, 3); last_L = L; }   // x^(8 * bytes per piece)
				u32 M = M0;
#pragma unroll 1
				for (u32 step = 1; step < 32; step <<= 1) {
					const u32 left = __shfl_up_sync(0xFFFFFFFFu, part, step);
					if ((lane & (2 * step - 1)) == (2 * step - 1)) part = multmodp(M, left) ^ part;
					M = multmodp(M, M);
				}
				u32 cc = ~__shfl_sync(0xFFFFFFFFu, part, 31);
				for (const u8 *p = B; p < e; p++) cc = crc_byte(tr_lane, cc, *p);      // unaligned tail continues the state
				crc = ~cc;
			}
		}
		if (sub == 1) {
			if (item < n_items && lane == 0) crc_out[seg] = crc;
			continue;
		}
		// the parts of a segment sit in neighbouring warps of this CTA: the first one's warp folds them left to right
		if (lane == 0) { part_crc[warp] = crc; part_len[warp] = my_len; }
		__syncthreads();
		if (item < n_items && part_no == 0 && lane == 0) {
			u32 acc = part_crc[warp];
			for (u32 j = 1; j < sub; j++)
				if (part_len[warp + j]) acc = multmodp(dev_x2nmodp(part_len[warp + j], 3), acc) ^ part_crc[warp + j];
			crc_out[seg] = acc;
		}
		__syncthreads();
	}
}

// crc(A||B) = crc(A) * x^(8|B|) + crc(B): one thread folds the piece CRCs left to right (n_pieces is small)
__global__ void crc32_fold_kernel(const u32 *__restrict__ piece_crc, u32 n_pieces, u64 piece, u64 total, u32 *__restrict__ out) {
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	if (n_pieces == 0) { *out = 0; return; }
	u32 crc = piece_crc[0];
	const u32 M = dev_x2nmodp(piece, 3);
	for (u32 i = 1; i < n_pieces; i++) {
		const u64 len = (i + 1 == n_pieces) ? total - (u64)i * piece : piece;
		const u32 m = len == piece ? M : dev_x2nmodp(len, 3);
		crc = multmodp(m, crc) ^ piece_crc[i];
	}
	*out = crc;
}

static cudaError_t upload_tables() {
	static bool uploaded[MAX_DEVICES] = {};             // __constant__ memory is per device
	const int slot = current_device_slot();
	if (uploaded[slot]) return cudaSuccess;
	host_init_tables();
	cudaError_t e = cudaMemcpyToSymbol(X2N, h_x2n, sizeof(h_x2n));
	if (e == cudaSuccess) uploaded[slot] = true;
	return e;
}

// one CTA per SM (fewer when there are fewer than 32 segments per SM); the kernel's 128 KiB of shared memory is opted
// into once per device
static cudaError_t crc32_grid(uint32_t n_seg, int *grid, uint32_t *sub) {
	static int sms[MAX_DEVICES] = {};
	const int slot = current_device_slot();
	if (!sms[slot]) {
		int dev = 0, n = 0;
		cudaError_t e = cudaGetDevice(&dev);
		if (e == cudaSuccess) e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
		if (e == cudaSuccess) e = cudaFuncSetAttribute(crc32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CRC_SMEM);
		if (e != cudaSuccess) return e;
		sms[slot] = n > 0 ? n : 1;
	}
	// few segments are shared by several warps each (a power of two, so that a segment's parts sit in one CTA)
	uint32_t s = 1;
	while (s < 32 && (uint64_t)n_seg * s < (uint64_t)sms[slot] * CRC_WARPS / 2) s <<= 1;
	*sub = s;
	const uint64_t want = ((uint64_t)n_seg * s + CRC_WARPS - 1) / CRC_WARPS;
	*grid = want < (uint64_t)sms[slot] ? (int)want : sms[slot];
	return cudaSuccess;
}

cudaError_t launch_crc32_segments(const uint8_t *d_data, const uint64_t *d_off, const uint64_t *d_len,
                                  uint32_t n_seg, uint32_t *d_crc, cudaStream_t st) {
	if (n_seg == 0) return cudaSuccess;
	cudaError_t e = upload_tables();
	if (e != cudaSuccess) return e;
	int grid = 0;
	uint32_t sub = 1;
	if ((e = crc32_grid(n_seg, &grid, &sub)) != cudaSuccess) return e;
	B2D_LAUNCH(crc32_kernel, grid, CRC_WARPS * 32, CRC_SMEM, st)(d_data, d_off, d_len, 0, 0, n_seg, sub, d_crc);
	return cudaGetLastError();
}

cudaError_t launch_crc32_pieces(const uint8_t *d_data, uint64_t total, uint64_t piece, uint32_t n_pieces,
                                uint32_t *d_crc, cudaStream_t st) {
	if (n_pieces == 0) return cudaSuccess;
	cudaError_t e = upload_tables();
	if (e != cudaSuccess) return e;
	int grid = 0;
	uint32_t sub = 1;
	if ((e = crc32_grid(n_pieces, &grid, &sub)) != cudaSuccess) return e;
	B2D_LAUNCH(crc32_kernel, grid, CRC_WARPS * 32, CRC_SMEM, st)(d_data, nullptr, nullptr, total, piece, n_pieces, sub, d_crc);
	return cudaGetLastError();
}

cudaError_t launch_crc32_fold(const uint32_t *d_piece_crc, uint32_t n_pieces, uint64_t piece, uint64_t total,
                              uint32_t *d_crc_out, cudaStream_t st) {
	cudaError_t e = upload_tables();
	if (e != cudaSuccess) return e;
	B2D_LAUNCH(crc32_fold_kernel, 1, 32, 0, st)(d_piece_crc, n_pieces, piece, total, d_crc_out);
	return cudaGetLastError();
}

// ---------------------------------------------------------------- Adler-32 (zlib container)
// Replaces java.util.zip.Adler32 at ZlibOutputStream.java:25,56,65 and ZlibInputStream.java:30,69,78.  For a piece d[0,n):
// A = sum d[i], B = sum (n - i) d[i]; pieces concatenate as A = Ax + Ay, B = Bx + ny * Ax + By, and a running state
// (s1, s2) advances as s1' = s1 + A, s2' = s2 + n * s1 + B  (all mod 65521).  One CTA per piece, one strip per thread,
// thread 0 folds the 256 strips.
constexpr u32 ADLER_MOD = 65521u;

__global__ void __launch_bounds__(CRC_THREADS)
adler32_kernel(const u8 *__restrict__ data, const u64 *__restrict__ off, const u64 *__restrict__ len,
               u64 total, u64 piece, u32 *__restrict__ out) {
	__shared__ u32 sa[CRC_THREADS], sb[CRC_THREADS];
	const int t = threadIdx.x;
	u64 s_off, s_len;
	if (off) { s_off = off[blockIdx.x]; s_len = len[blockIdx.x]; }
	else { s_off = (u64)blockIdx.x * piece; s_len = total - s_off < piece ? total - s_off : piece; }
	const u64 L = (((s_len + CRC_THREADS - 1) / CRC_THREADS) + 15) & ~(u64)15;      // strip length, multiple of 16
	const u64 lo = min(s_len, (u64)t * L), hi = min(s_len, lo + L);
	const u8 *p = data + s_off + lo;
	u64 A = 0, B = 0;                                   // running sums of this strip, reduced mod 65521 per block
	u64 done = 0;
	const u64 n = hi - lo;
	while (done < n) {
		const u32 blk = (u32)min((u64)2048, n - done);
		u32 a = 0, b = 0;                                // b = sum (blk - i) * d[i]  <= 255 * 2048 * 2049 / 2 < 2^32
		u32 i = 0;
		if ((((uintptr_t)(p + done)) & 15) == 0) {
			for (; i + 16 <= blk; i += 16) {
				const uint4 v = __ldg((const uint4 *)(p + done + i));
				const u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
				for (int k = 0; k < 4; k++) {
#pragma unroll
					for (int q = 0; q < 4; q++) {
						const u32 d = (w[k] >> (8 * q)) & 0xFF;
						a += d;
						b += (blk - (i + 4 * k + q)) * d;
					}
				}
			}
		}
		for (; i < blk; i++) { const u32 d = p[done + i]; a += d; b += (blk - i) * d; }
		// append the block: strip = strip || block
		B = (B + (u64)blk * A + b) % ADLER_MOD;
		A = (A + a) % ADLER_MOD;
		done += blk;
	}
	sa[t] = (u32)A;
	sb[t] = (u32)B;
	__syncthreads();
	if (t == 0) {
		u64 a = 0, b = 0;
		for (int k = 0; k < CRC_THREADS; k++) {
			const u64 klo = min(s_len, (u64)k * L), khi = min(s_len, klo + L);
			b = (b + ((khi - klo) % ADLER_MOD) * a + sb[k]) % ADLER_MOD;
			a = (a + sa[k]) % ADLER_MOD;
		}
		// as an Adler-32 with the standard start (s1 = 1, s2 = 0)
		const u64 s1 = (1 + a) % ADLER_MOD, s2 = ((s_len % ADLER_MOD) + b) % ADLER_MOD;
		out[blockIdx.x] = (u32)(s2 << 16 | s1);
	}
}

uint32_t host_adler32_combine(uint32_t ad1, uint32_t ad2, uint64_t len2) {    // adler(A||B) from adler(A), adler(B), |B|
	const uint64_t M = ADLER_MOD;
	const uint64_t a1 = ad1 & 0xFFFF, b1 = ad1 >> 16, a2 = ad2 & 0xFFFF, b2 = ad2 >> 16;
	const uint64_t rem = len2 % M;
	const uint64_t s1 = (a1 + a2 + M - 1) % M;
	const uint64_t s2 = (b1 + b2 + rem * ((a1 + M - 1) % M)) % M;
	return (uint32_t)(s2 << 16 | s1);
}

cudaError_t launch_adler32_segments(const uint8_t *d_data, const uint64_t *d_off, const uint64_t *d_len,
                                    uint32_t n_seg, uint32_t *d_out, cudaStream_t st) {
	if (n_seg == 0) return cudaSuccess;
	B2D_LAUNCH(adler32_kernel, n_seg, CRC_THREADS, 0, st)(d_data, d_off, d_len, 0, 0, d_out);
	return cudaGetLastError();
}

cudaError_t launch_adler32_pieces(const uint8_t *d_data, uint64_t total, uint64_t piece, uint32_t n_pieces,
                                  uint32_t *d_out, cudaStream_t st) {
	if (n_pieces == 0) return cudaSuccess;
	B2D_LAUNCH(adler32_kernel, n_pieces, CRC_THREADS, 0, st)(d_data, nullptr, nullptr, total, piece, d_out);
	return cudaGetLastError();
}

}  // namespace b2d
