// crc32.cu -- CRC-32 (IEEE 802.3, reflected 0xEDB88320) on sm_100a.
//
// Replaces java.util.zip.CRC32 at the reference's call sites GzipOutputStream.java:25,57,67 and
// GzipInputStream.java:32,72,83 (the reference itself has no CRC code; it uses the JDK's).
// One CTA per segment (a gzip member's output, or a 1 MiB chunk of deflate input).  Every thread folds
// its own run of 16-byte vectors with slice-by-4 tables held in shared memory, then the per-thread CRCs
// are combined in a log-depth tree with carry-less multiplications by x^(8*bytes) mod P -- the
// crc32_combine identity crc(A||B) = crc(A)*x^(8|B|) + crc(B) -- so no byte is read twice.
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

__device__ __forceinline__ u32 crc_byte(const u32 *T0, u32 c, u32 byte) {
	return T0[(c ^ byte) & 0xFF] ^ (c >> 8);
}
__device__ __forceinline__ u32 crc_word(const u32 *T, u32 c, u32 w) {   // T[k*256 + v] = slice-by-4 tables
	c ^= w;
	return T[768 + (c & 0xFF)] ^ T[512 + ((c >> 8) & 0xFF)] ^ T[256 + ((c >> 16) & 0xFF)] ^ T[c >> 24];
}

__global__ void __launch_bounds__(CRC_THREADS)
crc32_kernel(const u8 *__restrict__ data, const u64 *__restrict__ off, const u64 *__restrict__ len,
             u64 total, u64 piece, u32 *__restrict__ crc_out) {
	__shared__ u32 T[1024];
	__shared__ u32 part[CRC_THREADS];
	__shared__ u32 m0_sh;
	const int t = threadIdx.x;
	{   // slice-by-4 tables
		u32 c = (u32)t;
		for (int k = 0; k < 8; k++) c = (c >> 1) ^ (POLY & (0u - (c & 1u)));
		T[t] = c;
		__syncthreads();
		u32 v = c;
		for (int k = 1; k < 4; k++) { v = T[v & 0xFF] ^ (v >> 8); T[k * 256 + t] = v; }
	}
	u64 s_off, s_len;
	if (off) { s_off = off[blockIdx.x]; s_len = len[blockIdx.x]; }
	else { s_off = (u64)blockIdx.x * piece; s_len = total - s_off < piece ? total - s_off : piece; }
	const u8 *s = data + s_off, *e = s + s_len;
	const u8 *A = (const u8 *)(((uintptr_t)s + 15) & ~(uintptr_t)15);
	const u8 *B = (const u8 *)((uintptr_t)e & ~(uintptr_t)15);
	__syncthreads();
	if (B <= A) {                                   // tiny segment: one thread, bytewise
		if (t == 0) {
			u32 c = 0xFFFFFFFFu;
			for (const u8 *p = s; p < e; p++) c = crc_byte(T, c, *p);
			crc_out[blockIdx.x] = ~c;
		}
		return;
	}
	const u64 nv = (u64)(B - A) >> 4;                // 16-byte vectors in the aligned middle
	const u64 L = (nv + CRC_THREADS - 1) / CRC_THREADS;
	// pieces are right-aligned: thread t owns vectors [nv - (T-t)L, nv - (T-1-t)L) clipped at 0, so every
	// non-empty piece except the left-most is full and the combine multipliers are uniform per tree level
	long long lo = (long long)nv - (long long)(CRC_THREADS - t) * (long long)L;
	long long hi = lo + (long long)L;
	const int t0 = CRC_THREADS - (int)((nv + L - 1) / L);   // left-most non-empty thread
	if (lo < 0) lo = 0;
	u32 c = 0xFFFFFFFFu;
	if (t == t0) for (const u8 *p = s; p < A; p++) c = crc_byte(T, c, *p);   // unaligned head joins the first piece
	if (hi > lo) {
		const uint4 *v = (const uint4 *)A + lo;
		for (long long k = 0; k < hi - lo; k++) {
			uint4 w = __ldg(v + k);
			c = crc_word(T, c, w.x);
			c = crc_word(T, c, w.y);
			c = crc_word(T, c, w.z);
			c = crc_word(T, c, w.w);
		}
	}
	part[t] = (t >= t0) ? ~c : 0u;                   // crc of an empty piece is 0
	if (t == 0) m0_sh = dev_x2nmodp(L * 16, 3);      // x^(8 * bytes per piece)
	__syncthreads();
	u32 M = m0_sh;
	for (int step = 1; step < CRC_THREADS; step <<= 1) {
		u32 mine = 0;
		bool act = (t & (2 * step - 1)) == (2 * step - 1);
		if (act) mine = multmodp(M, part[t - step]) ^ part[t];
		__syncthreads();
		if (act) part[t] = mine;
		M = multmodp(M, M);
		__syncthreads();
	}
	if (t == CRC_THREADS - 1) {
		u32 cc = ~part[t];
		for (const u8 *p = B; p < e; p++) cc = crc_byte(T, cc, *p);          // unaligned tail continues the state
		crc_out[blockIdx.x] = ~cc;
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

cudaError_t launch_crc32_segments(const uint8_t *d_data, const uint64_t *d_off, const uint64_t *d_len,
                                  uint32_t n_seg, uint32_t *d_crc, cudaStream_t st) {
	if (n_seg == 0) return cudaSuccess;
	cudaError_t e = upload_tables();
	if (e != cudaSuccess) return e;
	B2D_LAUNCH(crc32_kernel, n_seg, CRC_THREADS, 0, st)(d_data, d_off, d_len, 0, 0, d_crc);
	return cudaGetLastError();
}

cudaError_t launch_crc32_pieces(const uint8_t *d_data, uint64_t total, uint64_t piece, uint32_t n_pieces,
                                uint32_t *d_crc, cudaStream_t st) {
	if (n_pieces == 0) return cudaSuccess;
	cudaError_t e = upload_tables();
	if (e != cudaSuccess) return e;
	B2D_LAUNCH(crc32_kernel, n_pieces, CRC_THREADS, 0, st)(d_data, nullptr, nullptr, total, piece, d_crc);
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
