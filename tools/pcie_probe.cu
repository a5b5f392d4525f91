// pcie_probe.cu -- how fast can SM stores reach mapped (pinned) host memory, by store shape?  Diagnostic for the
// host-mirror path of b2d_inflate_batch.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bin/pcie_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// every warp owns a contiguous region (like a member's output slot) and writes it front to back in `tile`-byte pieces,
// one 16-byte vector per lane and step
__global__ void write_regions(const uint4 *__restrict__ src, uint4 *dst, size_t region16, int also_dev, uint4 *dev) {
	const size_t w = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	const unsigned lane = threadIdx.x & 31;
	const size_t base = w * region16;
	for (size_t i = lane; i < region16; i += 32) {
		uint4 v = src[base + i];
		dst[base + i] = v;
		if (also_dev) dev[base + i] = v;
	}
}
// same, 4 bytes per lane
__global__ void write_regions_u32(const uint32_t *__restrict__ src, uint32_t *dst, size_t region4) {
	const size_t w = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	const unsigned lane = threadIdx.x & 31;
	const size_t base = w * region4;
	for (size_t i = lane; i < region4; i += 32) dst[base + i] = src[base + i];
}
// the decoder's flush pattern: pieces of ~900 bytes at arbitrary byte alignment, head and tail by byte stores, the
// 16-byte aligned body by vectors (store_tile in csrc/inflate.cu); align128 = 1 writes only whole 128-byte lines and
// carries the rest over to the next piece
__global__ void write_pieces(const uint8_t *__restrict__ src, uint8_t *dst, size_t region, int align128) {
	const size_t w = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	const unsigned lane = threadIdx.x & 31;
	const uint8_t *s = src + w * region;
	uint8_t *d = dst + w * region;
	size_t pos = 0, done = 0;
	unsigned seed = (unsigned)w * 2654435761u + 12345u;
	while (pos < region) {
		seed = seed * 1664525u + 1013904223u;
		size_t len = 800 + (seed >> 24);                 // 800..1055
		if (pos + len > region) len = region - pos;
		pos += len;
		size_t lo = done, hi = align128 && pos < region ? (pos & ~(size_t)127) : pos;
		if (hi <= lo) continue;
		size_t a = (lo + 15) & ~(size_t)15, b = hi & ~(size_t)15;
		if (a >= b) {
			for (size_t k = lo + lane; k < hi; k += 32) d[k] = s[k];
		} else {
			if (lo + lane < a) d[lo + lane] = s[lo + lane];
			for (size_t v = (a >> 4) + lane; v < (b >> 4); v += 32) ((uint4 *)d)[v] = ((const uint4 *)s)[v];
			if (b + lane < hi) d[b + lane] = s[b + lane];
		}
		done = hi;
	}
}
// The other direction: every warp owns a contiguous region of HOST memory (a member's compressed bytes) and pulls it
// into device memory front to back, `vecs` 16-byte vectors per lane and fetch (vecs * 512 bytes per warp and fetch),
// all loads of a fetch issued before the first store; `delay` clock ticks of busy work between fetches stand for the
// decoding that consumes the bytes.
__global__ void read_regions(const uint4 *__restrict__ host, uint4 *dev, size_t region16, int vecs, long long delay) {
	const size_t w = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	const unsigned lane = threadIdx.x & 31;
	const size_t base = w * region16;
	for (size_t i = 0; i < region16; i += 32 * (size_t)vecs) {
		uint4 v[8];
#pragma unroll
		for (int k = 0; k < 8; k++)
			if (k < vecs && i + 32 * k + lane < region16) v[k] = host[base + i + 32 * k + lane];
#pragma unroll
		for (int k = 0; k < 8; k++)
			if (k < vecs && i + 32 * k + lane < region16) dev[base + i + 32 * k + lane] = v[k];
		if (delay) { const long long t0 = clock64(); while (clock64() - t0 < delay) {} }
	}
}
// grid-stride streaming copy
__global__ void stream_copy(const uint4 *__restrict__ src, uint4 *dst, size_t n16) {
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

int main() {
	const size_t N = 1ull << 30;
	uint8_t *h, *d, *d2;
	CK(cudaHostAlloc(&h, N, cudaHostAllocMapped));
	CK(cudaMalloc(&d, N));
	CK(cudaMalloc(&d2, N));
	CK(cudaMemset(d, 0x5A, N));
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	float ms;
	auto report = [&](const char *name) { cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); printf("%-64s %8.3f ms  %7.2f GB/s\n", name, ms, N / ms / 1e6); };
	for (int rep = 0; rep < 2; rep++) {
		cudaEventRecord(e0); CK(cudaMemcpyAsync(h, d, N, cudaMemcpyDeviceToHost)); report("cudaMemcpyAsync D2H");
		cudaEventRecord(e0); stream_copy<<<148 * 8, 256>>>((const uint4 *)d, (uint4 *)h, N / 16); report("stream copy kernel, 16 B per lane, grid stride");
		cudaEventRecord(e0); write_pieces<<<1024, 128>>>(d, h, N / 4096, 0); report("4096 warps, ~900 B pieces at byte alignment (store_tile shape)");
		cudaEventRecord(e0); write_pieces<<<1024, 128>>>(d, h, N / 4096, 1); report("4096 warps, ~900 B pieces, whole 128 B lines only");
		for (int warps : {4096}) {
			char nm[128];
			cudaEventRecord(e0); write_regions<<<warps / 4, 128>>>((const uint4 *)d, (uint4 *)h, N / 16 / warps, 0, nullptr);
			snprintf(nm, sizeof nm, "%d warps, own region each, 16 B per lane", warps); report(nm);
			cudaEventRecord(e0); write_regions<<<warps / 4, 128>>>((const uint4 *)d, (uint4 *)h, N / 16 / warps, 1, (uint4 *)d2);
			snprintf(nm, sizeof nm, "%d warps, own region each, 16 B per lane + device copy", warps); report(nm);
			cudaEventRecord(e0); write_regions_u32<<<warps / 4, 128>>>((const uint32_t *)d, (uint32_t *)h, N / 4 / warps);
			snprintf(nm, sizeof nm, "%d warps, own region each, 4 B per lane", warps); report(nm);
		}
	}
	// SM reads of pinned host memory (the decoder pulling its own input): 4096 warps, 104 KB each = 416 MiB
	{
		const size_t R = 104 * 1024, warps = 4096, NB = R * warps;
		cudaStream_t s2;
		CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
		CK(cudaMemcpy(h, d, NB, cudaMemcpyDeviceToHost));
		for (int vecs : {1, 2, 4, 8}) {
			for (long long delay : {0ll, 20000ll, 200000ll}) {
				for (int duplex = 0; duplex < 2; duplex++) {
					if (duplex && delay == 20000) continue;
					char nm[160];
					CK(cudaDeviceSynchronize());
					cudaEventRecord(e0);
					if (duplex) CK(cudaMemcpyAsync(h + (N / 2), d2, N / 2, cudaMemcpyDeviceToHost, s2));
					read_regions<<<warps / 4, 128>>>((const uint4 *)h, (uint4 *)d, R / 16, vecs, delay);
					cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
					CK(cudaDeviceSynchronize());
					snprintf(nm, sizeof nm, "SM reads of host memory: %d B per warp and fetch, delay %lld ticks%s", vecs * 512, delay,
					         duplex ? ", 512 MiB D2H copy at the same time" : "");
					printf("%-96s %8.3f ms  %7.2f GB/s\n", nm, ms, NB / ms / 1e6);
				}
			}
		}
	}
	CK(cudaDeviceSynchronize());
	printf("check %d\n", h[12345] == 0x5A && h[N - 1] == 0x5A);
	return 0;
}
