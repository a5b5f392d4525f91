// api.cu -- C ABI of libb2deflate.so (include/b2deflate.h): the host runtime around the sm_100a kernels.
//
// This is the layer the reference's stream classes would call through Panama FFM:
//   InflaterInputStream.read  -> decomp/Open.read (Open.java:83-110)                 => b2d_inflate_batch
//   DeflaterOutputStream.writeBuffer -> Strategy.decide / Decision.compressTo
//                                    (DeflaterOutputStream.java:119-137)              => b2d_deflate_chunks
//   java.util.zip.CRC32 at GzipOutputStream.java:57 / GzipInputStream.java:72          => b2d_crc32*
// One context per GPU (b2d_init: one device; b2d_init_devices: several GPUs of the box behind the same entry points).
// A context owns its pipeline streams, a grow-only device scratch pool and the staging logic (H2D -> kernels -> D2H,
// sliced so copies overlap compute).  With several devices the host entry points partition the independent units
// (members, chunks) into contiguous ranges, one per device, driven by one host thread each; the compress direction
// joins the ranges at the scanned offsets.  There is no CPU codec in this library: without a usable
// sm_100 device every entry point returns B2D_ERR_NO_DEVICE.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <chrono>
#include <mutex>
#include <thread>
#include <vector>
#include <algorithm>
#include "kernels.h"

using namespace b2d;

namespace b2d {
static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace b2d

namespace {

// Device buffer of the scratch pool.  `last_use` orders reuse across streams: the *_dev entry points run on the
// caller's stream without synchronising, so the next user of the same buffer (possibly on another stream) first waits
// for the event the previous user recorded.
struct DevBuf {
	void *p = nullptr;
	size_t cap = 0;
	cudaEvent_t last_use = nullptr;
	bool used = false;
};

struct Ctx {
	bool ready = false;
	int device = -1;
	int sm_count = 0;
	std::mutex mu;                                                // serialises calls on this device
	cudaStream_t st[40] = {};                                    // pipeline streams (one per slice; st[32..] fixed roles)
	cudaEvent_t ev[48] = {};
	DevBuf in, out, meta, scratch, crc, scratch2, bits, scratch3;
	void *pinned_meta = nullptr;
	size_t pinned_meta_cap = 0;
	uint32_t *pinned_prog = nullptr;                              // per-member progress words the decoder writes (mapped)
	size_t pinned_prog_cap = 0;
	DevBuf *all_bufs[8] = {&in, &out, &meta, &scratch, &crc, &scratch2, &bits, &scratch3};
};

Ctx g_ctx[MAX_DEVICES];                  // g_ctx[0] is "GPU 0": the only one after b2d_init, the gather target otherwise
int g_ndev = 0;
std::mutex g_mu;                         // init / shutdown / multi-device calls
std::mutex g_err_mu;
char g_last_error[256] = "";

void set_error(const char *fmt, const char *a, const char *b) {
	std::lock_guard<std::mutex> lk(g_err_mu);
	snprintf(g_last_error, sizeof g_last_error, fmt, a, b);
}
int fail_cuda(cudaError_t e, const char *where) {
	set_error("%s: %s", where, cudaGetErrorString(e));
	cudaGetLastError();
	return e == cudaErrorMemoryAllocation ? B2D_ERR_OUT_OF_MEMORY : B2D_ERR_CUDA;
}
#define CK(call)                                                     \
	do {                                                             \
		cudaError_t e_ = (call);                                     \
		if (e_ != cudaSuccess) return fail_cuda(e_, #call);          \
	} while (0)

// Work that was enqueued before an error return may still read the caller's input or write the caller's output
// (H2D copies, mapped-memory stores, D2H copies): an armed guard waits for the device before the function returns.
struct SyncOnError {
	bool armed = true;
	~SyncOnError() { if (armed) { cudaDeviceSynchronize(); cudaGetLastError(); } }
};

int ensure(DevBuf &b, size_t bytes) {
	bytes = (bytes + 255) & ~(size_t)255;
	if (bytes <= b.cap) return 0;
	if (b.p) {
		if (b.used) cudaEventSynchronize(b.last_use);               // (cudaFree waits for the device anyway)
		cudaFree(b.p); b.p = nullptr; b.cap = 0;
	}
	size_t want = bytes + bytes / 8;
	cudaError_t e = cudaMalloc(&b.p, want);
	if (e != cudaSuccess) {
		cudaGetLastError();
		want = bytes;
		e = cudaMalloc(&b.p, want);
	}
	if (e != cudaSuccess) { b.p = nullptr; return fail_cuda(e, "cudaMalloc"); }
	b.cap = want;
	b.used = false;
	return 0;
}
// stream `st` is about to use / has just used the shared buffer
int acquire(DevBuf &b, cudaStream_t st) {
	if (b.used) CK(cudaStreamWaitEvent(st, b.last_use, 0));
	return 0;
}
int release(DevBuf &b, cudaStream_t st) {
	CK(cudaEventRecord(b.last_use, st));
	b.used = true;
	return 0;
}

int ensure_pinned_meta(Ctx &g, size_t bytes) {
	if (bytes <= g.pinned_meta_cap) return 0;
	if (g.pinned_meta) cudaFreeHost(g.pinned_meta);
	g.pinned_meta = nullptr;
	g.pinned_meta_cap = 0;
	CK(cudaMallocHost(&g.pinned_meta, bytes * 2));
	g.pinned_meta_cap = bytes * 2;
	return 0;
}

int ensure_pinned_prog(Ctx &g, size_t words) {
	if (words <= g.pinned_prog_cap) return 0;
	if (g.pinned_prog) cudaFreeHost(g.pinned_prog);
	g.pinned_prog = nullptr;
	g.pinned_prog_cap = 0;
	CK(cudaHostAlloc((void **)&g.pinned_prog, words * 2 * sizeof(uint32_t), cudaHostAllocMapped | cudaHostAllocPortable));
	g.pinned_prog_cap = words * 2;
	return 0;
}

void release_ctx(Ctx &g) {
	if (g.device >= 0) cudaSetDevice(g.device);
	for (DevBuf *b : g.all_bufs) {
		if (b->p) cudaFree(b->p);
		if (b->last_use) cudaEventDestroy(b->last_use);
		b->p = nullptr;
		b->cap = 0;
		b->last_use = nullptr;
		b->used = false;
	}
	if (g.pinned_meta) cudaFreeHost(g.pinned_meta);
	g.pinned_meta = nullptr;
	g.pinned_meta_cap = 0;
	if (g.pinned_prog) cudaFreeHost(g.pinned_prog);
	g.pinned_prog = nullptr;
	g.pinned_prog_cap = 0;
	for (auto &s : g.st) { if (s) cudaStreamDestroy(s); s = nullptr; }
	for (auto &e : g.ev) { if (e) cudaEventDestroy(e); e = nullptr; }
	g.ready = false;
	g.device = -1;
}

int init_ctx(Ctx &g, int device) {
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0 || device < 0 || device >= count) {
		char num[16];
		snprintf(num, sizeof num, "%d", device);
		set_error("no usable CUDA device %s (%s)", num, e == cudaSuccess ? "out of range" : cudaGetErrorString(e));
		cudaGetLastError();
		return B2D_ERR_NO_DEVICE;
	}
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, device));
	if (prop.major != 10) {
		char what[64];
		snprintf(what, sizeof what, "device %d is sm_%d%d", device, prop.major, prop.minor);
		set_error("%s; libb2deflate is built for sm_100a only%s", what, "");
		return B2D_ERR_NO_DEVICE;
	}
	CK(cudaSetDevice(device));
	g.device = device;
	for (auto &s : g.st) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
	for (auto &ev : g.ev) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
	for (DevBuf *b : g.all_bufs) CK(cudaEventCreateWithFlags(&b->last_use, cudaEventDisableTiming));
	// the decoder's shared-memory addressing assumes where the static segment starts in the shared window (inflate.cu):
	// verified here, once per device, instead of trapping inside the kernel
	uint32_t base = 0;
	CK(probe_inflate_smem_base(&base, g.st[0]));
	if (base != INFLATE_SMEM_WINDOW_BASE) {
		char got[32];
		snprintf(got, sizeof got, "0x%x", base);
		set_error("static shared memory starts at window address %s, the decoder is built for 0x400%s", got, "");
		return B2D_ERR_CUDA;
	}
	g.sm_count = prop.multiProcessorCount;
	g.ready = true;
	return B2D_OK;
}

int init_devices_locked(const int *devices, int n) {
	// The host pipelines keep a dozen streams busy at once (eight slices of a batch each decode on their own stream, next
	// to the copy streams).  CUDA maps streams onto 8 hardware queues by default and streams that share a queue
	// serialise; the setting is read when the process creates its first CUDA context, so it only helps if nobody
	// initialised CUDA before us (a harness that did -- bench.py imports torch first -- sets the variable itself).
	// (16, not 32: context creation takes 1.2 s with 8 or 16 queues and 2.2 s with 32.)
	setenv("CUDA_DEVICE_MAX_CONNECTIONS", "16", 0);
	bool same = n == g_ndev;
	for (int i = 0; same && i < n; i++) same = g_ctx[i].ready && g_ctx[i].device == devices[i];
	if (same) return B2D_OK;
	for (int i = 0; i < g_ndev; i++) {
		std::lock_guard<std::mutex> lk(g_ctx[i].mu);
		if (g_ctx[i].ready || g_ctx[i].device >= 0) { cudaSetDevice(g_ctx[i].device); cudaDeviceSynchronize(); release_ctx(g_ctx[i]); }
	}
	g_ndev = 0;
	for (int i = 0; i < n; i++) {
		int r = init_ctx(g_ctx[i], devices[i]);
		if (r != B2D_OK) {
			for (int j = 0; j <= i; j++) release_ctx(g_ctx[j]);
			return r;
		}
	}
	g_ndev = n;
	if (n > 1) {   // peer access towards GPU 0 for the gather (best effort: without it cudaMemcpyPeerAsync stages through the host)
		for (int i = 1; i < n; i++) {
			int can = 0;
			if (cudaDeviceCanAccessPeer(&can, g_ctx[i].device, g_ctx[0].device) == cudaSuccess && can) {
				cudaSetDevice(g_ctx[i].device);
				cudaDeviceEnablePeerAccess(g_ctx[0].device, 0);
				cudaSetDevice(g_ctx[0].device);
				cudaDeviceEnablePeerAccess(g_ctx[i].device, 0);
			}
			cudaGetLastError();
		}
		cudaSetDevice(g_ctx[0].device);
	}
	return B2D_OK;
}

Ctx *primary() { return g_ndev > 0 && g_ctx[0].ready ? &g_ctx[0] : nullptr; }
// the context of the device a device pointer lives on (the *_dev entry points); GPU 0's if it cannot be told
Ctx *ctx_for_pointer(const void *p) {
	Ctx *g = primary();
	if (!g || g_ndev == 1 || !p) return g;
	cudaPointerAttributes pa;
	if (cudaPointerGetAttributes(&pa, p) == cudaSuccess && pa.type == cudaMemoryTypeDevice)
		for (int i = 0; i < g_ndev; i++)
			if (g_ctx[i].device == pa.device) return &g_ctx[i];
	cudaGetLastError();
	return g;
}

int normalise_opts(const b2d_deflate_opts *o, uint64_t in_len, DeflateParams &p) {
	b2d_deflate_opts d;
	memset(&d, 0, sizeof d);
	d.lazy = -1;
	if (o) d = *o;
	p.framing = d.framing;
	if (p.framing != B2D_FRAMING_CHUNKED && p.framing != B2D_FRAMING_REFERENCE) return B2D_ERR_BAD_ARGUMENT;
	p.block_bytes = d.block_bytes ? d.block_bytes : (1u << 16);
	if (p.block_bytes < 4096 || p.block_bytes > (1u << 20)) return B2D_ERR_BAD_ARGUMENT;
	if (p.framing == B2D_FRAMING_REFERENCE) {
		// one unit: chunk = the whole input rounded up to whole blocks (must fit 32 bits)
		uint64_t nb = std::max<uint64_t>(1, (in_len + p.block_bytes - 1) / p.block_bytes);
		uint64_t cb = nb * p.block_bytes;
		if (cb > 0xFFFF0000ull) return B2D_ERR_BAD_ARGUMENT;
		p.chunk_bytes = (uint32_t)cb;
	} else {
		p.chunk_bytes = d.chunk_bytes ? d.chunk_bytes : (1u << 20);
		if (p.chunk_bytes % p.block_bytes != 0) return B2D_ERR_BAD_ARGUMENT;
	}
	p.mode = d.mode;
	if (p.mode < B2D_MODE_AUTO || p.mode > B2D_MODE_DYNAMIC) return B2D_ERR_BAD_ARGUMENT;
	p.search = d.search;
	if (p.search < B2D_SEARCH_DEFAULT || p.search > B2D_SEARCH_FULL) return B2D_ERR_BAD_ARGUMENT;
	p.depth = d.chain_depth > 0 ? d.chain_depth : 4;
	p.lazy = d.lazy < 0 ? 1 : (d.lazy ? 1 : 0);
	if (p.search != B2D_SEARCH_DEFAULT) p.lazy = d.lazy > 0 ? 1 : 0;     // the reference strategies are greedy
	p.is_last = d.is_last ? 1 : 0;
	p.checksum = d.checksum;
	if (p.checksum != B2D_CHECKSUM_CRC32 && p.checksum != B2D_CHECKSUM_ADLER32) return B2D_ERR_BAD_ARGUMENT;
	p.leaf_bytes = 0;
	if (d.split_min_bytes && d.split_min_bytes < p.block_bytes) {
		const uint32_t l = d.split_min_bytes;
		if (l < 4096 || (l & (l - 1)) || p.block_bytes % l != 0 || p.block_bytes / l > 16 ||
		    ((p.block_bytes / l) & (p.block_bytes / l - 1)))
			return B2D_ERR_BAD_ARGUMENT;
		p.leaf_bytes = l;
	}
	return 0;
}

uint32_t combine_checksum(const DeflateParams &p, uint32_t acc, uint32_t piece, uint64_t len) {
	return (p.checksum == B2D_CHECKSUM_ADLER32 ? host_adler32_combine : host_crc32_combine)(acc, piece, len);
}

// ---- device-pointer cores (caller holds g.mu and has made g.device current) ----

int inflate_dev_locked(const uint8_t *d_in, const uint64_t *d_in_off, uint32_t n, uint8_t *d_out,
                       const uint64_t *d_out_off, uint64_t *d_out_len, uint64_t *d_in_consumed, uint32_t *d_crc,
                       int32_t *d_status, uint32_t flags, cudaStream_t st, uint8_t *out_mirror = nullptr,
                       const uint64_t *d_in_end = nullptr, uint32_t *progress = nullptr, const uint8_t *in_host = nullptr,
                       const uint32_t *landed = nullptr, uint32_t group_size = 0) {
	if (n == 0) return B2D_OK;
	CK(launch_inflate(d_in, d_in_off, n, d_out, d_out_off, d_out_len, d_in_consumed, d_status, flags, st, out_mirror, d_in_end, progress,
	                  in_host, landed, group_size));
	if ((flags & B2D_INFLATE_ADLER32) && d_crc) CK(launch_adler32_segments(d_out, d_out_off, d_out_len, n, d_crc, st));
	else if ((flags & B2D_INFLATE_CRC32) && d_crc) CK(launch_crc32_segments(d_out, d_out_off, d_out_len, n, d_crc, st));
	return B2D_OK;
}

int deflate_dev_locked(Ctx &g, const uint8_t *d_in, uint64_t in_len, const DeflateParams &p, uint8_t *d_out, uint64_t out_cap,
                       uint64_t *d_total, uint64_t *d_chunk_len, uint32_t *d_chunk_crc, cudaStream_t st,
                       uint32_t *d_block_bits = nullptr, DevBuf *scratch = nullptr, bool overlap_stages = true) {
	if (!scratch) scratch = &g.scratch;
	DeflateAux aux;                                               // (one call at a time per context: g.mu is held)
	aux.stream = g.st[34];
	for (int i = 0; i < 6; i++) aux.ev[i] = g.ev[24 + i];
	size_t sb = deflate_scratch_bytes(in_len, p);
	int r = ensure(*scratch, sb);
	if (r) return r;
	if ((r = acquire(*scratch, st))) return r;
	CK(launch_deflate(d_in, in_len, p, d_out, out_cap, d_total, d_chunk_len, scratch->p, scratch->cap, st, d_block_bits,
	                  overlap_stages ? &aux : nullptr));
	if ((r = release(*scratch, st))) return r;
	if (d_chunk_crc) {
		uint32_t n_chunks = (uint32_t)((in_len + p.chunk_bytes - 1) / p.chunk_bytes);
		if (p.checksum == B2D_CHECKSUM_ADLER32) CK(launch_adler32_pieces(d_in, in_len, p.chunk_bytes, n_chunks, d_chunk_crc, st));
		else CK(launch_crc32_pieces(d_in, in_len, p.chunk_bytes, n_chunks, d_chunk_crc, st));
	}
	return B2D_OK;
}

// Contiguous ranges of `n` units for `parts` workers, balanced by a cumulative weight (weight_upto(i) = weight of
// units [0, i)); -> bounds[parts + 1]
template <typename W>
std::vector<uint32_t> balanced_ranges(uint32_t n, int parts, W weight_upto) {
	std::vector<uint32_t> b(parts + 1, n);
	b[0] = 0;
	const double total = (double)weight_upto(n);
	uint32_t i = 0;
	for (int k = 1; k < parts; k++) {
		const double want = total * k / parts;
		while (i < n && (double)weight_upto(i) < want) i++;
		b[k] = i;
	}
	return b;
}

// ---------------------------------------------------------------- inflate (host pointers, one device)
// Member i occupies in[begin[i], end[i]) (ranges in ascending order) and decodes to out[out_off[i], out_off[i+1]).
// One member decodes at a fixed, serial pace whatever the batch size, so the batch is cut into at most four large
// slices (>= 1024 members each, enough warps to fill the GPU together), each on its own stream: the slice kernels run
// side by side, slice k's H2D overlaps the decode of slices < k and its D2H overlaps the decode of slices > k.  Pinned
// buffers (b2d_alloc_pinned) make the copies asynchronous; PCIe is the end-to-end bound.  Members advance at the same
// pace, so a copy after the kernel could not overlap anything; a pinned output buffer is therefore filled WHILE the
// decode runs, in one of two ways:
//   progress ("ce"): slots of one size (the usual batch).  Every decoding warp reports to a mapped host word how many
//            32 KiB pieces of its member are final in device memory; this thread polls the words and, as soon as piece
//            j of every member of a slice is final, moves that piece of all of them with ONE strided copy-engine
//            transfer (cudaMemcpy2DAsync: rows of 32 KiB, pitch = slot size).  The copy engine reaches the full PCIe
//            rate (57 GB/s measured against 46-49 for SM stores) and the kernel spends no instructions on delivery.
//   mirror : ragged slots.  The kernel writes the finished output to the mapped host address itself, as whole
//            128-byte lines, next to the device copy it keeps for back-references and the checksum.
// B2D_INFLATE_D2H=ce|mirror|copy picks one for diagnosis (copy = plain D2H after the kernels, what pageable buffers get).
int inflate_host(Ctx &g, const uint8_t *in, const uint64_t *begin, const uint64_t *end, uint32_t n, uint8_t *out,
                 const uint64_t *out_off, uint64_t *out_len, uint64_t *in_consumed, uint32_t *crc32, int32_t *status,
                 uint32_t flags) {
	std::lock_guard<std::mutex> lk(g.mu);
	const auto t_entry = std::chrono::steady_clock::now();
	if (!g.ready) return B2D_ERR_NO_DEVICE;
	if (n == 0) return B2D_OK;
	for (uint32_t i = 0; i < n; i++)
		if (end[i] < begin[i] || out_off[i + 1] < out_off[i] || (i + 1 < n && begin[i + 1] < begin[i]) || (i + 1 < n && end[i + 1] < end[i]))
			return B2D_ERR_BAD_ARGUMENT;
	const uint64_t in0 = begin[0], in_total = end[n - 1] - in0;
	const uint64_t out0 = out_off[0], out_total = out_off[n] - out0;
	if ((in_total && !in) || (out_total && !out)) return B2D_ERR_BAD_ARGUMENT;
	CK(cudaSetDevice(g.device));
	int r;
	if ((r = ensure(g.in, in_total + 64 + 256))) return r;
	if ((r = ensure(g.out, out_total + 256))) return r;
	// meta layout (device): in_begin[n] in_end[n] out_off[n+1] | out_len[n] consumed[n] crc[n] status[n]
	const size_t m_begin = 0, m_end = (size_t)n * 8, m_off_out = m_end + (size_t)n * 8, m_len = m_off_out + (size_t)(n + 1) * 8,
	             m_cons = m_len + (size_t)n * 8, m_crc = m_cons + (size_t)n * 8, m_stat = m_crc + (size_t)n * 4,
	             m_total = m_stat + (size_t)n * 4, m_flags = (m_total + 15) & ~(size_t)15;      // + "group landed" words (device only)
	if ((r = ensure(g.meta, m_flags + 4 * 512))) return r;
	if ((r = ensure_pinned_meta(g, m_total))) return r;
	uint8_t *hm = (uint8_t *)g.pinned_meta, *dm = (uint8_t *)g.meta.p;
	uint64_t *h_begin = (uint64_t *)(hm + m_begin), *h_end = (uint64_t *)(hm + m_end), *h_out_off = (uint64_t *)(hm + m_off_out);
	for (uint32_t i = 0; i < n; i++) { h_begin[i] = begin[i] - in0; h_end[i] = end[i] - in0; }
	for (uint32_t i = 0; i <= n; i++) h_out_off[i] = out_off[i] - out0;
	uint8_t *d_in = (uint8_t *)g.in.p, *d_out = (uint8_t *)g.out.p;
	// A pinned INPUT buffer is not copied: the decoding warps pull their members' bytes from the mapped host address as
	// they go (inflate.cu, stage_input), so every member starts at once instead of queueing behind the host-to-device
	// copy of the bytes in front of it.  B2D_INFLATE_H2D=copy keeps the copy for diagnosis.
	const uint8_t *in_host = nullptr;
	bool hybrid = false;                                         // B2D_INFLATE_H2D=hybrid: pull AND copy (see below; slower, kept for diagnosis)
	if (in_total) {
		const char *hm_ = getenv("B2D_INFLATE_H2D");
		hybrid = hm_ && !strcmp(hm_, "hybrid");
		cudaPointerAttributes pa0, pa1;
		if (!(hm_ && !strcmp(hm_, "copy")) &&
		    cudaPointerGetAttributes(&pa0, in + in0) == cudaSuccess && pa0.type == cudaMemoryTypeHost && pa0.devicePointer &&
		    cudaPointerGetAttributes(&pa1, in + in0 + in_total - 1) == cudaSuccess && pa1.type == cudaMemoryTypeHost &&
		    (const uint8_t *)pa1.devicePointer - (const uint8_t *)pa0.devicePointer == (ptrdiff_t)(in_total - 1)) {
			in_host = (const uint8_t *)pa0.devicePointer;
			d_in += ((uintptr_t)in_host - (uintptr_t)d_in) & 127;       // same 128-byte phase: whole lines map to whole lines
		}
		cudaGetLastError();
	}
	uint8_t *mirror = nullptr;
	uint32_t *progress = nullptr;
	uint64_t slot = 0;                                           // the slots' common size (progress mode)
	if (out_total) {
		const char *nm_ = getenv("B2D_NO_MIRROR");
		const char *dm_ = getenv("B2D_INFLATE_D2H");
		int want = (nm_ && nm_[0] == '1') ? 0 : 3;                  // 0 copy, 1 mirror, 2 ce, 3 choose
		if (dm_) want = !strcmp(dm_, "copy") ? 0 : !strcmp(dm_, "mirror") ? 1 : !strcmp(dm_, "ce") ? 2 : want;
		cudaPointerAttributes pa0, pa1;
		if (want != 0 &&
		    cudaPointerGetAttributes(&pa0, out + out0) == cudaSuccess && pa0.type == cudaMemoryTypeHost && pa0.devicePointer &&
		    cudaPointerGetAttributes(&pa1, out + out0 + out_total - 1) == cudaSuccess && pa1.type == cudaMemoryTypeHost &&
		    (uint8_t *)pa1.devicePointer - (uint8_t *)pa0.devicePointer == (ptrdiff_t)(out_total - 1)) {
			bool uniform = n >= 64 && out_off[1] > out_off[0];
			for (uint32_t i = 1; uniform && i < n; i++) uniform = out_off[i + 1] - out_off[i] == out_off[1] - out_off[0];
			if (want != 1 && uniform) {
				if ((r = ensure_pinned_prog(g, n))) return r;
				progress = g.pinned_prog;
				memset(progress, 0, (size_t)n * sizeof(uint32_t));
				slot = out_off[1] - out_off[0];
			} else {
				mirror = (uint8_t *)pa0.devicePointer;
				d_out += ((uintptr_t)mirror - (uintptr_t)d_out) & 127;      // same 128-byte phase as the host buffer (256 B slack)
			}
		}
		cudaGetLastError();
	}
	for (DevBuf *b : {&g.in, &g.out, &g.meta}) if ((r = acquire(*b, g.st[0]))) return r;
	SyncOnError guard;
	CK(cudaMemcpyAsync(dm, hm, m_len, cudaMemcpyHostToDevice, g.st[0]));
	if (hybrid && in_host) CK(cudaMemsetAsync(dm + m_flags, 0, 4 * 512, g.st[0]));
	CK(cudaEventRecord(g.ev[0], g.st[0]));
	uint32_t max_slices = 8, min_per = 512;                       // (4096 members: 8 x 512 measured 24.4 ms, 4 x 1024 24.8, 1 x 4096 26.2)
	if (const char *sl_ = getenv("B2D_INFLATE_SLICES")) {      // diagnostic: "slices[,members per slice at least]"
		unsigned a_ = 0, b_ = 0;
		int got = sscanf(sl_, "%u,%u", &a_, &b_);
		if (got >= 1 && a_ >= 1 && a_ <= 32) max_slices = a_;
		if (got >= 2 && b_ >= 1) min_per = b_;
	}
	uint32_t n_slices = std::min<uint32_t>(max_slices, std::max<uint32_t>(1, n / min_per));
	uint32_t per = (n + n_slices - 1) / n_slices;
	// Hybrid (diagnostic): the blob is copied by the copy engine as well, in groups of members on a stream of its own,
	// and every group raises a flag when it has landed (the warps stop pulling their own input then).  Measured at
	// 4096 x 256 KiB: the kernels end 1.5 ms earlier, but the extra copy-engine traffic slows the output copies and the
	// call takes 27.0 ms against 24.4 with the warps pulling everything -- so it is not the default.
	hybrid = hybrid && in_host;
	const uint32_t group = hybrid ? std::max<uint32_t>(1, (per + 7) / 8) : 0;            // <= 8 groups per slice
	if (hybrid) {
		cudaStream_t sH = g.st[33];
		CK(cudaStreamWaitEvent(sH, g.ev[0], 0));
		for (uint32_t a = 0, sl = 0; a < n; a += per, sl++) {
			const uint32_t b = std::min(n, a + per);
			uint32_t fl = 8 * sl;                                // slice sl's flags: words 8 sl .. 8 sl + 7
			for (uint32_t ga = a; ga < b; ga += group, fl++) {
				const uint32_t gb = std::min(b, ga + group);
				const uint64_t ia = h_begin[ga], ib = h_end[gb - 1];
				if (ib > ia) CK(cudaMemcpyAsync(d_in + ia, in + in0 + ia, ib - ia, cudaMemcpyHostToDevice, sH));
				CK(cudaMemsetAsync(dm + m_flags + 4 * (size_t)fl, 1, 4, sH));
			}
		}
		CK(cudaEventRecord(g.ev[41], sH));
	}
	int k = 0;
	const char *tr_ = getenv("B2D_TRACE");                  // diagnostic: the slices' H2D / kernel / D2H timeline on stderr
	const bool trace = tr_ != nullptr && tr_[0] == '1';
	cudaEvent_t te[32][4];
	for (uint32_t a = 0; a < n; a += per, k++) {
		uint32_t b = std::min(n, a + per);
		cudaStream_t st = g.st[k];
		if (k > 0) CK(cudaStreamWaitEvent(st, g.ev[0], 0));
		uint64_t ia = h_begin[a], ib = h_end[b - 1], oa = h_out_off[a], ob = h_out_off[b];
		// (a kernel may read the aligned words around its slice while a neighbour's copy lands in them; those bytes
		// are shifted out / masked by the bit reader, so the race is benign)
		if (trace) { for (int q = 0; q < 4; q++) cudaEventCreate(&te[k][q]); cudaEventRecord(te[k][0], st); }
		if (ib > ia && !in_host) CK(cudaMemcpyAsync(d_in + ia, in + in0 + ia, ib - ia, cudaMemcpyHostToDevice, st));
		if (trace) cudaEventRecord(te[k][1], st);
		r = inflate_dev_locked(d_in, (const uint64_t *)(dm + m_begin) + a, b - a, d_out,
		                       (const uint64_t *)(dm + m_off_out) + a, (uint64_t *)(dm + m_len) + a,
		                       (uint64_t *)(dm + m_cons) + a, (uint32_t *)(dm + m_crc) + a,
		                       (int32_t *)(dm + m_stat) + a, flags, st, mirror, (const uint64_t *)(dm + m_end) + a,
		                       progress ? progress + a : nullptr, in_host,
		                       hybrid ? (const uint32_t *)(dm + m_flags) + 8 * (size_t)k : nullptr, group);
		if (r) return r;
		if (trace) cudaEventRecord(te[k][2], st);
		if (ob > oa && !mirror && !progress) CK(cudaMemcpyAsync(out + out0 + oa, d_out + oa, ob - oa, cudaMemcpyDeviceToHost, st));
		if (trace) cudaEventRecord(te[k][3], st);
	}
	if (progress) {
		// Deliver while the kernels run: piece j of a slice moves as soon as every member of the slice has reported it.
		const uint64_t P = (uint64_t)1 << INFLATE_PROGRESS_SHIFT;
		const uint32_t n_pieces = (uint32_t)((slot + P - 1) / P);
		cudaStream_t sD = g.st[32];
		std::vector<uint32_t> next(k, 0);
		volatile uint32_t *pv = progress;
		int open = k;
		uint64_t idle = 0;
		while (open > 0) {
			bool moved = false;
			open = 0;
			for (int s = 0; s < k; s++) {
				if (next[s] >= n_pieces) continue;
				const uint32_t a = (uint32_t)s * per, b = std::min(n, a + per);
				uint32_t m = 0xFFFFFFFFu;
				for (uint32_t i = a; i < b && m > next[s]; i++) m = std::min(m, (uint32_t)pv[i]);
				if (m > next[s]) {
					const uint32_t upto = std::min(m, n_pieces);
					const uint64_t x0 = (uint64_t)next[s] * P, x1 = std::min<uint64_t>(slot, (uint64_t)upto * P);
					CK(cudaMemcpy2DAsync(out + out0 + h_out_off[a] + x0, slot, d_out + h_out_off[a] + x0, slot, x1 - x0, b - a,
					                     cudaMemcpyDeviceToHost, sD));
					next[s] = upto;
					moved = true;
				}
				if (next[s] < n_pieces) open++;
			}
			if (moved) { idle = 0; continue; }
			if ((++idle & 0x3FFF) == 0) {                       // nothing new for a while: are the kernels still alive?
				bool all_done = true;
				for (int s = 0; s < k; s++) {
					cudaError_t q = cudaStreamQuery(g.st[s]);
					if (q == cudaErrorNotReady) all_done = false;
					else if (q != cudaSuccess) return fail_cuda(q, "inflate kernels");
				}
				if (all_done) {                                  // (every member reports 0x7FFFFFFF before its kernel ends)
					bool late = false;
					for (uint32_t i = 0; i < n; i++) late |= pv[i] != 0x7FFFFFFFu;
					if (late) { set_error("%s%s", "inflate: progress words incomplete after the kernels ended", ""); return B2D_ERR_CUDA; }
				}
			}
#if defined(__x86_64__)
			__builtin_ia32_pause();
#endif
		}
		CK(cudaEventRecord(g.ev[40], sD));
		CK(cudaStreamWaitEvent(g.st[0], g.ev[40], 0));
	}
	for (int s = 1; s < k; s++) {
		CK(cudaEventRecord(g.ev[s], g.st[s]));
		CK(cudaStreamWaitEvent(g.st[0], g.ev[s], 0));
	}
	if (hybrid) CK(cudaStreamWaitEvent(g.st[0], g.ev[41], 0));
	CK(cudaMemcpyAsync(hm + m_len, dm + m_len, m_total - m_len, cudaMemcpyDeviceToHost, g.st[0]));
	for (DevBuf *b : {&g.in, &g.out, &g.meta}) if ((r = release(*b, g.st[0]))) return r;
	const auto t_issued = std::chrono::steady_clock::now();
	CK(cudaStreamSynchronize(g.st[0]));
	guard.armed = false;
	if (trace) {
		const auto t_done = std::chrono::steady_clock::now();
		fprintf(stderr, "[b2d trace] dev %d host: entry -> all work issued %.3f ms, -> synchronized %.3f ms\n", g.device,
		        std::chrono::duration<double, std::milli>(t_issued - t_entry).count(),
		        std::chrono::duration<double, std::milli>(t_done - t_entry).count());
		for (int q = 0; q < k; q++) {
			float t[4] = {-1, -1, -1, -1};
			for (int j = 0; j < 4; j++) {
				cudaError_t ee = cudaEventElapsedTime(&t[j], te[0][0], te[q][j]);
				if (ee != cudaSuccess) { fprintf(stderr, "[b2d trace] event (%d,%d): %s\n", q, j, cudaGetErrorString(ee)); cudaGetLastError(); }
			}
			fprintf(stderr, "[b2d trace] slice %d: h2d %.3f..%.3f ms, kernels ..%.3f ms, d2h ..%.3f ms%s\n", q, t[0], t[1], t[2], t[3],
			        mirror ? " [output mirrored by the kernel]" : "");
		}
		for (int q = 0; q < k; q++)
			for (int j = 0; j < 4; j++) cudaEventDestroy(te[q][j]);
	}
	memcpy(out_len, hm + m_len, (size_t)n * 8);
	memcpy(in_consumed, hm + m_cons, (size_t)n * 8);
	if (crc32 && (flags & (B2D_INFLATE_CRC32 | B2D_INFLATE_ADLER32))) memcpy(crc32, hm + m_crc, (size_t)n * 4);
	memcpy(status, hm + m_stat, (size_t)n * 4);
	return B2D_OK;
}

// All initialised devices: contiguous member ranges balanced by output capacity, one host thread per device, no
// exchange step (the inflate direction needs no collective).
int inflate_host_all(const uint8_t *in, const uint64_t *begin, const uint64_t *end, uint32_t n, uint8_t *out,
                     const uint64_t *out_off, uint64_t *out_len, uint64_t *in_consumed, uint32_t *crc32, int32_t *status,
                     uint32_t flags) {
	if (g_ndev == 0) return B2D_ERR_NO_DEVICE;
	if (n == 0) return B2D_OK;
	if (!begin || !end || !out_off || !out_len || !in_consumed || !status) return B2D_ERR_BAD_ARGUMENT;
	if ((flags & (B2D_INFLATE_CRC32 | B2D_INFLATE_ADLER32)) && !crc32) return B2D_ERR_BAD_ARGUMENT;
	const int parts = (int)std::min<uint64_t>((uint64_t)g_ndev, std::max<uint32_t>(1, n / 64));
	if (parts <= 1) return inflate_host(g_ctx[0], in, begin, end, n, out, out_off, out_len, in_consumed, crc32, status, flags);
	for (uint32_t i = 0; i < n; i++) if (out_off[i + 1] < out_off[i]) return B2D_ERR_BAD_ARGUMENT;
	const std::vector<uint32_t> b = balanced_ranges(n, parts, [&](uint32_t i) { return out_off[i] - out_off[0] + i; });
	std::vector<int> rc(parts, B2D_OK);
	std::vector<std::thread> th;
	for (int d = 0; d < parts; d++) {
		th.emplace_back([&, d] {
			const uint32_t a = b[d], c = b[d + 1];
			if (c > a)
				rc[d] = inflate_host(g_ctx[d], in, begin + a, end + a, c - a, out, out_off + a, out_len + a, in_consumed + a,
				                     crc32 ? crc32 + a : nullptr, status + a, flags);
		});
	}
	for (auto &t : th) t.join();
	for (int d = 0; d < parts; d++) if (rc[d] != B2D_OK) return rc[d];
	return B2D_OK;
}

// ---------------------------------------------------------------- deflate (host pointers)
struct DeflateHostOut {           // where one device's share of a call ends up on the host
	uint64_t total = 0;           // compressed bytes of the range
	uint32_t crc = 0;             // running checksum after the range (only meaningful for a single-device call)
};

// One device, chunked framing, >= 256 chunks.  Pipelined: the chunks are cut into up to 4 slices of >= 256 chunks.
// Slice k+1's H2D (stream A) runs under slice k's kernels and slice k-1's D2H (stream C); the slices' kernels alternate
// between two streams with a scratch area each, so the latency-bound kernels of one slice (chains: one warp per 256 KiB
// segment, the same few milliseconds whatever the slice size) run under the issue-bound ones of its neighbour.  The
// slices are ordinary calls (is_last only on the final one), so the bytes are the same as one big call.
//   deliver = true : payload slices are copied to `out` as they finish (single device)
//   deliver = false: payloads stay in g.out (slice k at slice_bound * k); the caller places them once every device's
//                    total is known (deflate_pipelined_deliver)
struct SlicePlan {
	uint32_t per = 0, n_slices = 0;
	uint64_t slice_in = 0, slice_bound = 0;
	size_t ms_clen = 0, ms_ccrc = 0, ms_total = 0;
	std::vector<uint64_t> slice_total;
};

int deflate_pipelined(Ctx &g, const uint8_t *in, uint64_t in_len, const DeflateParams &p, uint8_t *out, uint64_t out_cap,
                      bool want_crc, uint32_t *chunk_crc_out, uint64_t *chunk_out_len, uint32_t *block_bits_out, bool deliver,
                      SlicePlan &sp, uint64_t *total_out) {
	int r;
	const uint32_t n_chunks = (uint32_t)((in_len + p.chunk_bytes - 1) / p.chunk_bytes);
	const uint32_t bpc = p.chunk_bytes / p.block_bytes;
	uint32_t max_slices = 4, min_per = 256;          // measured at 1 GiB: 2 slices 50.4 ms, 4: 48.1, 8: 56.2, 1: 62.5
	if (const char *sl_ = getenv("B2D_DEFLATE_SLICES")) {      // diagnostic: "slices[,chunks per slice at least]"
		unsigned a_ = 0, b_ = 0;
		int got = sscanf(sl_, "%u,%u", &a_, &b_);
		if (got >= 1 && a_ >= 1 && a_ <= 8) max_slices = a_;
		if (got >= 2 && b_ >= 1) min_per = b_;
	}
	if (n_chunks < min_per) min_per = std::max<uint32_t>(1, (n_chunks + 1) / 2);   // (a device's share of a multi-GPU call)
	sp.per = std::max<uint32_t>(min_per, (n_chunks + max_slices - 1) / max_slices);
	sp.n_slices = (n_chunks + sp.per - 1) / sp.per;
	sp.slice_in = (uint64_t)sp.per * p.chunk_bytes;
	sp.slice_bound = (deflate_bound_bytes(sp.slice_in, p.chunk_bytes, p.block_bytes) + 255) & ~(uint64_t)255;
	sp.ms_clen = 8;
	sp.ms_ccrc = sp.ms_clen + (size_t)(sp.per + 1) * 8;
	sp.ms_total = (sp.ms_ccrc + (size_t)(sp.per + 1) * 4 + 15) & ~(size_t)15;
	sp.slice_total.assign(sp.n_slices, 0);
	const uint32_t per = sp.per, n_slices = sp.n_slices;
	const uint64_t slice_in = sp.slice_in, slice_bound = sp.slice_bound;
	const size_t ms_clen = sp.ms_clen, ms_ccrc = sp.ms_ccrc, ms_total = sp.ms_total;
	if ((r = ensure(g.in, in_len + 64))) return r;
	if ((r = ensure(g.out, slice_bound * n_slices + 64))) return r;
	if ((r = ensure(g.meta, ms_total * n_slices))) return r;
	if ((r = ensure_pinned_meta(g, ms_total * n_slices))) return r;
	if (block_bits_out && (r = ensure(g.bits, (size_t)(n_chunks * bpc + 1) * 4))) return r;
	DeflateParams ps = p;
	if ((r = ensure(g.scratch, deflate_scratch_bytes(slice_in, ps)))) return r;
	if (n_slices > 1 && (r = ensure(g.scratch3, deflate_scratch_bytes(slice_in, ps)))) return r;
	uint8_t *dm = (uint8_t *)g.meta.p, *hm = (uint8_t *)g.pinned_meta;
	cudaStream_t sA = g.st[1], sC = g.st[2];
	cudaStream_t sK[2] = {g.st[0], g.st[3]};
	DevBuf *scr[2] = {&g.scratch, &g.scratch3};
	for (DevBuf *b : {&g.in, &g.out, &g.meta, &g.bits}) if ((r = acquire(*b, sA))) return r;
	SyncOnError guard;
	for (uint32_t k = 0; k < n_slices; k++) {
		const uint64_t a = (uint64_t)k * slice_in, len = std::min<uint64_t>(slice_in, in_len - a);
		CK(cudaMemcpyAsync((uint8_t *)g.in.p + a, in + a, len, cudaMemcpyHostToDevice, sA));
		CK(cudaEventRecord(g.ev[k], sA));
	}
	for (uint32_t k = 0; k < n_slices; k++) {
		const uint64_t a = (uint64_t)k * slice_in, len = std::min<uint64_t>(slice_in, in_len - a);
		ps.is_last = (p.is_last && k + 1 == n_slices) ? 1 : 0;
		uint8_t *dmk = dm + ms_total * k;
		cudaStream_t sB = sK[k & 1];
		CK(cudaStreamWaitEvent(sB, g.ev[k], 0));
		r = deflate_dev_locked(g, (const uint8_t *)g.in.p + a, len, ps, (uint8_t *)g.out.p + slice_bound * k, slice_bound,
		                       (uint64_t *)dmk, (uint64_t *)(dmk + ms_clen), want_crc ? (uint32_t *)(dmk + ms_ccrc) : nullptr, sB,
		                       block_bits_out ? (uint32_t *)g.bits.p + (size_t)k * per * bpc : nullptr, scr[k & 1],
		                       false /* the slices overlap each other already */);
		if (r) return r;
		CK(cudaMemcpyAsync(hm + ms_total * k, dmk, ms_total, cudaMemcpyDeviceToHost, sB));
		CK(cudaEventRecord(g.ev[8 + k], sB));
	}
	uint64_t total = 0;
	for (uint32_t k = 0; k < n_slices; k++) {
		CK(cudaEventSynchronize(g.ev[8 + k]));
		const uint8_t *hmk = hm + ms_total * k;
		const uint64_t tk = *(const uint64_t *)hmk;
		sp.slice_total[k] = tk;
		if (deliver) {
			if (total + tk > out_cap) return B2D_ERR_OUTPUT_OVERFLOW;
			if (tk) CK(cudaMemcpyAsync(out + total, (uint8_t *)g.out.p + slice_bound * k, tk, cudaMemcpyDeviceToHost, sC));
		}
		total += tk;
		const uint64_t a = (uint64_t)k * slice_in, len = std::min<uint64_t>(slice_in, in_len - a);
		const uint32_t nck = (uint32_t)((len + p.chunk_bytes - 1) / p.chunk_bytes);
		if (chunk_out_len) memcpy(chunk_out_len + (size_t)k * per, hmk + ms_clen, (size_t)nck * 8);
		if (chunk_crc_out) memcpy(chunk_crc_out + (size_t)k * per, hmk + ms_ccrc, (size_t)nck * 4);
	}
	if (block_bits_out && n_chunks)
		CK(cudaMemcpyAsync(block_bits_out, g.bits.p, (size_t)((in_len + p.block_bytes - 1) / p.block_bytes) * 4, cudaMemcpyDeviceToHost, sC));
	for (DevBuf *b : {&g.in, &g.out, &g.meta, &g.bits}) if ((r = release(*b, sC))) return r;
	CK(cudaStreamSynchronize(sC));
	guard.armed = false;
	*total_out = total;
	return B2D_OK;
}

// second half of deflate_pipelined(deliver = false): the payload slices go to their place in the joined stream --
// straight to the host buffer, or (to_dev0) into GPU 0's memory with one cudaMemcpyPeerAsync per slice
int deflate_pipelined_deliver(Ctx &g, const SlicePlan &sp, uint8_t *dst, bool dst_is_dev0, int dev0) {
	CK(cudaSetDevice(g.device));
	cudaStream_t sC = g.st[2];
	uint64_t o = 0;
	for (uint32_t k = 0; k < sp.n_slices; k++) {
		const uint64_t tk = sp.slice_total[k];
		const uint8_t *src = (const uint8_t *)g.out.p + sp.slice_bound * k;
		if (tk) {
			if (dst_is_dev0) CK(cudaMemcpyPeerAsync(dst + o, dev0, src, g.device, tk, sC));
			else CK(cudaMemcpyAsync(dst + o, src, tk, cudaMemcpyDeviceToHost, sC));
		}
		o += tk;
	}
	CK(cudaStreamSynchronize(sC));
	return B2D_OK;
}

// One device, any framing, one shot (small inputs, reference framing).
int deflate_oneshot(Ctx &g, const uint8_t *in, uint64_t in_len, const DeflateParams &p, uint8_t *out, uint64_t out_cap,
                    bool want_crc, uint32_t *chunk_crc_out, uint64_t *chunk_out_len, uint32_t *block_bits_out, uint64_t *total_out) {
	int r;
	const uint64_t bound = deflate_bound_bytes(in_len, p.chunk_bytes, p.block_bytes);
	const uint32_t n_chunks = (uint32_t)((in_len + p.chunk_bytes - 1) / p.chunk_bytes);
	const uint32_t n_blocks = (uint32_t)((in_len + p.block_bytes - 1) / p.block_bytes);
	if ((r = ensure(g.in, in_len + 64))) return r;
	if ((r = ensure(g.out, bound + 64))) return r;
	if (block_bits_out && (r = ensure(g.bits, (size_t)(n_blocks + 1) * 4))) return r;
	const size_t m_clen = 8, m_ccrc = m_clen + (size_t)(n_chunks + 1) * 8, m_total = m_ccrc + (size_t)(n_chunks + 1) * 4;
	if ((r = ensure(g.meta, m_total))) return r;
	if ((r = ensure_pinned_meta(g, m_total))) return r;
	uint8_t *dm = (uint8_t *)g.meta.p, *hm = (uint8_t *)g.pinned_meta;
	cudaStream_t st = g.st[0];
	for (DevBuf *b : {&g.in, &g.out, &g.meta, &g.bits}) if ((r = acquire(*b, st))) return r;
	SyncOnError guard;
	if (in_len) CK(cudaMemcpyAsync(g.in.p, in, in_len, cudaMemcpyHostToDevice, st));
	r = deflate_dev_locked(g, (const uint8_t *)g.in.p, in_len, p, (uint8_t *)g.out.p, bound, (uint64_t *)dm, (uint64_t *)(dm + m_clen),
	                       want_crc ? (uint32_t *)(dm + m_ccrc) : nullptr, st, block_bits_out ? (uint32_t *)g.bits.p : nullptr);
	if (r) return r;
	CK(cudaMemcpyAsync(hm, dm, m_total, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	const uint64_t total = *(uint64_t *)hm;
	if (total > out_cap) return B2D_ERR_OUTPUT_OVERFLOW;
	if (total) CK(cudaMemcpyAsync(out, g.out.p, total, cudaMemcpyDeviceToHost, st));
	if (block_bits_out && n_blocks) CK(cudaMemcpyAsync(block_bits_out, g.bits.p, (size_t)n_blocks * 4, cudaMemcpyDeviceToHost, st));
	if (chunk_out_len) {
		if (p.framing == B2D_FRAMING_REFERENCE) chunk_out_len[0] = total;
		else memcpy(chunk_out_len, hm + m_clen, (size_t)n_chunks * 8);
	}
	if (chunk_crc_out) memcpy(chunk_crc_out, hm + m_ccrc, (size_t)n_chunks * 4);
	for (DevBuf *b : {&g.in, &g.out, &g.meta, &g.bits}) if ((r = release(*b, st))) return r;
	CK(cudaStreamSynchronize(st));
	guard.armed = false;
	*total_out = total;
	return B2D_OK;
}

// The host entry of the compress direction: one device, or the chunks in contiguous ranges over all devices.
int64_t deflate_host_all(const uint8_t *in, uint64_t in_len, const b2d_deflate_opts *opts, uint8_t *out, uint64_t out_cap,
                         uint32_t *crc32_inout, uint64_t *chunk_out_len, uint32_t *block_bits, bool indexed) {
	if (g_ndev == 0) return B2D_ERR_NO_DEVICE;
	DeflateParams p;
	int r = normalise_opts(opts, in_len, p);
	if (r) return r;
	if ((in_len && !in) || !out) return B2D_ERR_BAD_ARGUMENT;
	if (indexed && (!block_bits || p.framing != B2D_FRAMING_CHUNKED)) return B2D_ERR_BAD_ARGUMENT;
	const uint32_t n_chunks = (uint32_t)((in_len + p.chunk_bytes - 1) / p.chunk_bytes);
	const bool want_crc = crc32_inout != nullptr;
	std::vector<uint32_t> ccrc(want_crc ? n_chunks + 1 : 0);
	uint64_t total = 0;
	const int parts = p.framing == B2D_FRAMING_CHUNKED ? (int)std::min<uint64_t>((uint64_t)g_ndev, n_chunks / 64) : 1;
	if (parts <= 1) {
		Ctx &g = g_ctx[0];
		std::lock_guard<std::mutex> lk(g.mu);
		if (!g.ready) return B2D_ERR_NO_DEVICE;
		CK(cudaSetDevice(g.device));
		if (p.framing == B2D_FRAMING_CHUNKED && n_chunks >= 256) {
			SlicePlan sp;
			r = deflate_pipelined(g, in, in_len, p, out, out_cap, want_crc, want_crc ? ccrc.data() : nullptr, chunk_out_len,
			                      indexed ? block_bits : nullptr, true, sp, &total);
		} else {
			r = deflate_oneshot(g, in, in_len, p, out, out_cap, want_crc, want_crc ? ccrc.data() : nullptr, chunk_out_len,
			                    indexed ? block_bits : nullptr, &total);
		}
		if (r) return r;
	} else {
		// Several devices: device d compresses the contiguous chunk range [b[d], b[d+1]); once every
		// range's size is known the offsets are a scan over them, and every device delivers its payload to its place in
		// the joined stream.  B2D_MULTI_GATHER=peer routes the payloads through GPU 0 (cudaMemcpyPeerAsync over NVLink
		// into one buffer there, then one copy to the host) -- what a consumer on GPU 0 would want; for a host
		// consumer the direct copies use every GPU's own PCIe link and are the default.
		const char *gm_ = getenv("B2D_MULTI_GATHER");
		const bool via_dev0 = gm_ && !strcmp(gm_, "peer");
		const uint32_t bpc = p.chunk_bytes / p.block_bytes;
		std::vector<uint32_t> b(parts + 1);
		for (int d = 0; d <= parts; d++) b[d] = (uint32_t)((uint64_t)n_chunks * d / parts);
		std::vector<SlicePlan> plans(parts);
		std::vector<uint64_t> totals(parts, 0);
		std::vector<int> rc(parts, B2D_OK);
		std::vector<std::unique_lock<std::mutex>> locks;
		for (int d = 0; d < parts; d++) locks.emplace_back(g_ctx[d].mu);
		for (int d = 0; d < parts; d++) if (!g_ctx[d].ready) return B2D_ERR_NO_DEVICE;
		{
			std::vector<std::thread> th;
			for (int d = 0; d < parts; d++)
				th.emplace_back([&, d] {
					Ctx &g = g_ctx[d];
					if (cudaSetDevice(g.device) != cudaSuccess) { rc[d] = B2D_ERR_CUDA; return; }
					const uint64_t a = (uint64_t)b[d] * p.chunk_bytes, e = std::min<uint64_t>(in_len, (uint64_t)b[d + 1] * p.chunk_bytes);
					DeflateParams pd = p;
					pd.is_last = (p.is_last && d + 1 == parts) ? 1 : 0;
					rc[d] = deflate_pipelined(g, in + a, e - a, pd, nullptr, 0, want_crc, want_crc ? ccrc.data() + b[d] : nullptr,
					                          chunk_out_len ? chunk_out_len + b[d] : nullptr,
					                          indexed ? block_bits + (size_t)b[d] * bpc : nullptr, false, plans[d], &totals[d]);
				});
			for (auto &t : th) t.join();
		}
		for (int d = 0; d < parts; d++) if (rc[d] != B2D_OK) return rc[d];
		std::vector<uint64_t> off(parts + 1, 0);
		for (int d = 0; d < parts; d++) off[d + 1] = off[d] + totals[d];
		total = off[parts];
		if (total > out_cap) return B2D_ERR_OUTPUT_OVERFLOW;
		uint8_t *gather = nullptr;
		if (via_dev0) {
			Ctx &g0 = g_ctx[0];
			CK(cudaSetDevice(g0.device));
			if ((r = ensure(g0.scratch2, total + 64))) return r;       // (not in use: g0.mu is held)
			if (g0.scratch2.used) CK(cudaEventSynchronize(g0.scratch2.last_use));
			gather = (uint8_t *)g0.scratch2.p;
		}
		{
			std::vector<std::thread> th;
			for (int d = 0; d < parts; d++)
				th.emplace_back([&, d] {
					rc[d] = deflate_pipelined_deliver(g_ctx[d], plans[d], via_dev0 ? gather + off[d] : out + off[d], via_dev0, g_ctx[0].device);
				});
			for (auto &t : th) t.join();
		}
		for (int d = 0; d < parts; d++) if (rc[d] != B2D_OK) return rc[d];
		if (via_dev0 && total) {
			CK(cudaSetDevice(g_ctx[0].device));
			CK(cudaMemcpyAsync(out, gather, total, cudaMemcpyDeviceToHost, g_ctx[0].st[0]));
			CK(cudaStreamSynchronize(g_ctx[0].st[0]));
		}
	}
	if (want_crc) {
		uint32_t crc = *crc32_inout;
		for (uint32_t c = 0; c < n_chunks; c++)
			crc = combine_checksum(p, crc, ccrc[c], std::min<uint64_t>(p.chunk_bytes, in_len - (uint64_t)c * p.chunk_bytes));
		*crc32_inout = crc;
	}
	return (int64_t)total;
}

// Block-parallel decode of our own stream on one device (host pointers).
int inflate_chunks_host(Ctx &g, const uint8_t *in, const uint64_t *chunk_in_len, uint32_t n_chunks, const uint32_t *block_bits,
                        uint32_t chunk_bytes, uint32_t block_bytes, uint8_t *out, uint64_t out_total, uint32_t *chunk_crc32,
                        int32_t *chunk_status, uint32_t flags, std::vector<uint64_t> &off) {
	std::lock_guard<std::mutex> lk(g.mu);
	if (!g.ready) return B2D_ERR_NO_DEVICE;
	const bool want_sum = (flags & (B2D_INFLATE_CRC32 | B2D_INFLATE_ADLER32)) != 0;
	CK(cudaSetDevice(g.device));
	const uint32_t bpc = chunk_bytes / block_bytes;
	off.assign(n_chunks + 1, 0);
	for (uint32_t c = 0; c < n_chunks; c++) off[c + 1] = off[c] + chunk_in_len[c];
	const uint64_t in_total = off[n_chunks];
	int r;
	if ((r = ensure(g.in, in_total + 64))) return r;
	if ((r = ensure(g.out, out_total + 256))) return r;
	if ((r = ensure(g.bits, (size_t)n_chunks * bpc * 4))) return r;
	if ((r = ensure(g.scratch2, inflate_units_scratch_bytes(out_total, chunk_bytes, block_bytes)))) return r;
	const size_t m_off = 0, m_crc = (size_t)(n_chunks + 1) * 8, m_st = m_crc + (size_t)n_chunks * 4, m_total = m_st + (size_t)n_chunks * 4;
	if ((r = ensure(g.meta, m_total))) return r;
	if ((r = ensure_pinned_meta(g, m_total))) return r;
	uint8_t *dm = (uint8_t *)g.meta.p, *hm = (uint8_t *)g.pinned_meta;
	memcpy(hm + m_off, off.data(), (size_t)(n_chunks + 1) * 8);
	cudaStream_t st = g.st[0];
	for (DevBuf *b : {&g.in, &g.out, &g.meta, &g.bits, &g.scratch2}) if ((r = acquire(*b, st))) return r;
	SyncOnError guard;
	CK(cudaMemcpyAsync(dm, hm, m_crc, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(g.bits.p, block_bits, (size_t)n_chunks * bpc * 4, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(g.in.p, in, in_total, cudaMemcpyHostToDevice, st));
	CK(launch_inflate_units((const uint8_t *)g.in.p, (const uint64_t *)(dm + m_off), n_chunks, (const uint32_t *)g.bits.p, chunk_bytes,
	                        block_bytes, out_total, (uint8_t *)g.out.p, (int *)(dm + m_st), g.scratch2.p, st));
	if (want_sum) {
		if (flags & B2D_INFLATE_ADLER32) CK(launch_adler32_pieces((const uint8_t *)g.out.p, out_total, chunk_bytes, n_chunks, (uint32_t *)(dm + m_crc), st));
		else CK(launch_crc32_pieces((const uint8_t *)g.out.p, out_total, chunk_bytes, n_chunks, (uint32_t *)(dm + m_crc), st));
	}
	CK(cudaMemcpyAsync(out, g.out.p, out_total, cudaMemcpyDeviceToHost, st));
	CK(cudaMemcpyAsync(hm + m_crc, dm + m_crc, m_total - m_crc, cudaMemcpyDeviceToHost, st));
	for (DevBuf *b : {&g.in, &g.out, &g.meta, &g.bits, &g.scratch2}) if ((r = release(*b, st))) return r;
	CK(cudaStreamSynchronize(st));
	guard.armed = false;
	memcpy(chunk_status, hm + m_st, (size_t)n_chunks * 4);
	if (want_sum) memcpy(chunk_crc32, hm + m_crc, (size_t)n_chunks * 4);
	return B2D_OK;
}

// One stream without an index, decoded speculatively in parallel (inflate.cu, "ONE foreign stream"); device pointers.
int inflate_stream_dev_locked(Ctx &g, const uint8_t *d_in, uint64_t in_len, uint8_t *d_out, uint64_t out_cap, uint64_t *d_result,
                              cudaStream_t st, uint32_t cap_scale = 1) {
	int r = ensure(g.scratch2, inflate_stream_scratch_bytes(in_len, cap_scale));
	if (r) return r;
	if ((r = acquire(g.scratch2, st))) return r;
	CK(launch_inflate_stream(d_in, in_len, d_out, out_cap, d_result, g.scratch2.p, st, cap_scale));
	return release(g.scratch2, st);
}

int checksum_host(bool adler, const uint8_t *data, uint64_t len, uint32_t *inout) {
	Ctx *gp = primary();
	if (!gp) { set_error("b2d_init not called or failed%s%s", "", ""); return B2D_ERR_NO_DEVICE; }
	Ctx &g = *gp;
	std::lock_guard<std::mutex> lk(g.mu);
	if (!inout || (len && !data)) return B2D_ERR_BAD_ARGUMENT;
	if (len == 0) return B2D_OK;
	CK(cudaSetDevice(g.device));
	int r;
	if ((r = ensure(g.in, len + 64))) return r;
	const uint64_t piece = 1u << 20;
	const uint32_t n_pieces = (uint32_t)((len + piece - 1) / piece);
	if ((r = ensure(g.crc, (size_t)(n_pieces + 1) * 4))) return r;
	if ((r = ensure_pinned_meta(g, (size_t)(n_pieces + 1) * 4))) return r;
	cudaStream_t st = g.st[0];
	for (DevBuf *b : {&g.in, &g.crc}) if ((r = acquire(*b, st))) return r;
	SyncOnError guard;
	CK(cudaMemcpyAsync(g.in.p, data, len, cudaMemcpyHostToDevice, st));
	if (adler) CK(launch_adler32_pieces((const uint8_t *)g.in.p, len, piece, n_pieces, (uint32_t *)g.crc.p, st));
	else CK(launch_crc32_pieces((const uint8_t *)g.in.p, len, piece, n_pieces, (uint32_t *)g.crc.p, st));
	CK(cudaMemcpyAsync(g.pinned_meta, g.crc.p, (size_t)n_pieces * 4, cudaMemcpyDeviceToHost, st));
	for (DevBuf *b : {&g.in, &g.crc}) if ((r = release(*b, st))) return r;
	CK(cudaStreamSynchronize(st));
	guard.armed = false;
	const uint32_t *pc = (const uint32_t *)g.pinned_meta;
	uint32_t v = *inout;
	for (uint32_t i = 0; i < n_pieces; i++) {
		const uint64_t l = std::min<uint64_t>(piece, len - (uint64_t)i * piece);
		v = adler ? host_adler32_combine(v, pc[i], l) : host_crc32_combine(v, pc[i], l);
	}
	*inout = v;
	return B2D_OK;
}

// GzipMetadata.read (GzipMetadata.java:73-146): validates a member header in the reference's order and returns its
// length, or a status (1 + Reason.ordinal()).
int parse_gzip_header(const uint8_t *p, uint64_t n, uint64_t *hdr_len) {
	uint64_t i = 0;
	auto need = [&](uint64_t k) { return i + k <= n; };
	if (!need(2)) return B2D_UNEXPECTED_END_OF_STREAM;
	if (p[0] != 0x1F || p[1] != 0x8B) return B2D_GZIP_INVALID_MAGIC_NUMBER;
	if (!need(3)) return B2D_UNEXPECTED_END_OF_STREAM;
	if (p[2] != 8) return B2D_UNSUPPORTED_COMPRESSION_METHOD;
	if (!need(4)) return B2D_UNEXPECTED_END_OF_STREAM;
	const int flags = p[3];
	if (flags & 0xE0) return B2D_GZIP_RESERVED_FLAGS_SET;
	if (!need(10)) return B2D_UNEXPECTED_END_OF_STREAM;
	const int os = p[9];
	if (os > 13 && os != 0xFF) return B2D_GZIP_UNSUPPORTED_OPERATING_SYSTEM;
	i = 10;
	if (flags & 4) {
		if (!need(2)) return B2D_UNEXPECTED_END_OF_STREAM;
		uint64_t xl = p[i] | (uint64_t)p[i + 1] << 8;
		i += 2;
		if (!need(xl)) return B2D_UNEXPECTED_END_OF_STREAM;
		i += xl;
	}
	for (int f = 8; f <= 16; f <<= 1) {
		if (!(flags & f)) continue;
		for (;;) {
			if (!need(1)) return B2D_UNEXPECTED_END_OF_STREAM;
			if (p[i++] == 0) break;
		}
	}
	if (flags & 2) {
		if (!need(2)) return B2D_UNEXPECTED_END_OF_STREAM;
		uint32_t expect = host_crc32_bytes(0, p, (size_t)i) & 0xFFFF;
		uint32_t actual = p[i] | (uint32_t)p[i + 1] << 8;
		if (actual != expect) return B2D_HEADER_CHECKSUM_MISMATCH;
		i += 2;
	}
	*hdr_len = i;
	return B2D_OK;
}

}  // namespace

extern "C" {

#define B2D_API __attribute__((visibility("default")))

B2D_API int b2d_init(int device) {
	std::lock_guard<std::mutex> lk(g_mu);
	return init_devices_locked(&device, 1);
}

B2D_API int b2d_init_devices(const int *devices, int n) {
	std::lock_guard<std::mutex> lk(g_mu);
	if (n <= 0) {                                  // every sm_100 device of the box
		int count = 0;
		if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
			cudaGetLastError();
			set_error("no usable CUDA device%s%s", "", "");
			return B2D_ERR_NO_DEVICE;
		}
		int all[MAX_DEVICES];
		count = std::min(count, MAX_DEVICES);
		for (int i = 0; i < count; i++) all[i] = i;
		return init_devices_locked(all, count);
	}
	if (!devices || n > MAX_DEVICES) return B2D_ERR_BAD_ARGUMENT;
	for (int i = 0; i < n; i++)
		for (int j = 0; j < i; j++) if (devices[i] == devices[j]) return B2D_ERR_BAD_ARGUMENT;
	return init_devices_locked(devices, n);
}

B2D_API int b2d_device_count(void) { return g_ndev; }

B2D_API void b2d_shutdown(void) {
	std::lock_guard<std::mutex> lk(g_mu);
	for (int i = 0; i < g_ndev; i++) {
		std::lock_guard<std::mutex> lk2(g_ctx[i].mu);
		if (g_ctx[i].ready) {
			cudaSetDevice(g_ctx[i].device);
			cudaDeviceSynchronize();
		}
		release_ctx(g_ctx[i]);
	}
	g_ndev = 0;
}

B2D_API uint64_t b2d_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }

B2D_API const char *b2d_strerror(int status) {
	switch (status) {
	case B2D_OK: return "OK";
	case B2D_UNEXPECTED_END_OF_STREAM: return "Unexpected end of stream";
	case B2D_RESERVED_BLOCK_TYPE: return "Reserved block type";
	case B2D_UNCOMPRESSED_BLOCK_LENGTH_MISMATCH: return "len/nlen mismatch in uncompressed block";
	case B2D_HUFFMAN_CODE_UNDER_FULL: return "This canonical code produces an under-full Huffman code tree";
	case B2D_HUFFMAN_CODE_OVER_FULL: return "This canonical code produces an over-full Huffman code tree";
	case B2D_NO_PREVIOUS_CODE_LENGTH_TO_COPY: return "No code length value to copy";
	case B2D_CODE_LENGTH_CODE_OVER_FULL: return "Run exceeds number of codes";
	case B2D_END_OF_BLOCK_CODE_ZERO_LENGTH: return "End-of-block symbol has zero code length";
	case B2D_RESERVED_LENGTH_SYMBOL: return "Reserved run length symbol";
	case B2D_RESERVED_DISTANCE_SYMBOL: return "Reserved distance symbol";
	case B2D_LENGTH_ENCOUNTERED_WITH_EMPTY_DISTANCE_CODE: return "Length symbol encountered with empty distance code";
	case B2D_COPY_FROM_BEFORE_DICTIONARY_START: return "Attempting to copy from before start of dictionary";
	case B2D_HEADER_CHECKSUM_MISMATCH: return "Header CRC-16 mismatch";
	case B2D_UNSUPPORTED_COMPRESSION_METHOD: return "Unsupported compression method";
	case B2D_DECOMPRESSED_CHECKSUM_MISMATCH: return "Decompression CRC-32 mismatch";
	case B2D_DECOMPRESSED_SIZE_MISMATCH: return "Decompressed size mismatch";
	case B2D_GZIP_INVALID_MAGIC_NUMBER: return "Invalid GZIP magic number";
	case B2D_GZIP_RESERVED_FLAGS_SET: return "Reserved flags are set";
	case B2D_GZIP_UNSUPPORTED_OPERATING_SYSTEM: return "Unsupported operating system value";
	case B2D_ERR_OUTPUT_OVERFLOW: return "Output capacity exceeded";
	case B2D_ERR_BAD_ARGUMENT: return "Bad argument";
	case B2D_ERR_NO_DEVICE: return "No usable sm_100 GPU (b2d_init not called or failed); there is no CPU fallback";
	case B2D_ERR_CUDA: return "CUDA runtime failure";
	case B2D_ERR_OUT_OF_MEMORY: return "Out of device memory";
	default: return "Unknown status";
	}
}

B2D_API const char *b2d_last_error(void) { return g_last_error; }

B2D_API int b2d_device_sm_count(void) { Ctx *g = primary(); return g ? g->sm_count : 0; }

B2D_API void *b2d_alloc_pinned(size_t bytes) {
	Ctx *g = primary();
	if (!g) return nullptr;
	std::lock_guard<std::mutex> lk(g->mu);
	if (cudaSetDevice(g->device) != cudaSuccess) { cudaGetLastError(); return nullptr; }
	void *p = nullptr;
	// portable: with several devices every GPU's kernels write their members' output into the same buffer
	if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) { cudaGetLastError(); return nullptr; }
	return p;
}

B2D_API void b2d_free_pinned(void *p) {
	if (p) cudaFreeHost(p);
}

// ---------------------------------------------------------------- inflate

B2D_API int b2d_inflate_batch_dev(const uint8_t *d_in, const uint64_t *d_in_off, uint32_t n, uint8_t *d_out,
                                  const uint64_t *d_out_off, uint64_t *d_out_len, uint64_t *d_in_consumed,
                                  uint32_t *d_crc32, int32_t *d_status, uint32_t flags, void *stream) {
	Ctx *g = ctx_for_pointer(d_out);
	if (!g) return B2D_ERR_NO_DEVICE;
	std::lock_guard<std::mutex> lk(g->mu);
	if (!g->ready) return B2D_ERR_NO_DEVICE;
	if (n && (!d_in || !d_in_off || !d_out || !d_out_off || !d_out_len || !d_in_consumed || !d_status))
		return B2D_ERR_BAD_ARGUMENT;
	if ((flags & (B2D_INFLATE_CRC32 | B2D_INFLATE_ADLER32)) && !d_crc32 && n) return B2D_ERR_BAD_ARGUMENT;
	CK(cudaSetDevice(g->device));
	cudaStream_t st = (cudaStream_t)stream;
	return inflate_dev_locked(d_in, d_in_off, n, d_out, d_out_off, d_out_len, d_in_consumed, d_crc32, d_status, flags, st);
}

B2D_API int b2d_inflate_batch(const uint8_t *in, const uint64_t *in_off, uint32_t n, uint8_t *out,
                              const uint64_t *out_off, uint64_t *out_len, uint64_t *in_consumed, uint32_t *crc32,
                              int32_t *status, uint32_t flags) {
	if (g_ndev == 0) return B2D_ERR_NO_DEVICE;
	if (n == 0) return B2D_OK;
	if (!in_off) return B2D_ERR_BAD_ARGUMENT;
	return inflate_host_all(in, in_off, in_off + 1, n, out, out_off, out_len, in_consumed, crc32, status, flags);
}

// ---------------------------------------------------------------- deflate

B2D_API uint64_t b2d_deflate_bound(uint64_t in_len, uint32_t chunk_bytes) {
	return deflate_bound_bytes(in_len, chunk_bytes ? chunk_bytes : (1u << 20), 1u << 16);
}

B2D_API int b2d_deflate_chunks_dev(const uint8_t *d_in, uint64_t in_len, const b2d_deflate_opts *opts, uint8_t *d_out,
                                   uint64_t out_cap, uint64_t *d_out_len_total, uint64_t *d_chunk_out_len,
                                   uint32_t *d_chunk_crc32, void *stream) {
	Ctx *g = ctx_for_pointer(d_out);
	if (!g) return B2D_ERR_NO_DEVICE;
	std::lock_guard<std::mutex> lk(g->mu);
	if (!g->ready) return B2D_ERR_NO_DEVICE;
	DeflateParams p;
	int r = normalise_opts(opts, in_len, p);
	if (r) return r;
	if ((in_len && !d_in) || !d_out || !d_out_len_total) return B2D_ERR_BAD_ARGUMENT;
	if (out_cap < deflate_bound_bytes(in_len, p.chunk_bytes, p.block_bytes)) return B2D_ERR_OUTPUT_OVERFLOW;
	CK(cudaSetDevice(g->device));
	cudaStream_t st = (cudaStream_t)stream;
	return deflate_dev_locked(*g, d_in, in_len, p, d_out, out_cap, d_out_len_total, d_chunk_out_len, d_chunk_crc32, st);
}

B2D_API int64_t b2d_deflate_chunks(const uint8_t *in, uint64_t in_len, const b2d_deflate_opts *opts, uint8_t *out,
                                   uint64_t out_cap, uint32_t *crc32_inout, uint64_t *chunk_out_len) {
	return deflate_host_all(in, in_len, opts, out, out_cap, crc32_inout, chunk_out_len, nullptr, false);
}

// ---------------------------------------------------------------- block-indexed streams (our own, fully parallel decode)

B2D_API int b2d_deflate_chunks_indexed_dev(const uint8_t *d_in, uint64_t in_len, const b2d_deflate_opts *opts, uint8_t *d_out,
                                           uint64_t out_cap, uint64_t *d_out_len_total, uint64_t *d_chunk_out_len,
                                           uint32_t *d_chunk_crc32, uint32_t *d_block_bits, void *stream) {
	Ctx *g = ctx_for_pointer(d_out);
	if (!g) return B2D_ERR_NO_DEVICE;
	std::lock_guard<std::mutex> lk(g->mu);
	if (!g->ready) return B2D_ERR_NO_DEVICE;
	DeflateParams p;
	int r = normalise_opts(opts, in_len, p);
	if (r) return r;
	if ((in_len && !d_in) || !d_out || !d_out_len_total || !d_block_bits || p.framing != B2D_FRAMING_CHUNKED) return B2D_ERR_BAD_ARGUMENT;
	if (out_cap < deflate_bound_bytes(in_len, p.chunk_bytes, p.block_bytes)) return B2D_ERR_OUTPUT_OVERFLOW;
	CK(cudaSetDevice(g->device));
	return deflate_dev_locked(*g, d_in, in_len, p, d_out, out_cap, d_out_len_total, d_chunk_out_len, d_chunk_crc32,
	                          (cudaStream_t)stream, d_block_bits);
}

B2D_API int b2d_inflate_chunks_dev(const uint8_t *d_in, const uint64_t *d_chunk_in_off, uint32_t n_chunks,
                                   const uint32_t *d_block_bits, uint32_t chunk_bytes, uint32_t block_bytes, uint64_t out_total,
                                   uint8_t *d_out, uint32_t *d_chunk_crc32, int32_t *d_chunk_status, uint32_t flags, void *stream) {
	Ctx *gp = ctx_for_pointer(d_out);
	if (!gp) return B2D_ERR_NO_DEVICE;
	Ctx &g = *gp;
	std::lock_guard<std::mutex> lk(g.mu);
	if (!g.ready) return B2D_ERR_NO_DEVICE;
	if (n_chunks == 0) return B2D_OK;
	if (!d_in || !d_chunk_in_off || !d_block_bits || !d_out || !d_chunk_status || block_bytes < 4096 || block_bytes > (1u << 20) ||
	    chunk_bytes % block_bytes != 0 || (uint64_t)n_chunks * chunk_bytes < out_total ||
	    (uint64_t)(n_chunks - 1) * chunk_bytes >= out_total)
		return B2D_ERR_BAD_ARGUMENT;
	if ((flags & (B2D_INFLATE_CRC32 | B2D_INFLATE_ADLER32)) && !d_chunk_crc32) return B2D_ERR_BAD_ARGUMENT;
	CK(cudaSetDevice(g.device));
	cudaStream_t st = (cudaStream_t)stream;
	int r = ensure(g.scratch2, inflate_units_scratch_bytes(out_total, chunk_bytes, block_bytes));
	if (r) return r;
	if ((r = acquire(g.scratch2, st))) return r;
	CK(launch_inflate_units(d_in, d_chunk_in_off, n_chunks, d_block_bits, chunk_bytes, block_bytes, out_total, d_out,
	                        d_chunk_status, g.scratch2.p, st));
	if ((r = release(g.scratch2, st))) return r;
	if (flags & (B2D_INFLATE_CRC32 | B2D_INFLATE_ADLER32)) {
		if (flags & B2D_INFLATE_ADLER32) CK(launch_adler32_pieces(d_out, out_total, chunk_bytes, n_chunks, d_chunk_crc32, st));
		else CK(launch_crc32_pieces(d_out, out_total, chunk_bytes, n_chunks, d_chunk_crc32, st));
	}
	return B2D_OK;
}

// Host form of the pair above.  Compress: like b2d_deflate_chunks plus the block index.  Decompress: chunk sizes + block
// index in, bytes out; a chunk whose block-parallel decode reports a problem is decoded again serially
// (B2D_INFLATE_CHUNK_INDEXED) so that status and delivered bytes are exactly the sequential decoder's.
B2D_API int64_t b2d_deflate_chunks_indexed(const uint8_t *in, uint64_t in_len, const b2d_deflate_opts *opts, uint8_t *out,
                                           uint64_t out_cap, uint32_t *crc32_inout, uint64_t *chunk_out_len, uint32_t *block_bits) {
	return deflate_host_all(in, in_len, opts, out, out_cap, crc32_inout, chunk_out_len, block_bits, true);
}

B2D_API int b2d_inflate_chunks(const uint8_t *in, const uint64_t *chunk_in_len, uint32_t n_chunks, const uint32_t *block_bits,
                               uint32_t chunk_bytes, uint32_t block_bytes, uint8_t *out, uint64_t out_total,
                               uint32_t *chunk_crc32, int32_t *chunk_status, uint32_t flags) {
	if (g_ndev == 0) return B2D_ERR_NO_DEVICE;
	if (n_chunks == 0) return B2D_OK;
	if (!in || !chunk_in_len || !block_bits || !out || !chunk_status || block_bytes < 4096 || chunk_bytes % block_bytes != 0 ||
	    (uint64_t)n_chunks * chunk_bytes < out_total || (uint64_t)(n_chunks - 1) * chunk_bytes >= out_total)
		return B2D_ERR_BAD_ARGUMENT;
	const bool want_sum = (flags & (B2D_INFLATE_CRC32 | B2D_INFLATE_ADLER32)) != 0;
	if (want_sum && !chunk_crc32) return B2D_ERR_BAD_ARGUMENT;
	const uint32_t bpc = chunk_bytes / block_bytes;
	std::vector<uint64_t> goff(n_chunks + 1, 0);                  // chunk offsets in `in`
	for (uint32_t c = 0; c < n_chunks; c++) goff[c + 1] = goff[c] + chunk_in_len[c];
	const int parts = (int)std::min<uint64_t>((uint64_t)g_ndev, std::max<uint32_t>(1, n_chunks / 64));
	int r = B2D_OK;
	if (parts <= 1) {
		std::vector<uint64_t> off;
		r = inflate_chunks_host(g_ctx[0], in, chunk_in_len, n_chunks, block_bits, chunk_bytes, block_bytes, out, out_total,
		                        chunk_crc32, chunk_status, flags, off);
	} else {
		std::vector<int> rc(parts, B2D_OK);
		std::vector<std::thread> th;
		for (int d = 0; d < parts; d++)
			th.emplace_back([&, d] {
				const uint32_t a = (uint32_t)((uint64_t)n_chunks * d / parts), e = (uint32_t)((uint64_t)n_chunks * (d + 1) / parts);
				if (e <= a) return;
				const uint64_t o0 = (uint64_t)a * chunk_bytes, o1 = std::min<uint64_t>(out_total, (uint64_t)e * chunk_bytes);
				std::vector<uint64_t> off;
				rc[d] = inflate_chunks_host(g_ctx[d], in + goff[a], chunk_in_len + a, e - a, block_bits + (size_t)a * bpc, chunk_bytes,
				                            block_bytes, out + o0, o1 - o0, chunk_crc32 ? chunk_crc32 + a : nullptr, chunk_status + a,
				                            flags, off);
			});
		for (auto &t : th) t.join();
		for (int d = 0; d < parts; d++) if (rc[d] != B2D_OK) r = rc[d];
	}
	if (r) return r;
	// exact outcome for chunks the parallel decode could not finish: one serial decode each
	for (uint32_t c = 0; c < n_chunks; c++) {
		if (chunk_status[c] == 0) continue;
		const uint64_t io[2] = {goff[c], goff[c + 1]};
		const uint64_t o0 = (uint64_t)c * chunk_bytes, oo[2] = {o0, std::min<uint64_t>(out_total, o0 + chunk_bytes)};
		uint64_t ol = 0, ic = 0;
		uint32_t cr = 0;
		int32_t s2 = 0;
		r = inflate_host(g_ctx[0], in, io, io + 1, 1, out, oo, &ol, &ic, &cr, &s2,
		                 (flags & (B2D_INFLATE_CRC32 | B2D_INFLATE_ADLER32)) | B2D_INFLATE_CHUNK_INDEXED);
		if (r) return r;
		chunk_status[c] = s2 != 0 ? s2 : (ol == oo[1] - oo[0] ? 0 : B2D_UNEXPECTED_END_OF_STREAM);
		if (want_sum) chunk_crc32[c] = cr;
	}
	return B2D_OK;
}

// ---------------------------------------------------------------- one stream from any producer, decoded in parallel

// Device pointers: d_in 4-byte aligned and readable to the next 16-byte boundary past in_len; d_result receives five
// u64 {out_len, in_consumed, status, 0, units}.  status != 0 means "not decodable in parallel, or not valid": decode it
// with b2d_inflate_batch_dev for the reference's exact outcome.
B2D_API int b2d_inflate_stream_dev(const uint8_t *d_in, uint64_t in_len, uint8_t *d_out, uint64_t out_cap, uint64_t *d_result, void *stream) {
	Ctx *gp = ctx_for_pointer(d_out);
	if (!gp) return B2D_ERR_NO_DEVICE;
	Ctx &g = *gp;
	std::lock_guard<std::mutex> lk(g.mu);
	if (!g.ready) return B2D_ERR_NO_DEVICE;
	if (!d_in || !d_result || (out_cap && !d_out) || ((uintptr_t)d_in & 3)) return B2D_ERR_BAD_ARGUMENT;
	CK(cudaSetDevice(g.device));
	return inflate_stream_dev_locked(g, d_in, in_len, d_out, out_cap, d_result, (cudaStream_t)stream);
}

// Host pointers.  Replaces InflaterInputStream.read -> Open.read (Open.java:83-110) for ONE raw-DEFLATE stream of any
// origin: results (bytes, out_len, in_consumed, status) are the sequential decoder's -- when the parallel decode cannot
// vouch for them the stream is decoded again by b2d_inflate_batch's one-warp decoder.  *parallel (optional) says which.
B2D_API int b2d_inflate_stream(const uint8_t *in, uint64_t in_len, uint8_t *out, uint64_t out_cap, uint64_t *out_len,
                               uint64_t *in_consumed, uint32_t *crc32, int32_t *status, uint32_t flags, int32_t *parallel) {
	Ctx *gp = primary();
	if (!gp) return B2D_ERR_NO_DEVICE;
	Ctx &g = *gp;
	if ((in_len && !in) || (out_cap && !out) || !out_len || !in_consumed || !status) return B2D_ERR_BAD_ARGUMENT;
	if ((flags & (B2D_INFLATE_CRC32 | B2D_INFLATE_ADLER32)) && !crc32) return B2D_ERR_BAD_ARGUMENT;
	if (parallel) *parallel = 0;
	bool done = false;
	const char *off_ = getenv("B2D_STREAM_PARALLEL");
	if (in_len >= (1u << 20) && !(off_ && off_[0] == '0')) {
		std::lock_guard<std::mutex> lk(g.mu);
		if (!g.ready) return B2D_ERR_NO_DEVICE;
		CK(cudaSetDevice(g.device));
		int r;
		if ((r = ensure(g.in, in_len + 64))) return r;
		if ((r = ensure(g.out, out_cap + 256))) return r;
		if ((r = ensure(g.meta, 64))) return r;
		if ((r = ensure_pinned_meta(g, 64 + (size_t)((out_cap >> 20) + 2) * 4))) return r;
		cudaStream_t st = g.st[0];
		for (DevBuf *b : {&g.in, &g.out, &g.meta}) if ((r = acquire(*b, st))) return r;
		SyncOnError guard;
		const char *tr_ = getenv("B2D_TRACE");
		const bool trace = tr_ && tr_[0] == '1';
		const auto t0 = std::chrono::steady_clock::now();
		auto lap = [&](const char *what) {
			if (trace) fprintf(stderr, "[b2d trace] inflate_stream: %s at %.3f ms\n", what,
			                   std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
		};
		CK(cudaMemcpyAsync(g.in.p, in, in_len, cudaMemcpyHostToDevice, st));
		// The units' buffers are sized for compression ratios up to ~15 inside a unit; a unit that outgrows its buffer
		// (long runs, long stretches of stored blocks without a dynamic block to restart at) is given 8 times as much
		// once, memory permitting, before the stream goes to the sequential decoder.
		uint64_t res_out = 0, res_cons = 0, res_status = 1;
		for (uint32_t scale = 1; scale <= 8; scale *= 8) {
			size_t free_b = 0, total_b = 0;
			cudaMemGetInfo(&free_b, &total_b);
			if (inflate_stream_scratch_bytes(in_len, scale) > free_b + g.scratch2.cap - (free_b >> 4)) break;
			if ((r = inflate_stream_dev_locked(g, (const uint8_t *)g.in.p, in_len, (uint8_t *)g.out.p, out_cap, (uint64_t *)g.meta.p, st, scale))) return r;
			lap("kernels enqueued");
			CK(cudaMemcpyAsync(g.pinned_meta, g.meta.p, 40, cudaMemcpyDeviceToHost, st));
			CK(cudaStreamSynchronize(st));
			lap("result on the host");
			res_out = ((const uint64_t *)g.pinned_meta)[0]; res_cons = ((const uint64_t *)g.pinned_meta)[1];
			res_status = ((const uint64_t *)g.pinned_meta)[2];
			if ((int64_t)res_status != B2D_ERR_OUTPUT_OVERFLOW || res_out > out_cap) break;
		}
		{
			if (res_status == 0 && res_out <= out_cap) {
				const uint64_t n = res_out;
				uint32_t sum = (flags & B2D_INFLATE_ADLER32) ? 1u : 0u;
				if (n) CK(cudaMemcpyAsync(out, g.out.p, n, cudaMemcpyDeviceToHost, st));
				if (n && (flags & (B2D_INFLATE_CRC32 | B2D_INFLATE_ADLER32))) {
					const bool adler = (flags & B2D_INFLATE_ADLER32) != 0;
					const uint64_t piece = 1u << 20;
					const uint32_t n_pieces = (uint32_t)((n + piece - 1) / piece);
					if ((r = ensure(g.crc, (size_t)(n_pieces + 1) * 4))) return r;
					if ((r = acquire(g.crc, st))) return r;
					if (adler) CK(launch_adler32_pieces((const uint8_t *)g.out.p, n, piece, n_pieces, (uint32_t *)g.crc.p, st));
					else CK(launch_crc32_pieces((const uint8_t *)g.out.p, n, piece, n_pieces, (uint32_t *)g.crc.p, st));
					CK(cudaMemcpyAsync((uint8_t *)g.pinned_meta + 64, g.crc.p, (size_t)n_pieces * 4, cudaMemcpyDeviceToHost, st));
					if ((r = release(g.crc, st))) return r;
					CK(cudaStreamSynchronize(st));
					const uint32_t *pc = (const uint32_t *)((const uint8_t *)g.pinned_meta + 64);
					for (uint32_t i = 0; i < n_pieces; i++) {
						const uint64_t l = std::min<uint64_t>(piece, n - (uint64_t)i * piece);
						sum = adler ? host_adler32_combine(sum, pc[i], l) : host_crc32_combine(sum, pc[i], l);
					}
				}
				*out_len = n;
				if (crc32) *crc32 = sum;
				done = true;
			}
			for (DevBuf *b : {&g.in, &g.out, &g.meta}) if ((r = release(*b, st))) return r;
			CK(cudaStreamSynchronize(st));
			lap("output and checksum on the host");
			guard.armed = false;
			if (done) {
				*in_consumed = res_cons;
				*status = 0;
				if (parallel) *parallel = 1;
			}
		}
	}
	if (done) return B2D_OK;
	// sequential: one warp, the reference's decoder restated (exact status, bytes delivered before a failure, consumed input)
	const uint64_t io[2] = {0, in_len}, oo[2] = {0, out_cap};
	uint32_t c = 0;
	int r = inflate_host(g, in, io, io + 1, 1, out, oo, out_len, in_consumed, &c, status, flags);
	if (crc32) *crc32 = c;
	return r;
}

// ---------------------------------------------------------------- gzip members

// ISIZE (mod 2^32) of each member, read from its last 4 bytes: what a caller sizes the output slots with.
B2D_API int b2d_gzip_isize(const uint8_t *in, const uint64_t *in_off, uint32_t n, uint64_t *isize) {
	if (n && (!in || !in_off || !isize)) return B2D_ERR_BAD_ARGUMENT;
	for (uint32_t i = 0; i < n; i++) {
		if (in_off[i + 1] < in_off[i]) return B2D_ERR_BAD_ARGUMENT;
		const uint64_t len = in_off[i + 1] - in_off[i];
		const uint8_t *e = in + in_off[i + 1];
		isize[i] = len >= 18 ? ((uint64_t)e[-4] | (uint64_t)e[-3] << 8 | (uint64_t)e[-2] << 16 | (uint64_t)e[-1] << 24) : 0;
	}
	return B2D_OK;
}

// n independent gzip members, each decoded like `new GzipInputStream(in)` read to the end (GzipInputStream.java:38-90):
// header validation on the host (tens of bytes), DEFLATE body + CRC-32 on the GPU, then the trailer checks in the
// reference's order (CRC first, then ISIZE mod 2^32).  Member i's DEFLATE data may use the bytes from the end of its
// header to in_off[i + 1] -- the end of ITS stream, as for the reference reading that member alone -- and never a
// neighbour's bytes: a truncated member ends in UNEXPECTED_END_OF_STREAM with the bytes decoded so far.
B2D_API int b2d_gunzip_batch(const uint8_t *in, const uint64_t *in_off, uint32_t n, uint8_t *out, const uint64_t *out_off,
                             uint64_t *out_len, uint64_t *in_consumed, int32_t *status) {
	if (n == 0) return B2D_OK;
	if (!in || !in_off || !out_off || !out_len || !in_consumed || !status) return B2D_ERR_BAD_ARGUMENT;
	std::vector<uint64_t> body(n), body_end(n), hdr(n, 0);
	std::vector<int32_t> hst(n, 0);
	std::vector<uint32_t> crc(n);
	for (uint32_t i = 0; i < n; i++) {
		if (in_off[i + 1] < in_off[i]) return B2D_ERR_BAD_ARGUMENT;
		hst[i] = parse_gzip_header(in + in_off[i], in_off[i + 1] - in_off[i], &hdr[i]);
		body_end[i] = in_off[i + 1];
		body[i] = hst[i] ? in_off[i + 1] : in_off[i] + hdr[i];      // a bad header decodes nothing (empty range)
	}
	int r = inflate_host_all(in, body.data(), body_end.data(), n, out, out_off, out_len, in_consumed, crc.data(), status, B2D_INFLATE_CRC32);
	if (r != B2D_OK) return r;
	for (uint32_t i = 0; i < n; i++) {
		if (hst[i]) { status[i] = hst[i]; out_len[i] = 0; in_consumed[i] = 0; continue; }
		if (status[i] != 0) { in_consumed[i] += hdr[i]; continue; }
		const uint64_t end = in_off[i] + hdr[i] + in_consumed[i];
		if (end + 8 > in_off[i + 1]) { status[i] = B2D_UNEXPECTED_END_OF_STREAM; in_consumed[i] += hdr[i]; continue; }
		const uint8_t *t = in + end;
		const uint32_t ec = (uint32_t)t[0] | (uint32_t)t[1] << 8 | (uint32_t)t[2] << 16 | (uint32_t)t[3] << 24;
		const uint32_t el = (uint32_t)t[4] | (uint32_t)t[5] << 8 | (uint32_t)t[6] << 16 | (uint32_t)t[7] << 24;
		if (crc[i] != ec) status[i] = B2D_DECOMPRESSED_CHECKSUM_MISMATCH;
		else if ((uint32_t)out_len[i] != el) status[i] = B2D_DECOMPRESSED_SIZE_MISMATCH;
		in_consumed[i] += hdr[i] + 8;
	}
	return B2D_OK;
}

// ---------------------------------------------------------------- CRC-32

B2D_API uint32_t b2d_crc32_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b) {
	return host_crc32_combine(crc_a, crc_b, len_b);
}

B2D_API int b2d_crc32_dev(const uint8_t *d_data, uint64_t len, uint32_t *d_crc_out, void *stream) {
	Ctx *gp = ctx_for_pointer(d_crc_out);
	if (!gp) return B2D_ERR_NO_DEVICE;
	Ctx &g = *gp;
	std::lock_guard<std::mutex> lk(g.mu);
	if (!g.ready) return B2D_ERR_NO_DEVICE;
	if ((len && !d_data) || !d_crc_out) return B2D_ERR_BAD_ARGUMENT;
	CK(cudaSetDevice(g.device));
	cudaStream_t st = (cudaStream_t)stream;
	const uint64_t piece = 1u << 20;
	const uint32_t n_pieces = (uint32_t)((len + piece - 1) / piece);
	int r = ensure(g.crc, (size_t)(n_pieces + 1) * 4);
	if (r) return r;
	if ((r = acquire(g.crc, st))) return r;
	CK(launch_crc32_pieces(d_data, len, piece, n_pieces, (uint32_t *)g.crc.p, st));
	CK(launch_crc32_fold(( const uint32_t *)g.crc.p, n_pieces, piece, len, d_crc_out, st));
	return release(g.crc, st);
}

B2D_API uint32_t b2d_adler32_combine(uint32_t adler_a, uint32_t adler_b, uint64_t len_b) {
	return host_adler32_combine(adler_a, adler_b, len_b);
}

// Checksums of host buffers.  The *_update forms report failure (no device, CUDA error); the value-returning forms
// are conveniences that leave the checksum unchanged on failure and record why in b2d_last_error().
B2D_API int b2d_adler32_update(const uint8_t *data, uint64_t len, uint32_t *adler_inout) {
	return checksum_host(true, data, len, adler_inout);
}

B2D_API int b2d_crc32_update(const uint8_t *data, uint64_t len, uint32_t *crc_inout) {
	return checksum_host(false, data, len, crc_inout);
}

B2D_API uint32_t b2d_adler32(uint32_t adler, const uint8_t *data, uint64_t len) {
	checksum_host(true, data, len, &adler);
	return adler;
}

B2D_API uint32_t b2d_crc32(uint32_t crc, const uint8_t *data, uint64_t len) {
	checksum_host(false, data, len, &crc);
	return crc;
}

}  // extern "C"
