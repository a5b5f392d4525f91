// kernels.h -- internal launch interface between the C-ABI host runtime (api.cu) and the kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/b2deflate.h"

namespace b2d {

// Every kernel launch of the library goes through B2D_LAUNCH, which counts it (b2d_kernel_launches(): bench.py's
// gpu_launches is read from here, not derived by hand).  Device-dependent one-time state (cudaFuncSetAttribute,
// __constant__ uploads) is kept per CUDA ordinal, so b2d_init on another device or several devices at once is fine.
constexpr int MAX_DEVICES = 16;
void count_launch();
#define B2D_LAUNCH(kernel, grid, block, smem, stream) ::b2d::count_launch(), kernel<<<(grid), (block), (smem), (stream)>>>
inline int current_device_slot() {
	int d = 0;
	if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= MAX_DEVICES) d = 0;
	return d;
}

// inflate.cu
// runs a one-warp kernel with the decoder's static shared-memory declaration and reports where the segment starts in the
// shared window; the decoder's LUT addressing (see inflate.cu) needs it at INFLATE_SMEM_WINDOW_BASE
constexpr uint32_t INFLATE_SMEM_WINDOW_BASE = 0x400;
#ifndef B2D_PROGRESS_SHIFT
#define B2D_PROGRESS_SHIFT 15
#endif
constexpr uint32_t INFLATE_PROGRESS_SHIFT = B2D_PROGRESS_SHIFT;          // 32 KiB pieces
cudaError_t probe_inflate_smem_base(uint32_t *base_out, cudaStream_t st);
cudaError_t launch_inflate(const uint8_t *d_in, const uint64_t *d_in_off, uint32_t n, uint8_t *d_out,
                           const uint64_t *d_out_off, uint64_t *d_out_len, uint64_t *d_in_consumed,
                           int *d_status, uint32_t flags, cudaStream_t st, uint8_t *out_mirror = nullptr,
                           const uint64_t *d_in_end = nullptr, uint32_t *progress = nullptr,
                           const uint8_t *in_host = nullptr, const uint32_t *landed = nullptr, uint32_t group_size = 0);
// in_host (optional): the compressed bytes are NOT in d_in yet but at this mapped host address (byte x of d_in <-> byte x
// of in_host, (in_host - d_in) % 128 == 0); the decoding warps pull them into d_in themselves as they go, until
// landed[i / group_size] (optional, device words) turns non-zero: the caller's own copy of member i's group has arrived
// progress (optional, instead of out_mirror): mapped host words, one per member; member i's word receives the number of
// whole (1 << INFLATE_PROGRESS_SHIFT)-byte pieces of its output that are final in d_out (0x7FFFFFFF when it is done)
// d_in_end (optional): member i occupies d_in[d_in_off[i], d_in_end[i]) instead of [d_in_off[i], d_in_off[i + 1])
// out_mirror (optional): a mapped host address for d_out[0] with (out_mirror - d_out) % 128 == 0; every output byte is
// then delivered there as well by the kernel itself (no device-to-host copy afterwards)

// block-parallel decode of a b2d_deflate_chunks stream (inflate.cu): chunk c occupies d_in[off[c], off[c+1]) and decodes
// to d_out[c * chunk_bytes ...); d_block_bits as produced by launch_deflate; d_chunk_status = 0 or the first failing
// block's status (such chunks are to be re-decoded serially for the exact outcome)
size_t inflate_units_scratch_bytes(uint64_t out_total, uint32_t chunk_bytes, uint32_t block_bytes);
cudaError_t launch_inflate_units(const uint8_t *d_in, const uint64_t *d_chunk_in_off, uint32_t n_chunks,
                                 const uint32_t *d_block_bits, uint32_t chunk_bytes, uint32_t block_bytes,
                                 uint64_t out_total, uint8_t *d_out, int *d_chunk_status, void *d_scratch, cudaStream_t st);

// speculative parallel decode of ONE raw-DEFLATE stream without an index (inflate.cu): d_in 4-byte aligned and readable
// up to 16 bytes past in_len; d_result = 5 x u64 {out_len, in_consumed, status (0 = done, else decode sequentially for
// the exact outcome), crc (0), n_live}
// cap_scale: multiplies the units' buffers (default 16 x the segment size: enough for ratios up to ~15 inside a unit)
size_t inflate_stream_scratch_bytes(uint64_t in_len, uint32_t cap_scale = 1);
cudaError_t launch_inflate_stream(const uint8_t *d_in, uint64_t in_len, uint8_t *d_out, uint64_t out_cap, void *d_result,
                                  void *d_scratch, cudaStream_t st, uint32_t cap_scale = 1);

// crc32.cu
// CRC-32 of n_seg independent segments: segment i = data[off[i], off[i] + len[i])  (len from d_len, u64)
cudaError_t launch_crc32_segments(const uint8_t *d_data, const uint64_t *d_off, const uint64_t *d_len,
                                  uint32_t n_seg, uint32_t *d_crc, cudaStream_t st);
// CRC-32 of fixed-size pieces of one buffer: piece i = data[i*piece, min((i+1)*piece, total))
cudaError_t launch_crc32_pieces(const uint8_t *d_data, uint64_t total, uint64_t piece, uint32_t n_pieces,
                                uint32_t *d_crc, cudaStream_t st);
// folds n_pieces piece CRCs (pieces of `piece` bytes, the last one shorter: total bytes) into one CRC on the device
cudaError_t launch_crc32_fold(const uint32_t *d_piece_crc, uint32_t n_pieces, uint64_t piece, uint64_t total,
                              uint32_t *d_crc_out, cudaStream_t st);
uint32_t host_crc32_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b);
// Adler-32 (same shapes as the CRC-32 launches); host_adler32_combine = adler32 of A||B from adler(A), adler(B), |B|
cudaError_t launch_adler32_segments(const uint8_t *d_data, const uint64_t *d_off, const uint64_t *d_len,
                                    uint32_t n_seg, uint32_t *d_out, cudaStream_t st);
cudaError_t launch_adler32_pieces(const uint8_t *d_data, uint64_t total, uint64_t piece, uint32_t n_pieces,
                                  uint32_t *d_out, cudaStream_t st);
uint32_t host_adler32_combine(uint32_t adler_a, uint32_t adler_b, uint64_t len_b);
uint32_t host_crc32_bytes(uint32_t crc, const uint8_t *p, size_t n);   // tiny inputs only (gzip headers)

// deflate.cu
struct DeflateParams {
	uint32_t chunk_bytes, block_bytes;
	int mode, search, depth, lazy, is_last;
	int checksum;     // B2D_CHECKSUM_*: what the per-chunk checksum array holds
	int framing;      // 0 = chunks closed by empty stored blocks; 1 = reference framing (one chunk, BFINAL on the last block)
	uint32_t leaf_bytes;   // 0 = no splitting; else the piece size of adaptive splitting (divides block_bytes)
};
uint64_t deflate_bound_bytes(uint64_t in_len, uint32_t chunk_bytes, uint32_t block_bytes);
size_t deflate_scratch_bytes(uint64_t in_len, const DeflateParams &p);
struct DeflateAux {            // a second stream and six events (timing disabled) the launch may use to overlap its stages
	cudaStream_t stream;
	cudaEvent_t ev[6];
};
cudaError_t launch_deflate(const uint8_t *d_in, uint64_t in_len, const DeflateParams &p, uint8_t *d_out,
                           uint64_t out_cap, uint64_t *d_out_len_total, uint64_t *d_chunk_out_len,
                           void *d_scratch, size_t scratch_bytes, cudaStream_t st, uint32_t *d_block_bits = nullptr,
                           const DeflateAux *aux = nullptr);
// d_block_bits (optional): bit offset of every block inside its chunk's output, one entry per block (the restart index
// of the block-parallel decoder); only meaningful with chunked framing.

}  // namespace b2d
