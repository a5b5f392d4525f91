// b2d_streams.hpp -- host-side mirror of the reference's stream API over the C ABI (include/b2deflate.h).
//
// The reference is Java (io.nayuki.deflate.*); this image has no JDK, so the host layer that sits above the C ABI is
// written in C++ with the SAME class names, constructor arguments, method meaning and error behaviour, so that the
// parity tests read like the reference's own.  (The Panama FFM Java twin is under ../java, uncompiled here.)
//   InflaterInputStream   <- InflaterInputStream.java:44,67,96-106 (ctors), :121-164 (read), :173-179 (close)
//   DeflaterOutputStream  <- DeflaterOutputStream.java:50-65 (ctors), :76-99 (write), :102-108 (finish), :111-116 (close)
//   GzipMetadata          <- GzipMetadata.java:28-66 (record + validation), :73-146 (read), :164-212 (write)
//   GzipInputStream       <- GzipInputStream.java:38-45 (ctor), :66-90 (read + trailer checks)
//   GzipOutputStream      <- GzipOutputStream.java:32-48 (ctors), :53-59 (write), :62-70 (finish)
//   DataFormatException   <- DataFormatException.java:15 (unchecked), :61-83 (Reason)
// What changes underneath: the codec engines (decomp/Open.java, comp/Lz77Huffman.java ...) are replaced by GPU batches.
// A decoder therefore buffers its whole input and decodes it in one call (or one call per chunk index), a
// compressor buffers `batch_bytes` of input per call.  There is no CPU codec here: without a B200 every stream
// operation throws IOException("No usable sm_100 GPU ...").
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <future>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/b2deflate.h"

namespace io_nayuki_deflate {

// ---------------------------------------------------------------- exceptions (Java names)
struct IOException : std::runtime_error { using std::runtime_error::runtime_error; };
struct IllegalStateException : std::logic_error { using std::logic_error::logic_error; };
struct IllegalArgumentException : std::invalid_argument { using std::invalid_argument::invalid_argument; };
struct IndexOutOfBoundsException : std::out_of_range { using std::out_of_range::out_of_range; };

// DataFormatException.java:15 -- unchecked there; a plain exception type here.  Reason ordinal = status - 1.
struct DataFormatException : std::runtime_error {
	enum class Reason {
		UNEXPECTED_END_OF_STREAM, RESERVED_BLOCK_TYPE, UNCOMPRESSED_BLOCK_LENGTH_MISMATCH, HUFFMAN_CODE_UNDER_FULL,
		HUFFMAN_CODE_OVER_FULL, NO_PREVIOUS_CODE_LENGTH_TO_COPY, CODE_LENGTH_CODE_OVER_FULL,
		END_OF_BLOCK_CODE_ZERO_LENGTH, RESERVED_LENGTH_SYMBOL, RESERVED_DISTANCE_SYMBOL,
		LENGTH_ENCOUNTERED_WITH_EMPTY_DISTANCE_CODE, COPY_FROM_BEFORE_DICTIONARY_START, HEADER_CHECKSUM_MISMATCH,
		UNSUPPORTED_COMPRESSION_METHOD, DECOMPRESSED_CHECKSUM_MISMATCH, DECOMPRESSED_SIZE_MISMATCH,
		GZIP_INVALID_MAGIC_NUMBER, GZIP_RESERVED_FLAGS_SET, GZIP_UNSUPPORTED_OPERATING_SYSTEM
	};
	DataFormatException(Reason r, const std::string &msg) : std::runtime_error(msg), reason(r) {}
	static DataFormatException fromStatus(int status) {
		return DataFormatException(static_cast<Reason>(status - 1), b2d_strerror(status));
	}
	static DataFormatException unexpectedEnd() {          // DataFormatException.java:44-48
		return DataFormatException(Reason::UNEXPECTED_END_OF_STREAM, "Unexpected end of stream");
	}
	Reason getReason() const { return reason; }
private:
	Reason reason;
};

inline void checkFromIndexSize(size_t off, size_t len, size_t size) {      // Objects.checkFromIndexSize
	if (off > size || len > size - off) throw IndexOutOfBoundsException("Range out of bounds");
}

// ---------------------------------------------------------------- java.io look-alikes (only what the path needs)
class InputStream {
public:
	virtual ~InputStream() {}
	virtual int read() { uint8_t b; return read(&b, 0, 1) == 1 ? b : -1; }
	virtual long read(uint8_t *b, size_t off, size_t len) = 0;       // bytes read, 0 iff len == 0, -1 at end
	virtual bool markSupported() const { return false; }
	virtual void mark() {}
	virtual void reset() { throw IOException("mark/reset not supported"); }
	virtual void close() {}
	void skipNBytes(uint64_t n) {
		uint8_t tmp[4096];
		while (n) {
			long r = read(tmp, 0, (size_t)std::min<uint64_t>(n, sizeof tmp));
			if (r <= 0) throw IOException("EOF while skipping");
			n -= (uint64_t)r;
		}
	}
	std::vector<uint8_t> readAllBytes() {
		std::vector<uint8_t> v;
		uint8_t tmp[1 << 16];
		for (long r; (r = read(tmp, 0, sizeof tmp)) > 0;) v.insert(v.end(), tmp, tmp + r);
		return v;
	}
};

class OutputStream {
public:
	virtual ~OutputStream() {}
	virtual void write(int b) { uint8_t x = (uint8_t)b; write(&x, 0, 1); }
	virtual void write(const uint8_t *b, size_t off, size_t len) = 0;
	virtual void flush() {}
	virtual void close() {}
};

class ByteArrayInputStream : public InputStream {
	const uint8_t *p; size_t n, pos = 0, marked = 0;
public:
	ByteArrayInputStream(const uint8_t *data, size_t len) : p(data), n(len) {}
	explicit ByteArrayInputStream(const std::vector<uint8_t> &v) : p(v.data()), n(v.size()) {}
	long read(uint8_t *b, size_t off, size_t len) override {
		if (len == 0) return 0;
		if (pos >= n) return -1;
		size_t k = std::min(len, n - pos);
		memcpy(b + off, p + pos, k);
		pos += k;
		return (long)k;
	}
	bool markSupported() const override { return true; }
	void mark() override { marked = pos; }
	void reset() override { pos = marked; }
	size_t position() const { return pos; }
};

class ByteArrayOutputStream : public OutputStream {
	std::vector<uint8_t> buf;
public:
	using OutputStream::write;
	void write(const uint8_t *b, size_t off, size_t len) override { buf.insert(buf.end(), b + off, b + off + len); }
	const std::vector<uint8_t> &toByteArray() const { return buf; }
};

// MarkableFileInputStream.java:30-70 -- a file stream whose mark/reset is a seek
class MarkableFileInputStream : public InputStream {
	FILE *f; long marked = 0;
public:
	explicit MarkableFileInputStream(const std::string &path) : f(fopen(path.c_str(), "rb")) {
		if (!f) throw IOException("Cannot open " + path);
	}
	~MarkableFileInputStream() override { if (f) fclose(f); }
	long read(uint8_t *b, size_t off, size_t len) override {
		if (len == 0) return 0;
		size_t k = fread(b + off, 1, len, f);
		if (k == 0) { if (ferror(f)) throw IOException("read error"); return -1; }
		return (long)k;
	}
	bool markSupported() const override { return true; }
	void mark() override { marked = ftell(f); }
	void reset() override { if (fseek(f, marked, SEEK_SET)) throw IOException("seek error"); }
	void close() override { if (f) { fclose(f); f = nullptr; } }
};

class FileOutputStream : public OutputStream {
	FILE *f;
public:
	using OutputStream::write;
	explicit FileOutputStream(const std::string &path) : f(fopen(path.c_str(), "wb")) {
		if (!f) throw IOException("Cannot open " + path);
	}
	~FileOutputStream() override { if (f) fclose(f); }
	void write(const uint8_t *b, size_t off, size_t len) override {
		if (len && fwrite(b + off, 1, len, f) != len) throw IOException("write error");
	}
	void flush() override { if (f) fflush(f); }
	void close() override { if (f) { if (fclose(f)) { f = nullptr; throw IOException("close error"); } f = nullptr; } }
};

// ---------------------------------------------------------------- device binding
// B2D_DEVICE=<ordinal> (default 0) binds one GPU; B2D_DEVICES=all or B2D_DEVICES=0,2,3 binds several, and the host entry
// points then shard chunks / members over them (b2d_init_devices).
inline int bindDevices() {
	if (const char *ds = getenv("B2D_DEVICES")) {
		if (!strcmp(ds, "all")) return b2d_init_devices(nullptr, 0);
		std::vector<int> v;
		for (const char *p = ds; *p;) {
			char *e;
			long d = strtol(p, &e, 10);
			if (e == p) break;
			v.push_back((int)d);
			p = *e == ',' ? e + 1 : e;
		}
		if (!v.empty()) return b2d_init_devices(v.data(), (int)v.size());
	}
	return b2d_init(getenv("B2D_DEVICE") ? atoi(getenv("B2D_DEVICE")) : 0);
}
inline void requireDevice() {
	static int rc = bindDevices();
	if (rc != B2D_OK) throw IOException(std::string(b2d_strerror(rc)) + " [" + b2d_last_error() + "]");
}

struct PinnedBuffer {                       // staging memory from b2d_alloc_pinned (falls back to nothing: throws)
	uint8_t *p = nullptr; size_t cap = 0;
	PinnedBuffer() {}
	PinnedBuffer(const PinnedBuffer &) = delete;
	~PinnedBuffer() { if (p) b2d_free_pinned(p); }
	void reserve(size_t n, size_t keep = 0) {
		if (n <= cap) return;
		size_t want = std::max(n, cap + cap / 2);
		uint8_t *q = (uint8_t *)b2d_alloc_pinned(want);
		if (!q) throw IOException("b2d_alloc_pinned failed");
		if (keep) memcpy(q, p, keep);
		if (p) b2d_free_pinned(p);
		p = q; cap = want;
	}
};

// Chunk index of a stream made by DeflaterOutputStream (sizes of the compressed chunks, each closed by an empty
// stored block).  With it a decoder hands every chunk to its own warp; without it a stream is one serial unit.
struct ChunkIndex {
	uint32_t chunk_bytes = 0;                // uncompressed bytes per chunk (the last may be shorter)
	std::vector<uint64_t> sizes;             // compressed bytes per chunk
	uint32_t block_bytes = 0;                // optional restart index: bit offset of every block_bytes-block inside its chunk
	std::vector<uint32_t> block_bits;
	bool empty() const { return sizes.empty(); }
};

// ---------------------------------------------------------------- InflaterInputStream
class InflaterInputStream : public InputStream {
public:
	static constexpr int DEFAULT_INPUT_BUFFER_SIZE = 16 * 1024;          // InflaterInputStream.java:72

	explicit InflaterInputStream(InputStream &in) : InflaterInputStream(in, false) {}
	InflaterInputStream(InputStream &in, bool endExactly) : InflaterInputStream(in, endExactly, DEFAULT_INPUT_BUFFER_SIZE) {}
	InflaterInputStream(InputStream &in, bool endExactly, int inBufLen) : input(&in), endExactly(endExactly) {
		if (inBufLen <= 0) throw IllegalArgumentException("Non-positive input buffer size");          // :98-99
		if (endExactly) {
			if (!in.markSupported()) throw IllegalArgumentException("Input stream not markable, cannot support endExactly");   // :100-103
			in.mark();
		}
	}
	// Extensions of the GPU build: a chunk index (parallel decode) and a size hint (e.g. gzip ISIZE).
	void setChunkIndex(ChunkIndex idx) { index = std::move(idx); }
	void setOutputSizeHint(uint64_t n) { sizeHint = n; }
	void setTrailerBytes(size_t n) { trailerBytes = n; }     // bytes after the DEFLATE data that belong to the caller
	void setChecksumAdler32(bool on) { adler = on; }         // checksum() is Adler-32 (zlib container) instead of CRC-32
	uint32_t crc32() { decodeEverything(); return crc; }     // checksum of everything decoded (computed on the GPU)
	uint32_t checksum() { return crc32(); }
	uint64_t consumedBytes() { decodeEverything(); return consumed; }

	int read() override {                                                                // :121-134
		uint8_t b;
		long r = read(&b, 0, 1);
		return r == 1 ? b : -1;
	}
	long read(uint8_t *b, size_t off, size_t len) override {                             // :147-164
		if (closed) throw IllegalStateException("Stream already closed");                 // :160-161
		if (sticky) throw IOException(*sticky);                                           // StickyException.java:17-26
		if (len == 0) return 0;
		try {
			decodeAll();
			while (pos >= out.size() && moreBatches()) adoptNextBatch(false);
		}
		catch (const IOException &e) { sticky = e.what(); throw; }                        // I/O errors are sticky (:152-157)
		if (pos < out.size()) {
			size_t k = std::min(len, out.size() - pos);
			memcpy(b + off, out.data() + pos, k);
			pos += k;
			return (long)k;
		}
		if (status != 0) throw DataFormatException::fromStatus(status);    // format errors are not sticky: thrown again on retry
		return -1;
	}
	void close() override {                                                              // :173-179, idempotent
		if (closed) return;
		closed = true;
		if (pending.valid()) pending.wait();            // (a batch on its way still reads the underlying stream)
		input->close();
	}

private:
	InputStream *input;
	bool endExactly, closed = false, decoded = false;
	std::optional<std::string> sticky;
	ChunkIndex index;
	bool adler = false;
	uint64_t sizeHint = 0, consumed = 0;
	size_t trailerBytes = 0;
	std::vector<uint8_t> out;
	size_t pos = 0;
	int status = 0;
	uint32_t crc = 0;

	// A stream with a block index is decoded in batches of batchChunks() chunks: while the caller consumes (writes out)
	// batch k, a second thread already reads the compressed bytes of batch k + 1 (the index says how many there are) and
	// decodes them on the GPU -- the read side of the overlapped file pipeline (the write side is bin/gzip's).
	static uint32_t batchChunks() {                      // B2D_GUNZIP_BATCH=<chunks per batch> (default 256)
		static const uint32_t n = getenv("B2D_GUNZIP_BATCH") ? std::max<uint32_t>(1, (uint32_t)strtoul(getenv("B2D_GUNZIP_BATCH"), nullptr, 10)) : 256u;
		return n;
	}
	struct Batch { std::vector<uint8_t> bytes; std::vector<uint32_t> crcs; bool ok = false, truncated = false; std::string error; };
	std::unique_ptr<PinnedBuffer> batchIn;               // the compressed input read so far, kept while batches are pending
	std::vector<uint64_t> batchInOff;                    // chunk offsets in it
	uint64_t batchInFilled = 0;                          // bytes of it that have been read from the underlying stream
	uint32_t nextChunk = 0;                              // first chunk that is neither delivered nor being decoded
	uint32_t pendingFirst = 0, pendingCount = 0;
	std::future<Batch> pending;
	size_t usableIn = 0;

	bool moreBatches() const { return pending.valid(); }
	// reads the underlying stream until the first `upTo` compressed bytes are in batchIn; false = the stream ended before
	bool fillTo(uint64_t upTo) {
		while (batchInFilled < upTo) {
			long r = input->read(batchIn->p, (size_t)batchInFilled, (size_t)std::min<uint64_t>(upTo - batchInFilled, 8u << 20));
			if (r <= 0) return false;
			batchInFilled += (uint64_t)r;
		}
		return true;
	}
	Batch decodeBatch(uint32_t c0, uint32_t nb) {
		Batch r;
		const uint32_t n = (uint32_t)index.sizes.size();
		const uint64_t bpc = index.chunk_bytes / index.block_bytes;
		try {
			if (!fillTo(batchInOff[c0 + nb])) { r.truncated = true; return r; }
		} catch (const std::exception &e) { r.error = e.what(); return r; }
		const uint64_t o0 = (uint64_t)c0 * index.chunk_bytes, o1 = std::min<uint64_t>(sizeHint, (uint64_t)(c0 + nb) * index.chunk_bytes);
		PinnedBuffer pout;
		pout.reserve(o1 - o0 + 64);
		r.crcs.resize(nb);
		std::vector<int32_t> st(nb);
		int rc = b2d_inflate_chunks(batchIn->p + batchInOff[c0], index.sizes.data() + c0, nb, index.block_bits.data() + (size_t)c0 * bpc,
		                            index.chunk_bytes, index.block_bytes, pout.p, o1 - o0, r.crcs.data(), st.data(),
		                            adler ? B2D_INFLATE_ADLER32 : B2D_INFLATE_CRC32);
		if (rc != B2D_OK) { r.error = std::string("b2d_inflate_chunks: ") + b2d_strerror(rc) + " [" + b2d_last_error() + "]"; return r; }
		r.ok = true;
		for (uint32_t i = 0; i < nb; i++) if (st[i] != 0) r.ok = false;     // the chunk-indexed path reports the exact outcome
		if (r.ok) r.bytes.assign(pout.p, pout.p + (o1 - o0));
		(void)n;
		return r;
	}
	void launchBatch() {
		const uint32_t n = (uint32_t)index.sizes.size();
		if (nextChunk >= n) return;
		pendingFirst = nextChunk;
		pendingCount = std::min<uint32_t>(batchChunks(), n - nextChunk);
		nextChunk += pendingCount;
		pending = std::async(std::launch::async, [this, c0 = pendingFirst, nb = pendingCount] { return decodeBatch(c0, nb); });
	}
	// takes over the batch that is being decoded; append = keep what is still unread of the current one in front of it
	void adoptNextBatch(bool append) {
		Batch b = pending.get();
		if (!b.error.empty()) throw IOException(b.error);
		const uint32_t c0 = pendingFirst, nb = pendingCount;
		if (b.truncated) {                              // the stream ends inside this batch: what was delivered stays, then the exception
			if (!append) { out.clear(); pos = 0; }
			status = B2D_UNEXPECTED_END_OF_STREAM;
			consumed = batchInFilled;
			batchIn.reset();
			return;
		}
		if (!b.ok) {                                    // something unusual in this batch: the rest goes the careful way, chunk by chunk
			std::vector<uint8_t> keep;
			if (append) keep.assign(out.begin() + (long)pos, out.end());
			PinnedBuffer pout;
			usableIn = fillTo(batchInOff.back()) ? (size_t)batchInOff.back() : (size_t)batchInFilled;
			decodeIndexedFrom(*batchIn, pout, usableIn, c0, keep);
			pos = 0;
			batchIn.reset();
			if (endExactly && status == 0) {                // Open.finish, Open.java:113-124
				input->reset();
				input->skipNBytes(consumed);
			}
			return;
		}
		for (uint32_t i = 0; i < nb; i++) {
			const uint64_t len = std::min<uint64_t>(index.chunk_bytes, sizeHint - (uint64_t)(c0 + i) * index.chunk_bytes);
			crc = adler ? b2d_adler32_combine(crc, b.crcs[i], len) : b2d_crc32_combine(crc, b.crcs[i], len);
		}
		if (append) { out.erase(out.begin(), out.begin() + (long)pos); out.insert(out.end(), b.bytes.begin(), b.bytes.end()); }
		else out = std::move(b.bytes);
		pos = 0;
		consumed = batchInOff[c0 + nb];
		launchBatch();
		if (!pending.valid()) {                          // that was the last batch
			status = 0;
			batchIn.reset();
			if (endExactly) {                            // Open.finish, Open.java:113-124
				input->reset();
				input->skipNBytes(consumed);
			}
		}
	}
	void decodeEverything() {
		decodeAll();
		while (moreBatches()) adoptNextBatch(true);
	}
	// the batched path applies when the index has block offsets and the exact size is known (what bin/gzip writes);
	// it reads the underlying stream itself, batch by batch
	bool startBatches() {
		const uint32_t n = (uint32_t)index.sizes.size();
		if (n <= batchChunks() || index.block_bytes == 0 || sizeHint == 0 || index.chunk_bytes % index.block_bytes != 0) return false;
		const uint64_t bpc = index.chunk_bytes / index.block_bytes;
		if (index.block_bits.size() != (size_t)n * bpc) return false;
		if ((uint64_t)n * index.chunk_bytes < sizeHint || (uint64_t)(n - 1) * index.chunk_bytes >= sizeHint) return false;
		batchInOff.assign(n + 1, 0);
		for (uint32_t i = 0; i < n; i++) batchInOff[i + 1] = batchInOff[i] + index.sizes[i];
		batchIn.reset(new PinnedBuffer());
		batchIn->reserve(batchInOff[n] + 64);
		batchInFilled = 0;
		usableIn = (size_t)batchInOff[n];
		crc = adler ? 1 : 0;
		nextChunk = 0;
		launchBatch();
		adoptNextBatch(false);
		return true;
	}

	void decodeAll() {
		if (decoded) return;
		decoded = true;
		requireDevice();
		if (!index.empty() && startBatches()) return;     // batches deliver as they come
		std::vector<uint8_t> raw = input->readAllBytes();
		const size_t usable = raw.size() >= trailerBytes ? raw.size() - trailerBytes : raw.size();
		std::unique_ptr<PinnedBuffer> pin(new PinnedBuffer());
		PinnedBuffer pout;
		pin->reserve(raw.size() + 64);
		if (!raw.empty()) memcpy(pin->p, raw.data(), raw.size());
		if (!index.empty()) decodeIndexed(*pin, pout, usable);
		else decodeSerial(*pin, pout, usable);
		if (endExactly && status == 0 && !moreBatches()) {               // Open.finish, Open.java:113-124
			input->reset();
			input->skipNBytes(consumed);
		}
	}

	void decodeSerial(PinnedBuffer &pin, PinnedBuffer &pout, size_t in_len) {
		uint64_t cap = sizeHint ? sizeHint : std::max<uint64_t>(1 << 16, (uint64_t)in_len * 6);
		for (;;) {
			pout.reserve(cap + 64);
			uint64_t out_len = 0, cons = 0;
			int32_t st = 0;
			// a stream nobody indexed: speculative parallel decode, sequential decoder behind it for anything unusual
			int rc = b2d_inflate_stream(pin.p, in_len, pout.p, cap, &out_len, &cons, &crc, &st,
			                            adler ? B2D_INFLATE_ADLER32 : B2D_INFLATE_CRC32, nullptr);
			if (rc != B2D_OK) throw IOException(std::string("b2d_inflate_stream: ") + b2d_strerror(rc) + " [" + b2d_last_error() + "]");
			if (st == B2D_ERR_OUTPUT_OVERFLOW) {                         // the size is unknown up front: retry larger
				if (cap > ((uint64_t)1 << 40)) throw IOException("decompressed size exceeds 1 TiB");
				cap = cap * 2 + (1 << 20);
				continue;
			}
			status = st;
			consumed = cons;
			out.assign(pout.p, pout.p + out_len);
			return;
		}
	}

	// chunk sizes + block offsets + exact size known: every block gets its own warp (b2d_inflate_chunks)
	bool decodeBlocks(PinnedBuffer &pin, PinnedBuffer &pout, size_t in_len) {
		const uint32_t n = (uint32_t)index.sizes.size();
		if (index.block_bytes == 0 || sizeHint == 0 || index.chunk_bytes % index.block_bytes != 0) return false;
		const uint64_t bpc = index.chunk_bytes / index.block_bytes;
		if (index.block_bits.size() != (size_t)n * bpc) return false;
		if ((uint64_t)n * index.chunk_bytes < sizeHint || (uint64_t)(n - 1) * index.chunk_bytes >= sizeHint) return false;
		uint64_t total_in = 0;
		for (uint64_t sz : index.sizes) total_in += sz;
		if (total_in > in_len) return false;
		pout.reserve(sizeHint + 64);
		std::vector<uint32_t> crcs(n);
		std::vector<int32_t> st(n);
		int rc = b2d_inflate_chunks(pin.p, index.sizes.data(), n, index.block_bits.data(), index.chunk_bytes, index.block_bytes,
		                            pout.p, sizeHint, crcs.data(), st.data(), adler ? B2D_INFLATE_ADLER32 : B2D_INFLATE_CRC32);
		if (rc != B2D_OK) throw IOException(std::string("b2d_inflate_chunks: ") + b2d_strerror(rc) + " [" + b2d_last_error() + "]");
		for (uint32_t i = 0; i < n; i++) if (st[i] != 0) return false;      // let the chunk-indexed path report the exact outcome
		out.assign(pout.p, pout.p + sizeHint);
		crc = adler ? 1 : 0;
		for (uint32_t i = 0; i < n; i++) {
			const uint64_t len = std::min<uint64_t>(index.chunk_bytes, sizeHint - (uint64_t)i * index.chunk_bytes);
			crc = adler ? b2d_adler32_combine(crc, crcs[i], len) : b2d_crc32_combine(crc, crcs[i], len);
		}
		consumed = total_in;
		status = 0;
		return true;
	}

	void decodeIndexed(PinnedBuffer &pin, PinnedBuffer &pout, size_t in_len) {
		if (decodeBlocks(pin, pout, in_len)) return;
		crc = adler ? 1 : 0;
		decodeIndexedFrom(pin, pout, in_len, 0, std::vector<uint8_t>());
	}
	// chunks c0 .. end, one warp per chunk (exact status and delivered bytes); `front` = bytes to deliver before them;
	// crc holds the checksum of everything in front of chunk c0
	void decodeIndexedFrom(PinnedBuffer &pin, PinnedBuffer &pout, size_t in_len, uint32_t c0, const std::vector<uint8_t> &front) {
		const uint32_t n_all = (uint32_t)index.sizes.size(), n = n_all - c0;
		uint64_t base_in = 0;
		for (uint32_t i = 0; i < c0; i++) base_in += index.sizes[i];
		std::vector<uint64_t> in_off(n + 1, base_in), out_off(n + 1, 0), out_len(n), cons(n);
		std::vector<uint32_t> crcs(n);
		std::vector<int32_t> st(n);
		for (uint32_t i = 0; i < n; i++) {
			in_off[i + 1] = in_off[i] + index.sizes[c0 + i];
			out_off[i + 1] = out_off[i] + index.chunk_bytes;
		}
		if (in_off[n] > in_len) throw DataFormatException::unexpectedEnd();
		pout.reserve(out_off[n] + 64);
		int rc = b2d_inflate_batch(pin.p, in_off.data(), n, pout.p, out_off.data(), out_len.data(), cons.data(), crcs.data(),
		                           st.data(), (adler ? B2D_INFLATE_ADLER32 : B2D_INFLATE_CRC32) | B2D_INFLATE_CHUNK_INDEXED);
		if (rc != B2D_OK) throw IOException(std::string("b2d_inflate_batch: ") + b2d_strerror(rc) + " [" + b2d_last_error() + "]");
		out = front;
		out.reserve(front.size() + out_off[n]);
		for (uint32_t i = 0; i < n; i++) {                               // deliver up to the first failing chunk, as a serial decode would
			out.insert(out.end(), pout.p + out_off[i], pout.p + out_off[i] + out_len[i]);
			crc = adler ? b2d_adler32_combine(crc, crcs[i], out_len[i]) : b2d_crc32_combine(crc, crcs[i], out_len[i]);
			consumed = in_off[i] + cons[i];
			if (st[i] != 0 || cons[i] != index.sizes[c0 + i] || (i + 1 < n && out_len[i] != index.chunk_bytes)) {
				status = st[i] != 0 ? st[i] : B2D_UNEXPECTED_END_OF_STREAM;
				return;
			}
		}
	}
};

// ---------------------------------------------------------------- DeflaterOutputStream
struct DeflaterOptions {                     // the GPU build's counterpart of (dataLookaheadLimit, historyLookbehindLimit, Strategy)
	uint32_t chunk_bytes = 1u << 20;         // independent unit, history reset (the reference carries 32 KiB across blocks)
	uint32_t block_bytes = 1u << 16;         // dataLookaheadLimit: one DEFLATE block per this many bytes (DeflaterOutputStream.java:50)
	int mode = B2D_MODE_AUTO;
	int search = B2D_SEARCH_DEFAULT;
	int chain_depth = 0;
	int lazy = -1;
	uint64_t batch_bytes = 256ull << 20;     // input buffered per GPU call
	int checksum = B2D_CHECKSUM_CRC32;       // which checksum of the input rides the GPU call (gzip: CRC-32, zlib: Adler-32)
	uint32_t split_min_bytes = 0;            // != 0: `new BinarySplit(strategy, n)`-style adaptive blocks (comp/BinarySplit.java)
};

class DeflaterOutputStream : public OutputStream {
public:
	explicit DeflaterOutputStream(OutputStream &out) : DeflaterOutputStream(out, DeflaterOptions()) {}
	DeflaterOutputStream(OutputStream &out, const DeflaterOptions &o)
	    : output(&out), opt(o), crc(o.checksum == B2D_CHECKSUM_ADLER32 ? 1u : 0u) {
		if (o.block_bytes < 4096 || o.chunk_bytes % o.block_bytes != 0 || o.batch_bytes < o.chunk_bytes ||
		    o.batch_bytes % o.chunk_bytes != 0)
			throw IllegalArgumentException("Invalid capacities");                          // :58-60
	}
	OutputStream &getUnderlyingStream() {                                                 // :69-73
		if (ended) throw IllegalStateException("Stream already ended");
		return *output;
	}
	using OutputStream::write;
	void write(int b) override { uint8_t x = (uint8_t)b; write(&x, 0, 1); }               // :76-83
	void write(const uint8_t *b, size_t off, size_t len) override {                       // :86-99
		if (ended) throw IllegalStateException("Stream already ended");
		requireDevice();
		while (len) {
			if (fill == opt.batch_bytes) flushBatch(false);
			stage.reserve(std::min<uint64_t>(opt.batch_bytes, std::max<uint64_t>(fill + len, 1 << 20)), fill);
			size_t k = (size_t)std::min<uint64_t>(len, std::min<uint64_t>(opt.batch_bytes, stage.cap) - fill);
			memcpy(stage.p + fill, b + off, k);
			fill += k; off += k; len -= k;
		}
	}
	void finish() {                                                                       // :102-108
		if (ended) throw IllegalStateException("Stream already ended");
		flushBatch(true);
		ended = true;
	}
	void close() override {                                                               // :111-116
		if (!ended) finish();
		output->close();
	}
	uint32_t crc32() const { return crc; }               // checksum (opt.checksum) of all bytes compressed so far
	uint32_t checksum() const { return crc; }
	uint64_t totalIn() const { return total_in; }
	const ChunkIndex &chunkIndex() const { return index; }

private:
	OutputStream *output;
	DeflaterOptions opt;
	PinnedBuffer stage, comp;
	uint64_t fill = 0, total_in = 0;
	uint32_t crc;
	bool ended = false;
	ChunkIndex index;

	void flushBatch(bool last) {
		requireDevice();
		if (fill == 0 && !last) return;
		b2d_deflate_opts o;
		memset(&o, 0, sizeof o);
		o.chunk_bytes = opt.chunk_bytes; o.block_bytes = opt.block_bytes; o.mode = opt.mode; o.search = opt.search;
		o.chain_depth = opt.chain_depth; o.lazy = opt.lazy; o.is_last = last ? 1 : 0; o.framing = B2D_FRAMING_CHUNKED;
		o.checksum = opt.checksum;
		o.split_min_bytes = opt.split_min_bytes;
		const uint64_t bound = b2d_deflate_bound(fill, opt.chunk_bytes);
		comp.reserve(bound);
		stage.reserve(1);
		const size_t n_chunks = (size_t)((fill + opt.chunk_bytes - 1) / opt.chunk_bytes);
		std::vector<uint64_t> sizes(std::max<size_t>(n_chunks, 1));
		const size_t n_blocks = (size_t)((fill + opt.block_bytes - 1) / opt.block_bytes);
		std::vector<uint32_t> bits(std::max<size_t>(n_blocks, 1));
		int64_t n = b2d_deflate_chunks_indexed(stage.p, fill, &o, comp.p, bound, &crc, sizes.data(), bits.data());
		if (n < 0) throw IOException(std::string("b2d_deflate_chunks: ") + b2d_strerror((int)n) + " [" + b2d_last_error() + "]");
		index.chunk_bytes = opt.chunk_bytes;
		index.block_bytes = opt.block_bytes;
		// a batch is a whole number of chunks (except the last), so block entries stay chunk-aligned across batches
		const size_t bpc = opt.chunk_bytes / opt.block_bytes;
		for (size_t c = 0; c < n_chunks; c++)
			for (size_t b = 0; b < bpc; b++) index.block_bits.push_back(c * bpc + b < n_blocks ? bits[c * bpc + b] : 0u);
		if (n_chunks) index.sizes.insert(index.sizes.end(), sizes.begin(), sizes.begin() + n_chunks);
		else if (n > 0) {                                 // empty final call: the 5-byte closing block joins the previous chunk
			if (index.sizes.empty()) index.sizes.push_back((uint64_t)n); else index.sizes.back() += (uint64_t)n;
		}
		output->write(comp.p, 0, (size_t)n);
		total_in += fill;
		fill = 0;
	}
};

// ---------------------------------------------------------------- GzipMetadata
struct GzipMetadata {
	enum class CompressionMethod { DEFLATE };
	enum class OperatingSystem {
		FAT_FILESYSTEM, AMIGA, VMS, UNIX, VM_CMS, ATARI_TOS, HPFS_FILESYSTEM, MACINTOSH, Z_SYSTEM, CPM, TOPS_20,
		NTFS_FILESYSTEM, QDOS, ACORN_RISCOS, UNKNOWN
	};
	CompressionMethod compressionMethod = CompressionMethod::DEFLATE;
	bool isFileText = false;
	std::optional<int32_t> modificationTimeUnixS;
	int extraFlags = 0;
	OperatingSystem operatingSystem = OperatingSystem::UNIX;
	std::optional<std::vector<uint8_t>> extraField;
	std::optional<std::string> fileName;
	std::optional<std::string> comment;
	bool hasHeaderCrc = false;

	GzipMetadata() {}
	GzipMetadata(CompressionMethod cm, bool text, std::optional<int32_t> mtime, int xfl, OperatingSystem os,
	             std::optional<std::vector<uint8_t>> extra, std::optional<std::string> name,
	             std::optional<std::string> comment_, bool hcrc)
	    : compressionMethod(cm), isFileText(text), modificationTimeUnixS(mtime), extraFlags(xfl), operatingSystem(os),
	      extraField(std::move(extra)), fileName(std::move(name)), comment(std::move(comment_)), hasHeaderCrc(hcrc) {
		validate();
	}
	void validate() const {                                                               // GzipMetadata.java:41-65
		if (modificationTimeUnixS && *modificationTimeUnixS == 0) throw IllegalArgumentException("Modification timestamp is zero");
		if ((unsigned)extraFlags >> 8 != 0) throw IllegalArgumentException("Invalid extra flags value");
		if (extraField && extraField->size() > 0xFFFF) throw IllegalArgumentException("Extra field too long");
	}

	// The header CRC-16 is the low half of the CRC-32 of the header bytes (GzipMetadata.java:134,211).  Headers are
	// tens of bytes; this is bookkeeping, not the hot path, so it is a plain bitwise loop on the host.
	static uint32_t headerCrc(const std::vector<uint8_t> &bytes) {
		uint32_t c = 0xFFFFFFFFu;
		for (uint8_t b : bytes) {
			c ^= b;
			for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
		}
		return ~c;
	}

	static GzipMetadata read(InputStream &in) {                                            // GzipMetadata.java:73-146
		std::vector<uint8_t> seen;
		auto u8 = [&]() -> int {
			int b = in.read();
			if (b < 0) throw DataFormatException::unexpectedEnd();                          // :143-145
			seen.push_back((uint8_t)b);
			return b;
		};
		if (((u8() << 8) | u8()) != 0x1F8B)
			throw DataFormatException(DataFormatException::Reason::GZIP_INVALID_MAGIC_NUMBER, "Invalid GZIP magic number");
		int cm = u8();
		if (cm != 8)
			throw DataFormatException(DataFormatException::Reason::UNSUPPORTED_COMPRESSION_METHOD,
			                          "Unsupported compression method: " + std::to_string(cm));
		int flags = u8();
		if (flags & 0xE0) throw DataFormatException(DataFormatException::Reason::GZIP_RESERVED_FLAGS_SET, "Reserved flags are set");
		GzipMetadata m;
		uint32_t mt = 0;
		for (int i = 0; i < 4; i++) mt |= (uint32_t)u8() << (8 * i);
		if (mt != 0) m.modificationTimeUnixS = (int32_t)mt;
		m.extraFlags = u8();
		int os = u8();
		if (os < (int)OperatingSystem::UNKNOWN) m.operatingSystem = (OperatingSystem)os;
		else if (os == 0xFF) m.operatingSystem = OperatingSystem::UNKNOWN;
		else throw DataFormatException(DataFormatException::Reason::GZIP_UNSUPPORTED_OPERATING_SYSTEM, "Unsupported operating system value");
		m.isFileText = flags & 1;
		if (flags & 4) {
			int len = u8(); len |= u8() << 8;
			std::vector<uint8_t> x(len);
			for (int i = 0; i < len; i++) x[i] = (uint8_t)u8();
			m.extraField = std::move(x);
		}
		auto zstr = [&]() { std::string s; for (int b; (b = u8()) != 0;) s.push_back((char)b); return s; };
		if (flags & 8) m.fileName = zstr();
		if (flags & 16) m.comment = zstr();
		m.hasHeaderCrc = flags & 2;
		if (m.hasHeaderCrc) {
			int expect = (int)(headerCrc(seen) & 0xFFFF);
			int actual = u8(); actual |= u8() << 8;
			if (actual != expect) throw DataFormatException(DataFormatException::Reason::HEADER_CHECKSUM_MISMATCH, "Header CRC-16 mismatch");
		}
		return m;
	}

	void write(OutputStream &out) const {                                                  // GzipMetadata.java:164-212
		std::vector<uint8_t> h = {0x1F, 0x8B, 8};
		h.push_back((uint8_t)((isFileText ? 1 : 0) | (hasHeaderCrc ? 2 : 0) | (extraField ? 4 : 0) | (fileName ? 8 : 0) | (comment ? 16 : 0)));
		uint32_t mt = modificationTimeUnixS ? (uint32_t)*modificationTimeUnixS : 0;
		for (int i = 0; i < 4; i++) h.push_back((uint8_t)(mt >> (8 * i)));
		h.push_back((uint8_t)extraFlags);
		h.push_back(operatingSystem == OperatingSystem::UNKNOWN ? 0xFF : (uint8_t)operatingSystem);
		if (extraField) {
			h.push_back((uint8_t)(extraField->size() & 0xFF));
			h.push_back((uint8_t)(extraField->size() >> 8));
			h.insert(h.end(), extraField->begin(), extraField->end());
		}
		if (fileName) { h.insert(h.end(), fileName->begin(), fileName->end()); h.push_back(0); }
		if (comment) { h.insert(h.end(), comment->begin(), comment->end()); h.push_back(0); }
		if (hasHeaderCrc) {
			uint32_t c = headerCrc(h);
			h.push_back((uint8_t)(c & 0xFF));
			h.push_back((uint8_t)((c >> 8) & 0xFF));
		}
		out.write(h.data(), 0, h.size());
	}

	// FEXTRA subfield "B2" (RFC 1952 2.3.1.1): u32 chunk_bytes, then one u32 compressed size per chunk.  The reference
	// parses and ignores extra fields (GzipMetadata.java:116-122), so files carrying it stay readable by it.
	// A second subfield "B3" (u32 block_bytes, one u32 bit offset per block) follows when both still fit 64 KiB: with it
	// gunzip decodes every block on its own warp.
	static std::optional<std::vector<uint8_t>> encodeChunkIndex(const ChunkIndex &idx) {
		size_t bytes = 4 + 4 + 4 * idx.sizes.size();
		if (idx.empty() || bytes > 0xFFFF) return std::nullopt;
		for (uint64_t s : idx.sizes) if (s > 0xFFFFFFFFull) return std::nullopt;
		std::vector<uint8_t> x = {'B', '2', (uint8_t)((bytes - 4) & 0xFF), (uint8_t)((bytes - 4) >> 8)};
		auto put32 = [&](uint32_t v) { for (int i = 0; i < 4; i++) x.push_back((uint8_t)(v >> (8 * i))); };
		put32(idx.chunk_bytes);
		for (uint64_t s : idx.sizes) put32((uint32_t)s);
		const size_t b3 = 4 + 4 + 4 * idx.block_bits.size();
		if (idx.block_bytes && !idx.block_bits.empty() && b3 - 4 <= 0xFFFF && x.size() + b3 <= 0xFFFF) {
			x.push_back('B'); x.push_back('3'); x.push_back((uint8_t)((b3 - 4) & 0xFF)); x.push_back((uint8_t)((b3 - 4) >> 8));
			put32(idx.block_bytes);
			for (uint32_t v : idx.block_bits) put32(v);
		}
		return x;
	}
	ChunkIndex chunkIndex() const {
		ChunkIndex idx;
		if (!extraField) return idx;
		const std::vector<uint8_t> &x = *extraField;
		for (size_t p = 0; p + 4 <= x.size();) {
			size_t len = x[p + 2] | (size_t)x[p + 3] << 8;
			if (p + 4 + len > x.size()) break;
			auto get32 = [&](size_t q) { return (uint32_t)x[q] | (uint32_t)x[q + 1] << 8 | (uint32_t)x[q + 2] << 16 | (uint32_t)x[q + 3] << 24; };
			if (x[p] == 'B' && x[p + 1] == '2' && len >= 4 && len % 4 == 0 && idx.sizes.empty()) {
				idx.chunk_bytes = get32(p + 4);
				for (size_t q = p + 8; q < p + 4 + len; q += 4) idx.sizes.push_back(get32(q));
				if (idx.chunk_bytes == 0) idx.sizes.clear();
			} else if (x[p] == 'B' && x[p + 1] == '3' && len >= 4 && len % 4 == 0) {
				idx.block_bytes = get32(p + 4);
				for (size_t q = p + 8; q < p + 4 + len; q += 4) idx.block_bits.push_back(get32(q));
			}
			p += 4 + len;
		}
		if (idx.sizes.empty()) { idx.block_bytes = 0; idx.block_bits.clear(); }
		return idx;
	}
};

// ---------------------------------------------------------------- GzipOutputStream
class GzipOutputStream : public OutputStream {
public:
	GzipOutputStream(OutputStream &out, const GzipMetadata &meta) : GzipOutputStream(out, meta, DeflaterOptions()) {}
	GzipOutputStream(OutputStream &out, const GzipMetadata &meta, const DeflaterOptions &o) : under(&out), deflater(out, o) {
		meta.write(out);                                                                   // GzipOutputStream.java:40
	}
	using OutputStream::write;
	void write(int b) override { uint8_t x = (uint8_t)b; write(&x, 0, 1); }
	void write(const uint8_t *b, size_t off, size_t len) override {                       // :53-59 (CRC and length ride the GPU call)
		if (ended) throw IllegalStateException("Stream already ended");
		deflater.write(b, off, len);
	}
	void finish() {                                                                       // :62-70
		if (ended) throw IllegalStateException("Stream already ended");
		deflater.finish();
		uint8_t t[8];
		uint32_t c = deflater.crc32(), n = (uint32_t)deflater.totalIn();                  // ISIZE is mod 2^32 (:69)
		for (int i = 0; i < 4; i++) { t[i] = (uint8_t)(c >> (8 * i)); t[4 + i] = (uint8_t)(n >> (8 * i)); }
		under->write(t, 0, 8);
		ended = true;
	}
	void close() override {
		if (!ended) finish();
		under->close();
	}
	const ChunkIndex &chunkIndex() const { return deflater.chunkIndex(); }
private:
	OutputStream *under;
	DeflaterOutputStream deflater;
	bool ended = false;
};

// ---------------------------------------------------------------- GzipInputStream
class GzipInputStream : public InputStream {
public:
	explicit GzipInputStream(InputStream &in) : raw(&in), metadata(GzipMetadata::read(in)) {        // GzipInputStream.java:38-45
		if (!in.markSupported()) throw IllegalArgumentException("Input stream not markable");        // (the reference wraps it in a BufferedInputStream)
		inflater.reset(new InflaterInputStream(in, true));
		inflater->setTrailerBytes(0);
		ChunkIndex idx = metadata.chunkIndex();
		if (!idx.empty()) inflater->setChunkIndex(std::move(idx));
	}
	const GzipMetadata &getMetadata() const { return metadata; }                          // :51
	void setOutputSizeHint(uint64_t n) { if (inflater) inflater->setOutputSizeHint(n); }
	int read() override { uint8_t b; return read(&b, 0, 1) == 1 ? b : -1; }
	long read(uint8_t *b, size_t off, size_t len) override {                              // :66-90
		if (!inflater) return -1;
		long r = inflater->read(b, off, len);
		if (r != -1) { length += (uint64_t)(r > 0 ? r : 0); return r; }
		const uint32_t crc = inflater->crc32();
		inflater.reset();                                  // the raw stream now stands right after the DEFLATE data
		uint8_t t[8];
		for (int i = 0; i < 8; i++) {
			int v = raw->read();
			if (v < 0) throw DataFormatException::unexpectedEnd();
			t[i] = (uint8_t)v;
		}
		uint32_t ec = (uint32_t)t[0] | (uint32_t)t[1] << 8 | (uint32_t)t[2] << 16 | (uint32_t)t[3] << 24;
		uint32_t el = (uint32_t)t[4] | (uint32_t)t[5] << 8 | (uint32_t)t[6] << 16 | (uint32_t)t[7] << 24;
		if (crc != ec) throw DataFormatException(DataFormatException::Reason::DECOMPRESSED_CHECKSUM_MISMATCH, "Decompression CRC-32 mismatch");
		if ((uint32_t)length != el) throw DataFormatException(DataFormatException::Reason::DECOMPRESSED_SIZE_MISMATCH, "Decompressed size mismatch");
		return -1;
	}
	void close() override { inflater.reset(); raw->close(); }
private:
	InputStream *raw;
	GzipMetadata metadata;
	std::unique_ptr<InflaterInputStream> inflater;
	uint64_t length = 0;
};

// ---------------------------------------------------------------- zlib container (RFC 1950)
struct ZlibMetadata {                                                                     // ZlibMetadata.java:19-47
	enum class CompressionMethod { DEFLATE, RESERVED };
	enum class CompressionLevel { FASTEST, FAST, DEFAULT, MAXIMUM };
	CompressionMethod compressionMethod = CompressionMethod::DEFLATE;
	int compressionInfo = 7;
	std::optional<uint32_t> presetDictionary;
	CompressionLevel compressionLevel = CompressionLevel::DEFAULT;

	ZlibMetadata() {}
	ZlibMetadata(CompressionMethod cm, int info, std::optional<uint32_t> dict, CompressionLevel lvl)
	    : compressionMethod(cm), compressionInfo(info), presetDictionary(dict), compressionLevel(lvl) {
		if (((unsigned)info >> 4) != 0 || (cm == CompressionMethod::DEFLATE && info > 7))
			throw IllegalArgumentException("Invalid compression info value");
	}
	static ZlibMetadata read(InputStream &in) {                                           // ZlibMetadata.java:47-80
		int cmf = in.read(), flg = in.read();
		if (flg == -1) throw DataFormatException::unexpectedEnd();
		if ((cmf << 8 | flg) % 31 != 0)
			throw DataFormatException(DataFormatException::Reason::HEADER_CHECKSUM_MISMATCH, "Header checksum mismatch");
		ZlibMetadata m;
		int method = cmf & 0xF;
		if (method == 8) m.compressionMethod = CompressionMethod::DEFLATE;
		else if (method == 15) m.compressionMethod = CompressionMethod::RESERVED;
		else throw DataFormatException(DataFormatException::Reason::UNSUPPORTED_COMPRESSION_METHOD,
		                               "Unsupported compression method: " + std::to_string(method));
		const int info = cmf >> 4;
		if ((flg >> 5) & 1) {
			uint32_t val = 0;
			for (int i = 0; i < 4; i++) {
				int b = in.read();
				if (b == -1) throw DataFormatException::unexpectedEnd();
				val = val << 8 | (uint32_t)b;
			}
			m.presetDictionary = val;
		}
		m.compressionLevel = (CompressionLevel)(flg >> 6);
		if (m.compressionMethod == CompressionMethod::DEFLATE && info > 7)                // the record constructor's check (:24-25)
			throw IllegalArgumentException("Invalid compression info value");
		m.compressionInfo = info;
		return m;
	}
	void write(OutputStream &out) const {                                                 // ZlibMetadata.java:86-104
		int cmf = (compressionMethod == CompressionMethod::DEFLATE ? 8 : 15) | compressionInfo << 4;
		int flg = (presetDictionary ? 1 : 0) << 5 | (int)compressionLevel << 6;
		flg |= (31 - (cmf << 8 | flg) % 31) % 31;
		out.write(cmf);
		out.write(flg);
		if (presetDictionary) for (int i = 3; i >= 0; i--) out.write((int)(*presetDictionary >> (i * 8)) & 0xFF);
	}
};

class ZlibOutputStream : public OutputStream {                                            // ZlibOutputStream.java:31-75
public:
	ZlibOutputStream(OutputStream &out, const ZlibMetadata &meta) : under(&out), deflater(out, adlerOptions()) { meta.write(out); }
	using OutputStream::write;
	void write(int b) override { uint8_t x = (uint8_t)b; write(&x, 0, 1); }
	void write(const uint8_t *b, size_t off, size_t len) override {
		if (ended) throw IllegalStateException("Stream already ended");
		deflater.write(b, off, len);                       // the Adler-32 rides the GPU call (ZlibOutputStream.java:56)
	}
	void finish() {                                                                       // :60-67, big-endian trailer
		if (ended) throw IllegalStateException("Stream already ended");
		deflater.finish();
		const uint32_t a = deflater.checksum();
		uint8_t t[4] = {(uint8_t)(a >> 24), (uint8_t)(a >> 16), (uint8_t)(a >> 8), (uint8_t)a};
		under->write(t, 0, 4);
		ended = true;
	}
	void close() override { if (!ended) finish(); under->close(); }
private:
	static DeflaterOptions adlerOptions() { DeflaterOptions o; o.checksum = B2D_CHECKSUM_ADLER32; return o; }
	OutputStream *under;
	DeflaterOutputStream deflater;
	bool ended = false;
};

class ZlibInputStream : public InputStream {                                              // ZlibInputStream.java:36-83
public:
	explicit ZlibInputStream(InputStream &in) : raw(&in), metadata(ZlibMetadata::read(in)) {
		if (!in.markSupported()) throw IllegalArgumentException("Input stream not markable");
		inflater.reset(new InflaterInputStream(in, true));
		inflater->setChecksumAdler32(true);
	}
	const ZlibMetadata &getMetadata() const { return metadata; }
	int read() override { uint8_t b; return read(&b, 0, 1) == 1 ? b : -1; }
	long read(uint8_t *b, size_t off, size_t len) override {                              // :64-83
		if (!inflater) return -1;
		long r = inflater->read(b, off, len);
		if (r != -1) return r;
		const uint32_t got = inflater->checksum();
		inflater.reset();
		uint32_t expect = 0;
		for (int i = 0; i < 4; i++) {
			int v = raw->read();
			if (v < 0) throw DataFormatException::unexpectedEnd();
			expect = expect << 8 | (uint32_t)v;
		}
		if (got != expect)
			throw DataFormatException(DataFormatException::Reason::DECOMPRESSED_CHECKSUM_MISMATCH, "Decompression Adler-32 mismatch");
		return -1;
	}
	void close() override { raw->close(); inflater.reset(); }
private:
	InputStream *raw;
	ZlibMetadata metadata;
	std::unique_ptr<InflaterInputStream> inflater;
};

}  // namespace io_nayuki_deflate
