// stream_tests.cpp -- the reference's own stream tests, re-run against the host mirror (b2d_streams.hpp) on the GPU.
//   test/io/nayuki/deflate/InflaterInputStreamTest.java:513-593  harness: decode every vector twice (byte-at-a-time
//       read(), then read(b, off, len) with random off/len incl. len == 0), always with endExactly, then require the
//       underlying stream to stand exactly at its end (:557-558); failures must carry the expected Reason (:587-593).
//   test/io/nayuki/deflate/DeflaterOutputStreamTest.java:24-115   five round-trip tests (empty, short random, mixed
//       write(int) / write(b,off,len), byte runs, long random with mostly single-byte writes).
//   plus the argument / state contract of InflaterInputStream.java:96-106,147-179 and DeflaterOutputStream.java:55-116,
//   and the GPU build's batched decode of indexed streams (input read batch by batch from the underlying stream).
// Usage: stream_tests <vectors.txt>   (lines: name hexbytes ok|fail hexoutput|REASON ; "-" = empty)
// Prints one line per failed check and "passed N checks" at the end; exit code 1 if any check failed.
#include <fstream>
#include <random>
#include <sstream>
#include "b2d_streams.hpp"

using namespace io_nayuki_deflate;

static int checks = 0, failures = 0;
#define CHECK(cond, what)                                                         \
	do {                                                                          \
		checks++;                                                                 \
		if (!(cond)) { failures++; fprintf(stderr, "FAIL %s: %s\n", what, #cond); } \
	} while (0)

static const char *REASONS[] = {
	"UNEXPECTED_END_OF_STREAM", "RESERVED_BLOCK_TYPE", "UNCOMPRESSED_BLOCK_LENGTH_MISMATCH", "HUFFMAN_CODE_UNDER_FULL",
	"HUFFMAN_CODE_OVER_FULL", "NO_PREVIOUS_CODE_LENGTH_TO_COPY", "CODE_LENGTH_CODE_OVER_FULL", "END_OF_BLOCK_CODE_ZERO_LENGTH",
	"RESERVED_LENGTH_SYMBOL", "RESERVED_DISTANCE_SYMBOL", "LENGTH_ENCOUNTERED_WITH_EMPTY_DISTANCE_CODE",
	"COPY_FROM_BEFORE_DICTIONARY_START", "HEADER_CHECKSUM_MISMATCH", "UNSUPPORTED_COMPRESSION_METHOD",
	"DECOMPRESSED_CHECKSUM_MISMATCH", "DECOMPRESSED_SIZE_MISMATCH", "GZIP_INVALID_MAGIC_NUMBER", "GZIP_RESERVED_FLAGS_SET",
	"GZIP_UNSUPPORTED_OPERATING_SYSTEM"};

static std::vector<uint8_t> unhex(const std::string &h) {
	std::vector<uint8_t> v;
	if (h == "-") return v;
	for (size_t i = 0; i + 1 < h.size(); i += 2) v.push_back((uint8_t)std::stoi(h.substr(i, 2), nullptr, 16));
	return v;
}

static std::mt19937 rng(20261018);
static int rnd(int n) { return (int)(rng() % (unsigned)n); }

// InflaterInputStreamTest.java:537-582
static void decodeBothWays(const std::string &name, const std::vector<uint8_t> &in, bool expectOk,
                           const std::vector<uint8_t> &expectOut, const std::string &expectReason) {
	for (int mode = 0; mode < 2; mode++) {
		ByteArrayInputStream sin(in);
		std::vector<uint8_t> got;
		std::string reason;
		try {
			InflaterInputStream iin(sin, true);
			if (mode == 0) {
				for (int b; (b = iin.read()) != -1;) got.push_back((uint8_t)b);
			} else {
				for (;;) {
					std::vector<uint8_t> buf(rnd(100) + 1);
					size_t off = (size_t)rnd((int)buf.size() + 1);
					size_t len = (size_t)rnd((int)(buf.size() - off) + 1);
					long n = iin.read(buf.data(), off, len);
					if (n == -1) break;
					CHECK(n >= 0 && (size_t)n <= len && (n > 0 || len == 0), name.c_str());      // 0 iff len == 0
					got.insert(got.end(), buf.begin() + off, buf.begin() + off + n);
				}
			}
		} catch (const DataFormatException &e) {
			reason = REASONS[(int)e.getReason()];
		}
		if (expectOk) {
			CHECK(reason.empty(), name.c_str());
			CHECK(got == expectOut, name.c_str());
			CHECK(sin.position() == in.size(), name.c_str());                            // end exactly (:557-558)
		} else {
			CHECK(reason == expectReason, name.c_str());
		}
	}
}

static std::vector<uint8_t> randomBytes(size_t n) {
	std::vector<uint8_t> v(n);
	for (auto &b : v) b = (uint8_t)rng();
	return v;
}

// DeflaterOutputStreamTest.checkInflate: the stream must decode back to the input
static void checkInflate(const std::vector<uint8_t> &data, const std::vector<uint8_t> &comp, const char *what) {
	ByteArrayInputStream bin(comp);
	InflaterInputStream iin(bin, true);
	std::vector<uint8_t> back = iin.readAllBytes();
	CHECK(back == data, what);
	CHECK(bin.position() == comp.size(), what);
}

static void deflaterTests() {
	{   // testEmpty (:24-29)
		ByteArrayOutputStream bout;
		DeflaterOutputStream dout(bout);
		dout.close();
		checkInflate({}, bout.toByteArray(), "testEmpty");
	}
	for (int t = 0; t < 30; t++) {   // testShortSingleWriteRandomly (:32-44), fewer trials (each is a GPU round trip)
		std::vector<uint8_t> data = randomBytes(rnd(100));
		ByteArrayOutputStream bout;
		DeflaterOutputStream dout(bout);
		dout.write(data.data(), 0, data.size());
		dout.close();
		checkInflate(data, bout.toByteArray(), "testShortSingleWriteRandomly");
	}
	for (int t = 0; t < 30; t++) {   // testShortMultiWriteRandomly (:47-64)
		std::vector<uint8_t> data = randomBytes(rnd(1000));
		ByteArrayOutputStream bout;
		DeflaterOutputStream dout(bout);
		for (size_t off = 0; off < data.size();) {
			if (rnd(2)) { dout.write(data[off]); off++; }
			else { size_t n = (size_t)rnd((int)(data.size() - off)) + 1; dout.write(data.data(), off, n); off += n; }
		}
		dout.close();
		checkInflate(data, bout.toByteArray(), "testShortMultiWriteRandomly");
	}
	{   // testByteRunsRandomly (:67-86)
		std::vector<uint8_t> data;
		for (int i = 0; i < 1000; i++) data.insert(data.end(), (size_t)rnd(1000) + 1, (uint8_t)rng());
		ByteArrayOutputStream bout;
		DeflaterOutputStream dout(bout);
		dout.write(data.data(), 0, data.size());
		dout.close();
		checkInflate(data, bout.toByteArray(), "testByteRunsRandomly");
		CHECK(bout.toByteArray().size() < data.size() / 20, "testByteRunsRandomly compresses");
	}
	for (int t = 0; t < 4; t++) {   // testLongRandomly (:89-115): mostly single-byte writes, bulk writes up to 300000
		std::vector<uint8_t> data = randomBytes((size_t)rnd(1000000));
		ByteArrayOutputStream bout;
		DeflaterOutputStream dout(bout);
		for (size_t off = 0; off < data.size();) {
			if (rnd(10) < 9) { dout.write(data[off]); off++; }
			else { size_t n = std::min<size_t>((size_t)rnd(300000) + 1, data.size() - off); dout.write(data.data(), off, n); off += n; }
		}
		dout.close();
		checkInflate(data, bout.toByteArray(), "testLongRandomly");
	}
}

// an InputStream that hands out a few hundred bytes per call (a pipe, a socket): the batched reader has to loop
struct DribbleStream : ByteArrayInputStream {
	using ByteArrayInputStream::ByteArrayInputStream;
	long read(uint8_t *b, size_t off, size_t len) override { return ByteArrayInputStream::read(b, off, std::min<size_t>(len, 777)); }
};

// The GPU build's own extension: a stream with a chunk + block index and a known size is decoded in batches, and the
// compressed bytes of batch k + 1 are READ from the underlying stream while batch k is consumed (B2D_GUNZIP_BATCH = 3
// chunks here).  Same bytes, same checksum, same endExactly position as the one-shot decode; a stream that ends early
// delivers what was whole and then throws UNEXPECTED_END_OF_STREAM.
static void indexedBatchTests() {
	std::vector<uint8_t> data;
	for (int i = 0; i < 3000; i++) {                       // runs and noise: chunks of very different compressed sizes
		if (rnd(3)) data.insert(data.end(), (size_t)rnd(400) + 1, (uint8_t)rng());
		else { std::vector<uint8_t> r = randomBytes((size_t)rnd(300)); data.insert(data.end(), r.begin(), r.end()); }
	}
	DeflaterOptions o;
	o.chunk_bytes = 1u << 16;
	o.block_bytes = 1u << 14;
	o.batch_bytes = 1u << 18;
	ByteArrayOutputStream bout;
	DeflaterOutputStream dout(bout, o);
	dout.write(data.data(), 0, data.size());
	dout.close();
	const std::vector<uint8_t> comp = bout.toByteArray();
	const ChunkIndex idx = dout.chunkIndex();
	CHECK(idx.sizes.size() == (data.size() + o.chunk_bytes - 1) / o.chunk_bytes && idx.sizes.size() > 6, "indexed: chunk count");
	uint32_t crcOneShot = 0;
	{   // one-shot reference: no size hint -> not batched
		ByteArrayInputStream bin(comp);
		InflaterInputStream iin(bin, true);
		iin.setChunkIndex(idx);
		CHECK(iin.readAllBytes() == data, "indexed one-shot bytes");
		crcOneShot = iin.crc32();
	}
	for (int variant = 0; variant < 2; variant++) {        // plain and dribbling underlying stream, 9 trailing bytes behind the data
		std::vector<uint8_t> in = comp;
		in.insert(in.end(), 9, (uint8_t)0xA5);
		std::unique_ptr<ByteArrayInputStream> bin(variant ? new DribbleStream(in) : new ByteArrayInputStream(in));
		InflaterInputStream iin(*bin, true);
		iin.setChunkIndex(idx);
		iin.setOutputSizeHint(data.size());
		std::vector<uint8_t> back;
		uint8_t buf[50000];
		for (long r; (r = iin.read(buf, 0, (size_t)rnd(50000) + 1)) != -1;) back.insert(back.end(), buf, buf + r);
		CHECK(back == data, "indexed batched bytes");
		CHECK(iin.crc32() == crcOneShot, "indexed batched checksum");
		CHECK(iin.consumedBytes() == comp.size(), "indexed batched consumed");
		CHECK(bin->position() == comp.size(), "indexed batched endExactly position");
	}
	{   // checksum asked for before everything was read: the remaining batches are taken over, nothing is lost
		ByteArrayInputStream bin(comp);
		InflaterInputStream iin(bin, true);
		iin.setChunkIndex(idx);
		iin.setOutputSizeHint(data.size());
		std::vector<uint8_t> back(1000);
		CHECK(iin.read(back.data(), 0, 1000) == 1000, "indexed early checksum first read");
		CHECK(iin.crc32() == crcOneShot, "indexed early checksum");
		std::vector<uint8_t> rest = iin.readAllBytes();
		back.insert(back.end(), rest.begin(), rest.end());
		CHECK(back == data, "indexed early checksum bytes");
	}
	for (int variant = 0; variant < 2; variant++) {        // the stream ends inside chunk 5 (second batch) / inside chunk 1 (first batch)
		uint64_t cut = 0;
		const size_t cutChunk = variant ? 1 : 5;
		for (size_t c = 0; c < cutChunk; c++) cut += idx.sizes[c];
		cut += idx.sizes[cutChunk] / 2;
		std::vector<uint8_t> in(comp.begin(), comp.begin() + (long)cut);
		DribbleStream bin(in);
		InflaterInputStream iin(bin, true);
		iin.setChunkIndex(idx);
		iin.setOutputSizeHint(data.size());
		std::vector<uint8_t> back;
		uint8_t buf[4096];
		bool thrown = false;
		try {
			for (long r; (r = iin.read(buf, 0, sizeof buf)) != -1;) back.insert(back.end(), buf, buf + r);
		} catch (const DataFormatException &e) {
			thrown = e.getReason() == DataFormatException::Reason::UNEXPECTED_END_OF_STREAM;
		}
		CHECK(thrown, "indexed truncated: UNEXPECTED_END_OF_STREAM");
		const size_t whole = (cutChunk / 3) * 3 * (size_t)o.chunk_bytes;         // the batches in front of the cut
		CHECK(back.size() == whole && std::equal(back.begin(), back.end(), data.begin()), "indexed truncated: whole batches delivered first");
	}
}

template <class E, class F> static bool throws(F f) {
	try { f(); } catch (const E &) { return true; } catch (...) { return false; }
	return false;
}

struct NoMarkStream : InputStream {
	long read(uint8_t *, size_t, size_t len) override { return len ? -1 : 0; }
};

static void contractTests() {
	std::vector<uint8_t> fixedEmpty = {0x03, 0x00};                                        // testFixedHuffmanEmpty
	{
		ByteArrayInputStream in(fixedEmpty);
		CHECK(throws<IllegalArgumentException>([&] { InflaterInputStream x(in, false, 0); }), "inBufLen <= 0");   // :98-99
		NoMarkStream nm;
		CHECK(throws<IllegalArgumentException>([&] { InflaterInputStream x(nm, true); }), "endExactly needs mark");   // :100-103
		InflaterInputStream iin(in, true);
		uint8_t b[4];
		CHECK(iin.read(b, 0, 0) == 0, "len 0 reads 0");
		CHECK(iin.read(b, 0, 4) == -1, "empty stream ends");
		CHECK(iin.read(b, 0, 4) == -1, "-1 again");
		iin.close();
		iin.close();                                                                       // idempotent (:173-179)
		CHECK(throws<IllegalStateException>([&] { iin.read(b, 0, 1); }), "read after close");   // :160-161
	}
	{   // a format error is delivered after the bytes that precede it, and is not sticky (DataFormatException is unchecked)
		std::vector<uint8_t> bad = {0x63, 0x18, 0x05, 0x40, 0x01};                          // testFixedHuffmanEofInDistanceExtensionBits
		ByteArrayInputStream in(bad);
		InflaterInputStream iin(in, true);
		std::vector<uint8_t> got;
		bool threw = false, threwAgain = false;
		try { for (int v; (v = iin.read()) != -1;) got.push_back((uint8_t)v); } catch (const DataFormatException &e) {
			threw = e.getReason() == DataFormatException::Reason::UNEXPECTED_END_OF_STREAM;
		}
		try { iin.read(); } catch (const DataFormatException &) { threwAgain = true; }
		CHECK(threw && threwAgain && got.size() == 262, "prefix delivered, then the Reason, again on retry");
	}
	{
		ByteArrayOutputStream bout;
		DeflaterOutputStream dout(bout);
		dout.write('x');
		dout.finish();
		CHECK(throws<IllegalStateException>([&] { dout.write('y'); }), "write after finish");            // :77-78
		CHECK(throws<IllegalStateException>([&] { dout.finish(); }), "finish twice");                    // :103-104
		CHECK(throws<IllegalStateException>([&] { dout.getUnderlyingStream(); }), "getUnderlyingStream after end");   // :70-71
		DeflaterOptions o;
		o.block_bytes = 1000;
		CHECK(throws<IllegalArgumentException>([&] { DeflaterOutputStream d2(bout, o); }), "invalid capacities");   // :58-60
		checkInflate({'x'}, bout.toByteArray(), "single byte");
	}
	{   // gzip / zlib containers: metadata round trip, trailer checks (GzipInputStream.java:73-88, ZlibInputStream.java:69-80)
		std::vector<uint8_t> data = randomBytes(50000);
		data.insert(data.end(), 100000, (uint8_t)'a');
		ByteArrayOutputStream bout;
		GzipMetadata meta(GzipMetadata::CompressionMethod::DEFLATE, true, 1700000000, 2, GzipMetadata::OperatingSystem::UNIX,
		                  std::vector<uint8_t>{1, 2, 3}, std::string("name.txt"), std::string("a comment"), true);
		GzipOutputStream gout(bout, meta);
		gout.write(data.data(), 0, data.size());
		gout.close();
		std::vector<uint8_t> gz = bout.toByteArray();
		ByteArrayInputStream bin(gz);
		GzipInputStream gin(bin);
		CHECK(gin.getMetadata().fileName.value() == "name.txt" && gin.getMetadata().comment.value() == "a comment" &&
		      gin.getMetadata().isFileText && gin.getMetadata().extraFlags == 2 && gin.getMetadata().hasHeaderCrc &&
		      gin.getMetadata().extraField->size() == 3 && *gin.getMetadata().modificationTimeUnixS == 1700000000, "gzip metadata");
		CHECK(gin.readAllBytes() == data, "gzip round trip");
		CHECK(bin.position() == gz.size(), "gzip consumed everything");
		gz[gz.size() - 6] ^= 1;
		ByteArrayInputStream bin2(gz);
		GzipInputStream gin2(bin2);
		CHECK(throws<DataFormatException>([&] { gin2.readAllBytes(); }), "gzip CRC mismatch");
		CHECK(throws<IllegalArgumentException>([&] { GzipMetadata m2(GzipMetadata::CompressionMethod::DEFLATE, false, 0, 0,
			GzipMetadata::OperatingSystem::UNIX, std::nullopt, std::nullopt, std::nullopt, false); }), "mtime zero rejected");   // GzipMetadata.java:45-48
		ByteArrayOutputStream zout;
		ZlibOutputStream zo(zout, ZlibMetadata());
		zo.write(data.data(), 0, data.size());
		zo.close();
		ByteArrayInputStream zin(zout.toByteArray());
		ZlibInputStream zi(zin);
		CHECK(zi.readAllBytes() == data, "zlib round trip");
		CHECK(zin.position() == zout.toByteArray().size(), "zlib consumed everything");
	}
}

int main(int argc, char **argv) {
	if (argc != 2) { fprintf(stderr, "Usage: stream_tests vectors.txt\n"); return 2; }
	setenv("B2D_GUNZIP_BATCH", "3", 1);                       // indexedBatchTests: batches of three chunks
	try {
		std::ifstream f(argv[1]);
		std::string line;
		int nvec = 0;
		while (std::getline(f, line)) {
			std::istringstream ss(line);
			std::string name, hex, kind, expect;
			if (!(ss >> name >> hex >> kind >> expect)) continue;
			decodeBothWays(name, unhex(hex), kind == "ok", kind == "ok" ? unhex(expect) : std::vector<uint8_t>(), expect);
			nvec++;
		}
		CHECK(nvec > 0, "vectors loaded");
		deflaterTests();
		indexedBatchTests();
		contractTests();
	} catch (const std::exception &e) {
		fprintf(stderr, "FAIL unexpected exception: %s\n", e.what());
		failures++;
	}
	printf("passed %d of %d checks\n", checks - failures, checks);
	return failures ? 1 : 0;
}
