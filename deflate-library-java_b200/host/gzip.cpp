// gzip.cpp -- `gzip InputFile OutputFile.gz`: the reference's compression CLI (src/gzip.java:27-77) on the GPU codec.
// Same two positional arguments, same messages, exit code 1 + message on stderr on error, same two speed lines.
// Difference under the hood: the body is compressed by b2d_deflate_chunks in 1 MiB chunks and the per-chunk sizes are
// stored in a gzip FEXTRA subfield ("B2") so that gunzip can decode the chunks in parallel; the reference's decoder
// parses and skips extra fields (GzipMetadata.java:116-122), so it still reads these files.  B2D_GZIP_INDEX=0 writes the
// header first and streams the body without an index (exactly the reference's header: FNAME + FHCRC, OS = Unix).
// The file is streamed like gzip.java:67-68 streams it (in.transferTo(out)), in batches (B2D_GZIP_BATCH MiB, default
// 256): while the GPU compresses batch k a second thread reads batch k + 1 into pinned memory and a third writes batch
// k - 1's output; the index goes into a header whose size is known from the file size and which is rewritten in place
// once the chunk sizes are.
#include <chrono>
#include <future>
#include <sys/stat.h>
#include "b2d_streams.hpp"

using namespace io_nayuki_deflate;

static std::string baseName(const std::string &p) {
	size_t k = p.find_last_of('/');
	return k == std::string::npos ? p : p.substr(k + 1);
}

static std::optional<std::string> submain(int argc, char **argv) {
	if (argc != 3) return "Usage: gzip InputFile OutputFile.gz";
	const std::string inPath = argv[1], outPath = argv[2];
	struct stat st;
	if (stat(inPath.c_str(), &st) != 0) return "Input path does not exist: " + inPath;
	if (S_ISDIR(st.st_mode)) return "Input path is a directory: " + inPath;
	struct stat so;
	if (stat(outPath.c_str(), &so) == 0 && S_ISDIR(so.st_mode)) return "Output path is a directory: " + outPath;

	int32_t modTime = (int32_t)st.st_mtime;                                        // gzip.java:51-62
	GzipMetadata meta(GzipMetadata::CompressionMethod::DEFLATE, false,
	                  modTime != 0 ? std::optional<int32_t>(modTime) : std::nullopt, 0, GzipMetadata::OperatingSystem::UNIX,
	                  std::nullopt, baseName(inPath), std::nullopt, true);
	const char *ix = getenv("B2D_GZIP_INDEX");
	const bool withIndex = !(ix && ix[0] == '0');
	DeflaterOptions dopt;                                  // B2D_GZIP_SPLIT=<bytes>: BinarySplit-style adaptive blocks
	if (const char *sp = getenv("B2D_GZIP_SPLIT")) dopt.split_min_bytes = (uint32_t)strtoul(sp, nullptr, 10);

	auto t0 = std::chrono::steady_clock::now();
	uint64_t outBytes = 0;
	try {
		MarkableFileInputStream in(inPath);
		FileOutputStream fout(outPath);
		std::vector<uint8_t> buf(8 << 20);
		if (!withIndex) {
			GzipOutputStream out(fout, meta, dopt);
			for (long r; (r = in.read(buf.data(), 0, buf.size())) > 0;) out.write(buf.data(), 0, (size_t)r);   // in.transferTo(out)
			out.close();
		} else {
			const uint64_t size = (uint64_t)st.st_size;
			const uint32_t chunk = dopt.chunk_bytes, bpc = dopt.chunk_bytes / dopt.block_bytes;
			const uint64_t n_chunks = (size + chunk - 1) / chunk;
			ChunkIndex index;                                  // placeholder of the final size first
			index.chunk_bytes = chunk;
			index.block_bytes = dopt.block_bytes;
			index.sizes.assign((size_t)std::max<uint64_t>(n_chunks, 1), 0);       // (an empty file: the 5-byte closing block is its one entry)
			index.block_bits.assign((size_t)(n_chunks * bpc), 0);
			meta.extraField = GzipMetadata::encodeChunkIndex(index);                 // absent if it does not fit 64 KiB
			if (!S_ISREG(st.st_mode) || !meta.extraField) {
				meta.extraField = std::nullopt;
				GzipOutputStream out(fout, meta, dopt);
				for (long r; (r = in.read(buf.data(), 0, buf.size())) > 0;) out.write(buf.data(), 0, (size_t)r);
				out.close();
			} else {
				fout.close();
				requireDevice();
				FILE *fi = fopen(inPath.c_str(), "rb"), *fo = fopen(outPath.c_str(), "wb");
				if (!fi || !fo) { if (fi) fclose(fi); if (fo) fclose(fo); throw IOException("Cannot open " + (fi ? outPath : inPath)); }
				struct Closer { FILE *a, *b; ~Closer() { if (a) fclose(a); if (b) fclose(b); } } closer{fi, fo};
				ByteArrayOutputStream hdr;
				meta.write(hdr);
				const size_t hdr_len = hdr.toByteArray().size();
				if (fwrite(hdr.toByteArray().data(), 1, hdr_len, fo) != hdr_len) throw IOException("write error");
				uint64_t batch = 256ull << 20;
				if (const char *bm = getenv("B2D_GZIP_BATCH")) batch = std::max<uint64_t>(1, strtoull(bm, nullptr, 10)) << 20;
				batch = std::max<uint64_t>(chunk, batch / chunk * chunk);
				const uint64_t n_batches = std::max<uint64_t>(1, (size + batch - 1) / batch);
				PinnedBuffer inb[2], outb[2];
				auto read_batch = [&](uint64_t k) {
					const uint64_t len = std::min<uint64_t>(batch, size - k * batch);
					inb[k & 1].reserve(len + 64);
					if (fread(inb[k & 1].p, 1, len, fi) != len) throw IOException("read error");
					return len;
				};
				b2d_deflate_opts o;
				memset(&o, 0, sizeof o);
				o.chunk_bytes = chunk; o.block_bytes = dopt.block_bytes; o.mode = dopt.mode; o.search = dopt.search;
				o.chain_depth = dopt.chain_depth; o.lazy = dopt.lazy; o.framing = B2D_FRAMING_CHUNKED; o.checksum = B2D_CHECKSUM_CRC32;
				o.split_min_bytes = dopt.split_min_bytes;
				uint32_t crc = 0;
				size_t ci = 0;
				std::future<uint64_t> reading = std::async(std::launch::async, read_batch, (uint64_t)0);
				std::future<void> writing;
				for (uint64_t k = 0; k < n_batches; k++) {
					const uint64_t len = reading.get();
					if (k + 1 < n_batches) reading = std::async(std::launch::async, read_batch, k + 1);
					const uint64_t bound = b2d_deflate_bound(len, chunk);
					if (writing.valid() && k >= 2) writing.get();              // outb[k & 1] was batch k - 2's: its write is long done,
					outb[k & 1].reserve(bound);                                // but make sure before the buffer is reused
					const size_t nc = (size_t)((len + chunk - 1) / chunk), nb = (size_t)((len + dopt.block_bytes - 1) / dopt.block_bytes);
					std::vector<uint64_t> sizes(std::max<size_t>(nc, 1));
					std::vector<uint32_t> bits(std::max<size_t>(nb, 1));
					o.is_last = k + 1 == n_batches;
					const int64_t n = b2d_deflate_chunks_indexed(inb[k & 1].p, len, &o, outb[k & 1].p, bound, &crc, sizes.data(), bits.data());
					if (n < 0) throw IOException(std::string("b2d_deflate_chunks: ") + b2d_strerror((int)n) + " [" + b2d_last_error() + "]");
					for (size_t c = 0; c < nc; c++, ci++) {
						index.sizes[ci] = sizes[c];
						for (size_t b = 0; b < bpc; b++) index.block_bits[ci * bpc + b] = c * bpc + b < nb ? bits[c * bpc + b] : 0u;
					}
					if (nc == 0) index.sizes[0] = (uint64_t)n;
					if (writing.valid()) writing.get();
					const uint8_t *src = outb[k & 1].p;
					writing = std::async(std::launch::async, [fo, src, n] { if (fwrite(src, 1, (size_t)n, fo) != (size_t)n) throw IOException("write error"); });
				}
				if (writing.valid()) writing.get();
				uint8_t t[8];
				const uint32_t n32 = (uint32_t)size;                                      // GzipOutputStream.java:62-70
				for (int i = 0; i < 4; i++) { t[i] = (uint8_t)(crc >> (8 * i)); t[4 + i] = (uint8_t)(n32 >> (8 * i)); }
				if (fwrite(t, 1, 8, fo) != 8) throw IOException("write error");
				meta.extraField = GzipMetadata::encodeChunkIndex(index);                  // the real index, same size
				ByteArrayOutputStream hdr2;
				meta.write(hdr2);
				if (hdr2.toByteArray().size() != hdr_len) throw IOException("gzip header changed size");
				if (fseek(fo, 0, SEEK_SET) != 0 || fwrite(hdr2.toByteArray().data(), 1, hdr_len, fo) != hdr_len) throw IOException("write error");
				if (fflush(fo) != 0) throw IOException("write error");
			}
		}
		in.close();
		struct stat s2;
		if (stat(outPath.c_str(), &s2) == 0) outBytes = (uint64_t)s2.st_size;
	} catch (const IOException &e) {
		return std::string("I/O exception: ") + e.what();
	}
	double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	fprintf(stderr, "Input  speed: %.2f MB/s\n", (double)st.st_size / 1e6 / sec);   // gzip.java:73-74
	fprintf(stderr, "Output speed: %.2f MB/s\n", (double)outBytes / 1e6 / sec);
	return std::nullopt;
}

int main(int argc, char **argv) {
	std::optional<std::string> msg;
	try { msg = submain(argc, argv); }
	catch (const std::exception &e) { msg = std::string("Exception: ") + e.what(); }
	if (msg) { fprintf(stderr, "%s\n", msg->c_str()); return 1; }                   // gzip.java:27-33
	return 0;
}
