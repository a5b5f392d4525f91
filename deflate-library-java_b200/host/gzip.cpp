// gzip.cpp -- `gzip InputFile OutputFile.gz`: the reference's compression CLI (src/gzip.java:27-77) on the GPU codec.
// Same two positional arguments, same messages, exit code 1 + message on stderr on error, same two speed lines.
// Difference under the hood: the body is compressed by b2d_deflate_chunks in 1 MiB chunks and the per-chunk sizes are
// stored in a gzip FEXTRA subfield ("B2") so that gunzip can decode the chunks in parallel; the reference's decoder
// parses and skips extra fields (GzipMetadata.java:116-122), so it still reads these files.  B2D_GZIP_INDEX=0 writes the
// header first and streams the body without an index (exactly the reference's header: FNAME + FHCRC, OS = Unix).
#include <chrono>
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
			ByteArrayOutputStream body;
			DeflaterOutputStream def(body, dopt);
			for (long r; (r = in.read(buf.data(), 0, buf.size())) > 0;) def.write(buf.data(), 0, (size_t)r);
			def.finish();
			meta.extraField = GzipMetadata::encodeChunkIndex(def.chunkIndex());      // absent if it does not fit 64 KiB
			meta.write(fout);
			fout.write(body.toByteArray().data(), 0, body.toByteArray().size());
			uint8_t t[8];
			uint32_t c = def.crc32(), n = (uint32_t)def.totalIn();                   // GzipOutputStream.java:62-70
			for (int i = 0; i < 4; i++) { t[i] = (uint8_t)(c >> (8 * i)); t[4 + i] = (uint8_t)(n >> (8 * i)); }
			fout.write(t, 0, 8);
			fout.close();
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
