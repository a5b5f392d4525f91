// gunzip.cpp -- `gunzip InputFile.gz OutputFile`: the reference's decompression CLI (src/gunzip.java:27-109) on the GPU
// codec.  Same arguments, same metadata lines on stderr, same error convention, same speed lines.  Only the first
// member is read and trailing bytes are not checked, like GzipInputStream.java:66-74.
#include <chrono>
#include <ctime>
#include <sys/stat.h>
#include "b2d_streams.hpp"

using namespace io_nayuki_deflate;

static std::optional<std::string> submain(int argc, char **argv) {
	if (argc != 3) return "Usage: gunzip InputFile.gz OutputFile";
	const std::string inPath = argv[1], outPath = argv[2];
	struct stat st;
	if (stat(inPath.c_str(), &st) != 0) return "Input path does not exist: " + inPath;
	if (S_ISDIR(st.st_mode)) return "Input path is a directory: " + inPath;
	struct stat so;
	if (stat(outPath.c_str(), &so) == 0 && S_ISDIR(so.st_mode)) return "Output path is a directory: " + outPath;
	uint64_t outBytes = 0;
	double sec = 0;
	try {
		MarkableFileInputStream fin(inPath);
		GzipInputStream in(fin);
		const GzipMetadata &meta = in.getMetadata();
		if (meta.modificationTimeUnixS) {                                         // gunzip.java:55-93
			time_t t = (time_t)*meta.modificationTimeUnixS;
			char iso[64];
			struct tm tmv;
			gmtime_r(&t, &tmv);
			strftime(iso, sizeof iso, "%Y-%m-%dT%H:%M:%SZ", &tmv);
			fprintf(stderr, "Last modified: %s\n", iso);
		} else fprintf(stderr, "Last modified: N/A\n");
		if (meta.extraFlags == 2) fprintf(stderr, "Extra flags: Maximum compression\n");
		else if (meta.extraFlags == 4) fprintf(stderr, "Extra flags: Fastest compression\n");
		else fprintf(stderr, "Extra flags: Unknown (%d)\n", meta.extraFlags);
		static const char *OS[] = {"FAT filesystem", "Amiga", "VMS", "Unix", "VM/CMS", "Atari TOS", "HPFS filesystem", "Macintosh",
		                           "Z-System", "CP/M", "TOPS-20", "NTFS filesystem", "QDOS", "Acorn RISCOS", "Unknown"};
		fprintf(stderr, "Operating system: %s\n", OS[(int)meta.operatingSystem]);
		fprintf(stderr, "File mode: %s\n", meta.isFileText ? "Text" : "Binary");
		if (meta.extraField) fprintf(stderr, "Extra field: %zu bytes\n", meta.extraField->size());
		if (meta.fileName) fprintf(stderr, "File name: %s\n", meta.fileName->c_str());
		if (meta.comment) fprintf(stderr, "Comment: %s\n", meta.comment->c_str());

		// ISIZE (last 4 bytes of the file, mod 2^32) sizes the output buffer; a wrong hint only costs a retry
		if (st.st_size >= 18) {
			FILE *f = fopen(inPath.c_str(), "rb");
			if (f) {
				uint8_t t[4];
				if (fseek(f, -4, SEEK_END) == 0 && fread(t, 1, 4, f) == 4) {
					uint64_t isize = (uint64_t)t[0] | (uint64_t)t[1] << 8 | (uint64_t)t[2] << 16 | (uint64_t)t[3] << 24;
					if (isize > 0) in.setOutputSizeHint(isize);
				}
				fclose(f);
			}
		}
		auto t0 = std::chrono::steady_clock::now();
		{
			FileOutputStream out(outPath);
			std::vector<uint8_t> buf(8 << 20);
			for (long r; (r = in.read(buf.data(), 0, buf.size())) != -1;) { out.write(buf.data(), 0, (size_t)r); outBytes += (uint64_t)r; }
			out.close();
		}
		sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		in.close();
	} catch (const IOException &e) {
		return std::string("I/O exception: ") + e.what();
	}
	fprintf(stderr, "Input  speed: %.2f MB/s\n", (double)st.st_size / 1e6 / sec);   // gunzip.java:102-103
	fprintf(stderr, "Output speed: %.2f MB/s\n", (double)outBytes / 1e6 / sec);
	return std::nullopt;
}

int main(int argc, char **argv) {
	std::optional<std::string> msg;
	try { msg = submain(argc, argv); }
	catch (const DataFormatException &e) { msg = std::string("Exception: DataFormatException: ") + e.what(); }
	catch (const std::exception &e) { msg = std::string("Exception: ") + e.what(); }
	if (msg) { fprintf(stderr, "%s\n", msg->c_str()); return 1; }
	return 0;
}
