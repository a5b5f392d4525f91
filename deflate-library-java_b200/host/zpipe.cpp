// zpipe.cpp -- `zpipe -c In Out.zz` / `zpipe -d In.zz Out`: file front-end of ZlibOutputStream / ZlibInputStream (the
// reference ships no zlib CLI; this exists so the zlib container of the host mirror can be driven from the tests).
#include "b2d_streams.hpp"

using namespace io_nayuki_deflate;

int main(int argc, char **argv) {
	if (argc != 4 || (std::string(argv[1]) != "-c" && std::string(argv[1]) != "-d")) {
		fprintf(stderr, "Usage: zpipe -c|-d InputFile OutputFile\n");
		return 1;
	}
	try {
		MarkableFileInputStream in(argv[2]);
		FileOutputStream fout(argv[3]);
		std::vector<uint8_t> buf(8 << 20);
		if (std::string(argv[1]) == "-c") {
			ZlibOutputStream out(fout, ZlibMetadata());
			for (long r; (r = in.read(buf.data(), 0, buf.size())) > 0;) out.write(buf.data(), 0, (size_t)r);
			out.close();
		} else {
			ZlibInputStream zin(in);
			for (long r; (r = zin.read(buf.data(), 0, buf.size())) != -1;) fout.write(buf.data(), 0, (size_t)r);
			fout.close();
		}
	} catch (const DataFormatException &e) {
		fprintf(stderr, "Exception: DataFormatException: %s\n", e.what());
		return 1;
	} catch (const std::exception &e) {
		fprintf(stderr, "Exception: %s\n", e.what());
		return 1;
	}
	return 0;
}
