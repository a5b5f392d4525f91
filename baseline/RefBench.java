/*
 * RefBench.java -- the reference's own Java path on the host cores, for a box that has a JDK >= 18 (this image has
 * none: `command -v java` finds nothing, so bench.py's CPU arm is the C restatement under oracle/ and says so).
 *
 *   javac -cp /root/reference/src -d /tmp/refbench baseline/RefBench.java $(find /root/reference/src/io -name '*.java')
 *   java  -cp /tmp/refbench RefBench inflate <dir with member_*.deflate files> [threads]
 *   java  -cp /tmp/refbench RefBench deflate <raw input file> [threads] [RLE_DYNAMIC|FULL_DYNAMIC]
 *
 * inflate: every file is one raw-DEFLATE member (bench.py's members without the gzip header / trailer); each thread
 * decodes its share with io.nayuki.deflate.InflaterInputStream into a reused byte[] (no per-byte sink: the
 * FileOutputStream.write(int) artefact of src/gzip.java:67 is deliberately not measured) and CRC-32s the output
 * (java.util.zip.CRC32, what GzipInputStream.java:72 does).  Prints single-thread and all-thread GB/s of uncompressed data.
 * deflate: the input is cut into independent 1 MiB chunks, each through DeflaterOutputStream(out, 65536, 32768, strategy)
 * into a ByteArrayOutputStream.
 */
import java.io.ByteArrayInputStream;
import java.io.ByteArrayOutputStream;
import java.io.IOException;
import java.nio.file.Files;
import java.nio.file.Path;
import java.util.ArrayList;
import java.util.List;
import java.util.concurrent.ExecutorService;
import java.util.concurrent.Executors;
import java.util.concurrent.Future;
import java.util.zip.CRC32;

import io.nayuki.deflate.DeflaterOutputStream;
import io.nayuki.deflate.InflaterInputStream;
import io.nayuki.deflate.comp.Lz77Huffman;
import io.nayuki.deflate.comp.Strategy;

public final class RefBench {

	public static void main(String[] args) throws Exception {
		if (args.length < 2) {
			System.err.println("usage: RefBench inflate <dir> [threads] | deflate <file> [threads] [strategy]");
			System.exit(2);
		}
		int threads = args.length > 2 ? Integer.parseInt(args[2]) : Runtime.getRuntime().availableProcessors();
		if (args[0].equals("inflate")) inflate(Path.of(args[1]), threads);
		else deflate(Path.of(args[1]), threads, args.length > 3 ? args[3] : "RLE_DYNAMIC");
	}

	private static void inflate(Path dir, int threads) throws Exception {
		List<byte[]> members = new ArrayList<>();
		try (var s = Files.list(dir)) {
			for (Path p : (Iterable<Path>)s.sorted()::iterator) members.add(Files.readAllBytes(p));
		}
		for (int t : new int[]{1, threads}) {
			for (int rep = 0; rep < 3; rep++) {                 // the last repetition is reported (JIT warm)
				long t0 = System.nanoTime();
				long bytes = run(t, members.size(), i -> {
					byte[] buf = new byte[1 << 16];
					CRC32 crc = new CRC32();
					long n = 0;
					try (InflaterInputStream in = new InflaterInputStream(new ByteArrayInputStream(members.get(i)))) {
						for (int r; (r = in.read(buf, 0, buf.length)) != -1;) { crc.update(buf, 0, r); n += r; }
					}
					return n;
				});
				double sec = (System.nanoTime() - t0) / 1e9;
				if (rep == 2) System.out.printf("inflate threads=%d: %.4f GB/s uncompressed (%d members)%n", t, bytes / sec / 1e9, members.size());
			}
		}
	}

	private static void deflate(Path file, int threads, String strategyName) throws Exception {
		byte[] data = Files.readAllBytes(file);
		Strategy strategy = (Strategy)Lz77Huffman.class.getField(strategyName).get(null);
		int chunk = 1 << 20, n = (data.length + chunk - 1) / chunk;
		for (int t : new int[]{1, threads}) {
			for (int rep = 0; rep < 3; rep++) {
				long t0 = System.nanoTime();
				long out = run(t, n, i -> {
					ByteArrayOutputStream sink = new ByteArrayOutputStream(chunk);
					try (DeflaterOutputStream d = new DeflaterOutputStream(sink, 65536, 32768, strategy)) {
						d.write(data, i * chunk, Math.min(chunk, data.length - i * chunk));
					}
					return (long)sink.size();
				});
				double sec = (System.nanoTime() - t0) / 1e9;
				if (rep == 2) System.out.printf("deflate %s threads=%d: %.4f GB/s uncompressed, ratio %.4f%n", strategyName, t,
					data.length / sec / 1e9, (double)data.length / out);
			}
		}
	}

	private interface Unit { long run(int i) throws IOException; }

	private static long run(int threads, int units, Unit u) throws Exception {
		ExecutorService ex = Executors.newFixedThreadPool(threads);
		List<Future<Long>> fs = new ArrayList<>();
		for (int t = 0; t < threads; t++) {
			final int t0 = t;
			fs.add(ex.submit(() -> { long s = 0; for (int i = t0; i < units; i += threads) s += u.run(i); return s; }));
		}
		long total = 0;
		for (Future<Long> f : fs) total += f.get();
		ex.shutdown();
		return total;
	}
}
