/*
 * Batch decode of independent raw-DEFLATE members (gzip members, or the chunks of a GpuDeflaterOutputStream stream with
 * its chunk index): the high-throughput entry point -- one warp per member on the GPU.  Each member gets the same
 * outcome a `new InflaterInputStream(in, true)` would give it: bytes, consumed input, CRC-32, and a Reason on failure
 * (a bad member does not disturb its neighbours).  UNCOMPILED IN THIS REPOSITORY'S IMAGE (no JDK).
 */
package io.nayuki.deflate.gpu;

import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

import java.io.IOException;
import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import io.nayuki.deflate.DataFormatException;
import io.nayuki.deflate.DataFormatException.Reason;

public final class BatchInflater {

	public record Result(byte[] data, long consumed, int crc32, Reason failure) {
		public byte[] dataOrThrow() {
			if (failure != null) throw new DataFormatException(failure, failure.toString());
			return data;
		}
	}

	/** members[i] is decoded into at most capacities[i] bytes. */
	public static Result[] inflate(byte[][] members, long[] capacities, boolean chunkIndexed) throws IOException {
		int n = members.length;
		long inTotal = 0, outTotal = 0;
		for (int i = 0; i < n; i++) { inTotal += members[i].length; outTotal += capacities[i]; }
		B2Deflate.requireDevice();
		MemorySegment in = B2Deflate.allocPinned(inTotal + 64), out = B2Deflate.allocPinned(outTotal + 64);
		try (Arena a = Arena.ofConfined()) {
			MemorySegment inOff = a.allocate(JAVA_LONG, n + 1L), outOff = a.allocate(JAVA_LONG, n + 1L);
			MemorySegment oLen = a.allocate(JAVA_LONG, Math.max(n, 1)), cons = a.allocate(JAVA_LONG, Math.max(n, 1));
			MemorySegment crc = a.allocate(JAVA_INT, Math.max(n, 1)), st = a.allocate(JAVA_INT, Math.max(n, 1));
			long ip = 0, op = 0;
			for (int i = 0; i < n; i++) {
				MemorySegment.copy(members[i], 0, in, JAVA_BYTE, ip, members[i].length);
				ip += members[i].length; op += capacities[i];
				inOff.setAtIndex(JAVA_LONG, i + 1L, ip);
				outOff.setAtIndex(JAVA_LONG, i + 1L, op);
			}
			int flags = B2Deflate.INFLATE_CRC32 | (chunkIndexed ? B2Deflate.INFLATE_CHUNK_INDEXED : 0);
			int rc = B2Deflate.inflateBatch(in, inOff, n, out, outOff, oLen, cons, crc, st, flags);
			if (rc != 0) throw new IOException("b2d_inflate_batch: " + B2Deflate.strerror(rc) + " [" + B2Deflate.lastError() + "]");
			Result[] res = new Result[n];
			for (int i = 0; i < n; i++) {
				int s = st.getAtIndex(JAVA_INT, i);
				if (s < 0) throw new IOException("member " + i + ": " + B2Deflate.strerror(s));   // capacity too small
				byte[] d = new byte[(int)oLen.getAtIndex(JAVA_LONG, i)];
				MemorySegment.copy(out, JAVA_BYTE, outOff.getAtIndex(JAVA_LONG, i), d, 0, d.length);
				res[i] = new Result(d, cons.getAtIndex(JAVA_LONG, i), crc.getAtIndex(JAVA_INT, i), s == 0 ? null : Reason.values()[s - 1]);
			}
			return res;
		} finally {
			B2Deflate.freePinned(in);
			B2Deflate.freePinned(out);
		}
	}

	private BatchInflater() {}
}
