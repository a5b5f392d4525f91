/*
 * The GPU replacement of decomp/Open.java behind InflaterInputStream.read (InflaterInputStream.java:147-164).
 * It implements the same contract as the reference's State.read(byte[],int,int): returns the number of bytes
 * delivered, -1 only after the final block was delivered, throws DataFormatException(Reason) at the point where the
 * reference's decoder would have hit the bad symbol (the bytes before it are delivered first).
 * UNCOMPILED IN THIS REPOSITORY'S IMAGE (no JDK).  INTEGRATION.md shows the three-line change in
 * InflaterInputStream / State that plugs it in.
 */
package io.nayuki.deflate.gpu;

import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

import java.io.IOException;
import java.io.InputStream;
import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import io.nayuki.deflate.DataFormatException;
import io.nayuki.deflate.DataFormatException.Reason;

public final class GpuInflater {

	private final InputStream input;
	private final boolean endExactly;
	private MemorySegment out;          // pinned; decoded bytes [0, outLen)
	private long outLen, pos, consumed;
	private int status, crc32;
	private boolean decoded;

	public GpuInflater(InputStream in, boolean endExactly) {
		this.input = in;
		this.endExactly = endExactly;     // the caller has already called in.mark(0) (InflaterInputStream.java:100-104)
	}

	public int read(byte[] b, int off, int len) throws IOException {
		if (len == 0) return 0;
		if (!decoded) decodeAll();
		if (pos < outLen) {
			int k = (int)Math.min(len, outLen - pos);
			MemorySegment.copy(out, JAVA_BYTE, pos, b, off, k);
			pos += k;
			return k;
		}
		if (status != 0)
			throw new DataFormatException(Reason.values()[status - 1], B2Deflate.strerror(status));
		return -1;
	}

	public int crc32() { return crc32; }
	public long consumedBytes() { return consumed; }

	public void close() throws IOException {
		if (out != null) { B2Deflate.freePinned(out); out = null; }
		input.close();
	}

	private void decodeAll() throws IOException {
		decoded = true;
		byte[] raw = input.readAllBytes();
		B2Deflate.requireDevice();
		MemorySegment in = B2Deflate.allocPinned(raw.length + 64L);
		try (Arena a = Arena.ofConfined()) {
			MemorySegment.copy(raw, 0, in, JAVA_BYTE, 0, raw.length);
			MemorySegment oLen = a.allocate(JAVA_LONG), cons = a.allocate(JAVA_LONG);
			MemorySegment crc = a.allocate(JAVA_INT), st = a.allocate(JAVA_INT);
			long cap = Math.max(1 << 16, 6L * raw.length);
			while (true) {                       // the decompressed size is unknown up front: grow and retry on overflow
				out = B2Deflate.allocPinned(cap + 64);
				// one stream nobody indexed: speculative parallel decode, the sequential decoder behind it (b2d_inflate_stream)
				int rc = B2Deflate.inflateStream(in, raw.length, out, cap, oLen, cons, crc, st, B2Deflate.INFLATE_CRC32);
				if (rc != 0) throw new IOException("b2d_inflate_stream: " + B2Deflate.strerror(rc) + " [" + B2Deflate.lastError() + "]");
				if (st.get(JAVA_INT, 0) != B2Deflate.ERR_OUTPUT_OVERFLOW) break;
				B2Deflate.freePinned(out);
				cap = cap * 2 + (1 << 20);
			}
			status = st.get(JAVA_INT, 0);
			outLen = oLen.get(JAVA_LONG, 0);
			consumed = cons.get(JAVA_LONG, 0);
			crc32 = crc.get(JAVA_INT, 0);
		} finally {
			B2Deflate.freePinned(in);
		}
		if (endExactly && status == 0) {          // Open.finish (Open.java:113-124)
			input.reset();
			input.skipNBytes(consumed);
		}
	}
}
