/*
 * The GPU replacement of DeflaterOutputStream.writeBuffer -> Strategy.decide -> Decision.compressTo
 * (DeflaterOutputStream.java:119-137): same public surface (write / finish / close, IllegalStateException after the
 * end), but input is buffered `batchBytes` at a time in pinned memory and compressed by ONE b2d_deflate_chunks call
 * into independent 1 MiB chunks joined by empty stored blocks.  UNCOMPILED IN THIS REPOSITORY'S IMAGE (no JDK).
 */
package io.nayuki.deflate.gpu;

import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

import java.io.IOException;
import java.io.OutputStream;
import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.util.Objects;

public final class GpuDeflaterOutputStream extends OutputStream {

	private OutputStream output;
	private final int chunkBytes, blockBytes;
	private int splitMinBytes = 0;
	private final long batchBytes;
	private final MemorySegment stage, comp;
	private long fill, totalIn;
	private int crc32;
	private boolean isEnded;

	public GpuDeflaterOutputStream(OutputStream out) {
		this(out, 1 << 20, 1 << 16, 256L << 20);
	}

	public GpuDeflaterOutputStream(OutputStream out, int chunkBytes, int blockBytes, long batchBytes) {
		this.output = Objects.requireNonNull(out);
		if (blockBytes < 4096 || chunkBytes % blockBytes != 0 || batchBytes < chunkBytes || batchBytes % chunkBytes != 0)
			throw new IllegalArgumentException("Invalid capacities");
		this.chunkBytes = chunkBytes;
		this.blockBytes = blockBytes;
		this.batchBytes = batchBytes;
		stage = B2Deflate.allocPinned(batchBytes);
		comp = B2Deflate.allocPinned(B2Deflate.deflateBound(batchBytes, chunkBytes));
	}

	public int crc32() { return crc32; }          // CRC-32 of everything compressed so far (GzipOutputStream.java:57,67)
	public long totalIn() { return totalIn; }

	/** Counterpart of wrapping the strategy in {@code new BinarySplit(strategy, n)}: adaptive blocks with pieces of n bytes
	 *  (a power of two >= 4096 dividing the block size, at most 16 per block); 0 switches it off. */
	public void setSplitMinBytes(int n) {
		if (n != 0 && (n < 4096 || (n & (n - 1)) != 0 || blockBytes % n != 0 || blockBytes / n > 16))
			throw new IllegalArgumentException("Invalid piece size");
		splitMinBytes = n;
	}

	@Override public void write(int b) throws IOException {
		write(new byte[]{(byte)b}, 0, 1);
	}

	@Override public void write(byte[] b, int off, int len) throws IOException {
		if (isEnded) throw new IllegalStateException("Stream already ended");
		Objects.checkFromIndexSize(off, len, b.length);
		while (len > 0) {
			if (fill == batchBytes) flushBatch(false);
			int k = (int)Math.min(len, batchBytes - fill);
			MemorySegment.copy(b, off, stage, JAVA_BYTE, fill, k);
			fill += k; off += k; len -= k;
		}
	}

	public void finish() throws IOException {
		if (isEnded) throw new IllegalStateException("Stream already ended");
		flushBatch(true);
		isEnded = true;
	}

	@Override public void close() throws IOException {
		if (!isEnded) finish();
		B2Deflate.freePinned(stage);
		B2Deflate.freePinned(comp);
		output.close();
		output = null;
	}

	private void flushBatch(boolean last) throws IOException {
		if (fill == 0 && !last) return;
		try (Arena a = Arena.ofConfined()) {
			MemorySegment o = a.allocate(B2Deflate.DEFLATE_OPTS);
			o.set(JAVA_INT, 0, chunkBytes);
			o.set(JAVA_INT, 4, blockBytes);
			o.set(JAVA_INT, 8, 0);       // mode auto
			o.set(JAVA_INT, 12, 0);      // search default
			o.set(JAVA_INT, 16, 0);      // depth default
			o.set(JAVA_INT, 20, -1);     // lazy default
			o.set(JAVA_INT, 24, last ? 1 : 0);
			o.set(JAVA_INT, 28, 0);      // chunked framing
			o.set(JAVA_INT, 32, 0);      // CRC-32 alongside
			o.set(JAVA_INT, 36, splitMinBytes);   // 0, or BinarySplit-style adaptive blocks (comp/BinarySplit.java)
			MemorySegment crc = a.allocate(JAVA_INT);
			crc.set(JAVA_INT, 0, crc32);
			long n = B2Deflate.deflateChunks(stage, fill, o, comp, comp.byteSize(), crc, MemorySegment.NULL);
			if (n < 0) throw new IOException("b2d_deflate_chunks: " + B2Deflate.strerror((int)n) + " [" + B2Deflate.lastError() + "]");
			crc32 = crc.get(JAVA_INT, 0);
			byte[] buf = new byte[1 << 20];
			for (long p = 0; p < n; p += buf.length) {
				int k = (int)Math.min(buf.length, n - p);
				MemorySegment.copy(comp, JAVA_BYTE, p, buf, 0, k);
				output.write(buf, 0, k);
			}
		}
		totalIn += fill;
		fill = 0;
	}
}
