/*
 * Panama FFM (java.lang.foreign, Java >= 22) binding of libb2deflate.so -- include/b2deflate.h.
 * No JNI glue: every entry point is one downcall handle.  UNCOMPILED IN THIS REPOSITORY'S IMAGE (no JDK there);
 * the same ABI is exercised by the ctypes binding (binding.py) and the C++ host mirror (host/b2d_streams.hpp).
 */
package io.nayuki.deflate.gpu;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemoryLayout;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.StructLayout;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;
import java.nio.file.Path;

public final class B2Deflate {

	/* b2d_deflate_opts (b2deflate.h) */
	public static final StructLayout DEFLATE_OPTS = MemoryLayout.structLayout(
		JAVA_INT.withName("chunk_bytes"), JAVA_INT.withName("block_bytes"), JAVA_INT.withName("mode"),
		JAVA_INT.withName("search"), JAVA_INT.withName("chain_depth"), JAVA_INT.withName("lazy"),
		JAVA_INT.withName("is_last"), JAVA_INT.withName("framing"), JAVA_INT.withName("checksum"),
		JAVA_INT.withName("split_min_bytes"));

	public static final int INFLATE_CRC32 = 1, INFLATE_CHUNK_INDEXED = 2;
	public static final int ERR_OUTPUT_OVERFLOW = -1;

	private static final Linker LINKER = Linker.nativeLinker();
	private static final SymbolLookup LIB = lookup();

	/* null when the library is not there: available() then says no instead of the class failing to initialise */
	private static SymbolLookup lookup() {
		try {
			return SymbolLookup.libraryLookup(Path.of(System.getProperty("b2deflate.library", "libb2deflate.so")), Arena.global());
		} catch (IllegalArgumentException e) {
			return null;
		}
	}

	private static MethodHandle h(String name, FunctionDescriptor fd) {
		if (LIB == null) return null;
		return LINKER.downcallHandle(LIB.find(name).orElseThrow(() -> new UnsatisfiedLinkError(name)), fd);
	}

	private static final MethodHandle INIT = h("b2d_init", FunctionDescriptor.of(JAVA_INT, JAVA_INT));
	private static final MethodHandle INIT_DEVICES = h("b2d_init_devices", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
	private static final MethodHandle INFLATE_STREAM = h("b2d_inflate_stream", FunctionDescriptor.of(JAVA_INT,
		ADDRESS, JAVA_LONG, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, ADDRESS));
	private static final MethodHandle STRERROR = h("b2d_strerror", FunctionDescriptor.of(ADDRESS, JAVA_INT));
	private static final MethodHandle LAST_ERROR = h("b2d_last_error", FunctionDescriptor.of(ADDRESS));
	private static final MethodHandle ALLOC_PINNED = h("b2d_alloc_pinned", FunctionDescriptor.of(ADDRESS, JAVA_LONG));
	private static final MethodHandle FREE_PINNED = h("b2d_free_pinned", FunctionDescriptor.ofVoid(ADDRESS));
	private static final MethodHandle INFLATE_BATCH = h("b2d_inflate_batch", FunctionDescriptor.of(JAVA_INT,
		ADDRESS, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_INT));
	private static final MethodHandle DEFLATE_BOUND = h("b2d_deflate_bound", FunctionDescriptor.of(JAVA_LONG, JAVA_LONG, JAVA_INT));
	private static final MethodHandle DEFLATE_CHUNKS = h("b2d_deflate_chunks", FunctionDescriptor.of(JAVA_LONG,
		ADDRESS, JAVA_LONG, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS));
	private static final MethodHandle CRC32_COMBINE = h("b2d_crc32_combine", FunctionDescriptor.of(JAVA_INT, JAVA_INT, JAVA_INT, JAVA_LONG));

	private static volatile boolean ready;

	/** Binds the process to one GPU -- or, with -Db2deflate.devices=all, to every GPU of the box (the host entry points
	 *  then shard their units over them) -- IOException-worthy if there is none (there is no CPU fallback). */
	public static void requireDevice() {
		if (ready) return;
		synchronized (B2Deflate.class) {
			if (ready) return;
			if (LIB == null) throw new IllegalStateException("libb2deflate.so not found (-Db2deflate.library=<path>)");
			final int rc;
			if ("all".equals(System.getProperty("b2deflate.devices"))) {
				rc = call(() -> (int)INIT_DEVICES.invokeExact(MemorySegment.NULL, 0));
			} else {
				// (invokeExact is signature-polymorphic: the argument has to be a primitive int, not the Integer that
				// Integer.getInteger returns)
				final int device = Integer.getInteger("b2deflate.device", 0).intValue();
				rc = call(() -> (int)INIT.invokeExact(device));
			}
			if (rc != 0) throw new IllegalStateException(strerror(rc) + " [" + lastError() + "]");
			ready = true;
		}
	}

	/** What the patched constructors of InflaterInputStream / DeflaterOutputStream ask (INTEGRATION.md sections 2 and 3):
	 *  true when the library loads and a GPU binds; the reference's own CPU classes stay in charge otherwise. */
	public static boolean available() {
		try {
			requireDevice();
			return true;
		} catch (IllegalStateException e) {
			return false;
		}
	}

	/** One raw-DEFLATE stream of any origin (b2d_inflate_stream): speculative parallel decode, sequential decoder behind it. */
	public static int inflateStream(MemorySegment in, long inLen, MemorySegment out, long outCap, MemorySegment outLen,
			MemorySegment inConsumed, MemorySegment crc32, MemorySegment status, int flags) {
		return call(() -> (int)INFLATE_STREAM.invokeExact(in, inLen, out, outCap, outLen, inConsumed, crc32, status, flags, MemorySegment.NULL));
	}

	public static String strerror(int status) {
		return call(() -> ((MemorySegment)STRERROR.invokeExact(status)).reinterpret(256).getString(0));
	}
	public static String lastError() {
		return call(() -> ((MemorySegment)LAST_ERROR.invokeExact()).reinterpret(256).getString(0));
	}

	/** Pinned staging memory (b2d_alloc_pinned); free with freePinned. */
	public static MemorySegment allocPinned(long bytes) {
		requireDevice();
		MemorySegment p = call(() -> (MemorySegment)ALLOC_PINNED.invokeExact(bytes));
		if (p.equals(MemorySegment.NULL)) throw new OutOfMemoryError("b2d_alloc_pinned(" + bytes + ")");
		return p.reinterpret(bytes);
	}
	public static void freePinned(MemorySegment p) { call(() -> { FREE_PINNED.invokeExact(p); return 0; }); }

	public static int inflateBatch(MemorySegment in, MemorySegment inOff, int n, MemorySegment out, MemorySegment outOff,
			MemorySegment outLen, MemorySegment inConsumed, MemorySegment crc32, MemorySegment status, int flags) {
		return call(() -> (int)INFLATE_BATCH.invokeExact(in, inOff, n, out, outOff, outLen, inConsumed, crc32, status, flags));
	}
	public static long deflateBound(long inLen, int chunkBytes) {
		return call(() -> (long)DEFLATE_BOUND.invokeExact(inLen, chunkBytes));
	}
	public static long deflateChunks(MemorySegment in, long inLen, MemorySegment opts, MemorySegment out, long outCap,
			MemorySegment crcInOut, MemorySegment chunkOutLen) {
		return call(() -> (long)DEFLATE_CHUNKS.invokeExact(in, inLen, opts, out, outCap, crcInOut, chunkOutLen));
	}
	public static int crc32Combine(int a, int b, long lenB) {
		return call(() -> (int)CRC32_COMBINE.invokeExact(a, b, lenB));
	}

	private interface Downcall<T> { T run() throws Throwable; }
	private static <T> T call(Downcall<T> d) {
		try { return d.run(); }
		catch (RuntimeException | Error e) { throw e; }
		catch (Throwable t) { throw new AssertionError(t); }
	}

	private B2Deflate() {}
}
