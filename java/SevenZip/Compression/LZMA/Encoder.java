package SevenZip.Compression.LZMA;

import SevenZip.ICodeProgress;

import java.io.IOException;
import java.io.InputStream;
import java.io.OutputStream;
import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;

import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

/**
 * Drop-in for the reference's SevenZip.Compression.LZMA.Encoder: identical public surface
 * (Code, WriteCoderProperties, SetDictionarySize, SetNumFastBytes, SetMatchFinder, SetLcLpPb,
 * SetEndMarkerMode, SetAlgorithm), implemented over liblzma_b200.so (include/lzma_b200.h).
 * Code() drains the InputStream into pinned memory, runs the CUDA encoder and writes the
 * payload -- bit-identical to what the Java loops would have written.
 */
public class Encoder implements AutoCloseable {
    private MemorySegment handle;

    public Encoder() {
        this(Integer.getInteger("lzma.b200.device", 0));
    }

    public Encoder(int device) {
        try {
            handle = (MemorySegment) LzmaB200.ENC_CREATE.invokeExact(device);
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
        if (handle.equals(MemorySegment.NULL)) {
            throw new RuntimeException("lzb_enc_create: " + LzmaB200.lastError());  // no CPU fallback
        }
    }

    private static boolean ok(int rc) {
        if (rc < 0) {
            throw new RuntimeException("lzma_b200: " + LzmaB200.lastError());
        }
        return rc == LzmaB200.OK;
    }

    public static boolean SetAlgorithm(int algorithm) {
        return true;
    }

    public boolean SetDictionarySize(int dictionarySize) {
        try {
            return ok((int) LzmaB200.ENC_SET_DICT.invokeExact(handle, dictionarySize));
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }

    public boolean SetNumFastBytes(int numFastBytes) {
        try {
            return ok((int) LzmaB200.ENC_SET_FB.invokeExact(handle, numFastBytes));
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }

    public boolean SetMatchFinder(int matchFinderIndex) {
        try {
            return ok((int) LzmaB200.ENC_SET_MF.invokeExact(handle, matchFinderIndex));
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }

    public boolean SetLcLpPb(int lc, int lp, int pb) {
        try {
            return ok((int) LzmaB200.ENC_SET_LCLPPB.invokeExact(handle, lc, lp, pb));
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }

    public void SetEndMarkerMode(boolean endMarkerMode) {
        try {
            ok((int) LzmaB200.ENC_SET_EOS.invokeExact(handle, endMarkerMode ? 1 : 0));
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }

    public void WriteCoderProperties(OutputStream outStream) throws IOException {
        try (Arena arena = Arena.ofConfined()) {
            final MemorySegment props = arena.allocate(5);
            ok((int) LzmaB200.ENC_PROPS.invokeExact(handle, props));
            outStream.write(props.toArray(JAVA_BYTE), 0, 5);
        } catch (IOException | RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new IOException(t);
        }
    }

    /** inSize / outSize are ignored exactly as in the reference (Encoder.java:1046-1062). */
    public void Code(InputStream inStream, OutputStream outStream, long inSize, long outSize, ICodeProgress progress) throws IOException {
        final byte[] data = inStream.readAllBytes();
        MemorySegment in = null;
        MemorySegment out = null;
        try (Arena arena = Arena.ofConfined()) {
            final long cap = (long) LzmaB200.ENC_BOUND.invokeExact((long) data.length);
            in = LzmaB200.pinned(data.length);
            out = LzmaB200.pinned(cap);
            MemorySegment.copy(data, 0, in, JAVA_BYTE, 0, data.length);
            final MemorySegment outLen = arena.allocate(JAVA_LONG);
            if (progress != null) {
                // ICodeProgress as an upcall: the library calls it from this thread with running totals while the parser
                // runs (Encoder.java:929-933, 1070-1072 call it after every CodeOneBlock)
                final java.lang.invoke.MethodHandle target = java.lang.invoke.MethodHandles.lookup().bind(
                        new ProgressUpcall(progress), "call",
                        java.lang.invoke.MethodType.methodType(void.class, MemorySegment.class, long.class, long.class));
                final MemorySegment stub = LzmaB200.LINKER.upcallStub(target, LzmaB200.PROGRESS_FD, arena);
                ok((int) LzmaB200.ENC_SET_PROGRESS.invokeExact(handle, stub, MemorySegment.NULL));
            }
            final int rc;
            try {
                rc = (int) LzmaB200.ENC_CODE.invokeExact(handle, in, (long) data.length, out, cap, outLen);
            } finally {
                if (progress != null) {
                    ok((int) LzmaB200.ENC_SET_PROGRESS.invokeExact(handle, MemorySegment.NULL, MemorySegment.NULL));
                }
            }
            if (rc != LzmaB200.OK) {
                throw new IOException("lzb_enc_code (" + rc + "): " + LzmaB200.lastError());
            }
            final long n = outLen.get(JAVA_LONG, 0);
            final byte[] chunk = new byte[1 << 20];
            for (long off = 0; off < n; off += chunk.length) {
                final int len = (int) Math.min(chunk.length, n - off);
                MemorySegment.copy(out, JAVA_BYTE, off, chunk, 0, len);
                outStream.write(chunk, 0, len);
            }
            outStream.flush();  // RangeEncoder.flush() flushes the stream (RangeEncoder.java:35)
            if (progress != null) {
                progress.SetProgress(data.length, n);  // LzmaBench needs one call with inSize >= dict (LzmaBench.java:219-223)
            }
        } catch (IOException e) {
            throw e;
        } catch (Throwable t) {
            throw new IOException(t);
        } finally {
            if (in != null) {
                LzmaB200.free(in);
            }
            if (out != null) {
                LzmaB200.free(out);
            }
        }
    }

    /** Receiver of the lzb_progress_fn upcall. */
    private static final class ProgressUpcall {
        private final ICodeProgress progress;

        ProgressUpcall(ICodeProgress progress) {
            this.progress = progress;
        }

        @SuppressWarnings("unused")  // bound by name in Code
        void call(MemorySegment user, long inSize, long outSize) {
            progress.SetProgress(inSize, outSize);
        }
    }

    @Override
    public void close() {
        if (handle != null && !handle.equals(MemorySegment.NULL)) {
            try {
                LzmaB200.ENC_DESTROY.invokeExact(handle);
            } catch (Throwable ignored) {
            }
            handle = MemorySegment.NULL;
        }
    }
}
