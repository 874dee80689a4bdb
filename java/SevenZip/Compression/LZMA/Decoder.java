package SevenZip.Compression.LZMA;

import java.io.IOException;
import java.io.InputStream;
import java.io.OutputStream;
import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;

import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

/**
 * Drop-in for the reference's SevenZip.Compression.LZMA.Decoder (SetDecoderProperties, Code)
 * over liblzma_b200.so.  Code() returns false exactly where the Java loops would.
 */
public class Decoder implements AutoCloseable {
    private MemorySegment handle;

    public Decoder() {
        this(Integer.getInteger("lzma.b200.device", 0));
    }

    public Decoder(int device) {
        try {
            handle = (MemorySegment) LzmaB200.DEC_CREATE.invokeExact(device);
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
        if (handle.equals(MemorySegment.NULL)) {
            throw new RuntimeException("lzb_dec_create: " + LzmaB200.lastError());  // no CPU fallback
        }
    }

    public boolean SetDecoderProperties(byte... properties) {
        try (Arena arena = Arena.ofConfined()) {
            final MemorySegment p = arena.allocate(Math.max(properties.length, 1));
            MemorySegment.copy(properties, 0, p, JAVA_BYTE, 0, properties.length);
            final int rc = (int) LzmaB200.DEC_SET_PROPS.invokeExact(handle, p, properties.length);
            if (rc < 0) {
                throw new RuntimeException("lzma_b200: " + LzmaB200.lastError());
            }
            return rc == LzmaB200.OK;
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }

    /**
     * outSize < 0 decodes until the end marker.  An unknown size has no a-priori bound, so the output
     * buffer grows geometrically and the stream is decoded again when it did not fit.
     */
    public boolean Code(InputStream inStream, OutputStream outStream, long outSize) throws IOException {
        final byte[] data = inStream.readAllBytes();
        long cap = outSize >= 0 ? outSize + 273 : Math.max(1 << 20, 8L * data.length);
        MemorySegment in = null;
        try (Arena arena = Arena.ofConfined()) {
            in = LzmaB200.pinned(data.length);
            MemorySegment.copy(data, 0, in, JAVA_BYTE, 0, data.length);
            final MemorySegment written = arena.allocate(JAVA_LONG);
            while (true) {
                final MemorySegment out = LzmaB200.pinned(cap);
                try {
                    final int rc = (int) LzmaB200.DEC_CODE.invokeExact(handle, in, (long) data.length, out, cap, outSize, written);
                    if (rc == -4 && outSize < 0) {  // LZB_E_CAPACITY
                        cap *= 4;
                        continue;
                    }
                    if (rc < 0) {
                        throw new IOException("lzb_dec_code (" + rc + "): " + LzmaB200.lastError());
                    }
                    if (rc == 0) {
                        return false;  // the reference returns without flushing its window (Decoder.java:281,290)
                    }
                    final long n = written.get(JAVA_LONG, 0);
                    final byte[] chunk = new byte[1 << 20];
                    for (long off = 0; off < n; off += chunk.length) {
                        final int len = (int) Math.min(chunk.length, n - off);
                        MemorySegment.copy(out, JAVA_BYTE, off, chunk, 0, len);
                        outStream.write(chunk, 0, len);
                    }
                    return true;
                } finally {
                    LzmaB200.free(out);
                }
            }
        } catch (IOException e) {
            throw e;
        } catch (Throwable t) {
            throw new IOException(t);
        } finally {
            if (in != null) {
                LzmaB200.free(in);
            }
        }
    }

    @Override
    public void close() {
        if (handle != null && !handle.equals(MemorySegment.NULL)) {
            try {
                LzmaB200.DEC_DESTROY.invokeExact(handle);
            } catch (Throwable ignored) {
            }
            handle = MemorySegment.NULL;
        }
    }
}
