package SevenZip.Compression.LZMA;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

/**
 * Panama FFM (JDK 22+) bindings of include/lzma_b200.h.  One downcall handle per C entry point;
 * no JNI glue and no generated code.  The library path comes from -Dlzma.b200.lib=... or
 * java.library.path ("lzma_b200").
 *
 * NOTE: this source ships for the reference's maintainers; the build image has no JDK, so it
 * has been neither compiled nor run here (see INTEGRATION.md).
 */
final class LzmaB200 {
    private LzmaB200() {
    }

    static final int OK = 1;

    static final Linker LINKER = Linker.nativeLinker();
    private static final SymbolLookup LIB = open();

    private static SymbolLookup open() {
        final String path = System.getProperty("lzma.b200.lib");
        if (path != null) {
            return SymbolLookup.libraryLookup(path, Arena.global());
        }
        System.loadLibrary("lzma_b200");
        return SymbolLookup.loaderLookup();
    }

    private static MethodHandle fn(String name, FunctionDescriptor fd) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(() -> new UnsatisfiedLinkError(name)), fd);
    }

    static final MethodHandle LAST_ERROR = fn("lzb_last_error", FunctionDescriptor.of(ADDRESS));
    static final MethodHandle HOST_ALLOC = fn("lzb_host_alloc", FunctionDescriptor.of(ADDRESS, JAVA_LONG));
    static final MethodHandle HOST_FREE = fn("lzb_host_free", FunctionDescriptor.ofVoid(ADDRESS));
    static final MethodHandle ENC_BOUND = fn("lzb_enc_bound", FunctionDescriptor.of(JAVA_LONG, JAVA_LONG));

    static final MethodHandle ENC_CREATE = fn("lzb_enc_create", FunctionDescriptor.of(ADDRESS, JAVA_INT));
    static final MethodHandle ENC_DESTROY = fn("lzb_enc_destroy", FunctionDescriptor.ofVoid(ADDRESS));
    static final MethodHandle ENC_SET_DICT = fn("lzb_enc_set_dictionary_size", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
    static final MethodHandle ENC_SET_FB = fn("lzb_enc_set_num_fast_bytes", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
    static final MethodHandle ENC_SET_MF = fn("lzb_enc_set_match_finder", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
    static final MethodHandle ENC_SET_LCLPPB = fn("lzb_enc_set_lc_lp_pb", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT));
    static final MethodHandle ENC_SET_EOS = fn("lzb_enc_set_end_marker_mode", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
    static final MethodHandle ENC_PROPS = fn("lzb_enc_write_coder_properties", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle ENC_SET_PROGRESS = fn("lzb_enc_set_progress", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    /** lzb_progress_fn: void (*)(void *user, uint64_t in_size, uint64_t out_size) */
    static final FunctionDescriptor PROGRESS_FD = FunctionDescriptor.ofVoid(ADDRESS, JAVA_LONG, JAVA_LONG);
    static final MethodHandle ENC_CODE = fn("lzb_enc_code",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS, JAVA_LONG, ADDRESS));

    static final MethodHandle DEC_CREATE = fn("lzb_dec_create", FunctionDescriptor.of(ADDRESS, JAVA_INT));
    static final MethodHandle DEC_DESTROY = fn("lzb_dec_destroy", FunctionDescriptor.ofVoid(ADDRESS));
    static final MethodHandle DEC_SET_PROPS = fn("lzb_dec_set_decoder_properties", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT));
    static final MethodHandle DEC_CODE = fn("lzb_dec_code",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS, JAVA_LONG, JAVA_LONG, ADDRESS));

    static String lastError() {
        try {
            final MemorySegment p = (MemorySegment) LAST_ERROR.invokeExact();
            return p.reinterpret(512).getString(0);
        } catch (Throwable t) {
            return t.toString();
        }
    }

    /** Pinned host memory owned by the CUDA runtime (lzb_host_alloc); H2D/D2H run at full PCIe speed from it. */
    static MemorySegment pinned(long bytes) throws java.io.IOException {
        try {
            final MemorySegment p = (MemorySegment) HOST_ALLOC.invokeExact(bytes);
            if (p.equals(MemorySegment.NULL)) {
                throw new java.io.IOException("lzb_host_alloc: " + lastError());
            }
            return p.reinterpret(Math.max(bytes, 1));
        } catch (java.io.IOException e) {
            throw e;
        } catch (Throwable t) {
            throw new java.io.IOException(t);
        }
    }

    static void free(MemorySegment p) {
        try {
            HOST_FREE.invokeExact(p);
        } catch (Throwable ignored) {
        }
    }
}
