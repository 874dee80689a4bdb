/*
 * lzma_b200.h -- C ABI of the B200-native LZMA block codec (liblzma_b200.so).
 *
 * This is the drop-in boundary for the LZMA Encoder / Decoder surface of
 * rfalke/lzma-java.  Every entry point names the reference interface it
 * replaces; citations are relative to
 *   /root/reference/src/main/java/SevenZip/Compression/LZMA/
 * A Java (Panama FFM or JNI), ctypes or cgo binding needs nothing but this
 * file: plain pointers and sizes, no CUDA or torch types.  INTEGRATION.md
 * shows the reference-side binding.
 *
 * Return convention (all int-returning calls):
 *     1  the reference would have returned true / completed normally
 *     0  the reference would have returned false (rejected setter value,
 *        corrupt stream) -- state unchanged for setters
 *   < 0  infrastructure error (LZB_E_*): no CUDA device, allocation failure,
 *        bad argument, output capacity too small.  A Java binding maps these
 *        to IOException.  lzb_last_error() describes the last one on the
 *        calling thread.
 *
 * There is NO CPU fallback: without a usable sm_100 device every *_create
 * fails and every code call returns LZB_E_CUDA.
 *
 * Threading: a handle owns one CUDA stream and its device scratch and is not
 * thread-safe (like the reference's instances); distinct handles are
 * independent.
 */
#ifndef LZMA_B200_H
#define LZMA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LZB_OK            1
#define LZB_FALSE         0
#define LZB_E_CUDA      (-1)   /* CUDA runtime / no device                */
#define LZB_E_ARG       (-2)   /* NULL handle, bad sizes                  */
#define LZB_E_NOMEM     (-3)   /* host or device allocation failed        */
#define LZB_E_CAPACITY  (-4)   /* an output buffer was too small          */
#define LZB_E_UNSUPPORTED (-5) /* valid for the reference, not for the kernels yet */

#define LZB_HEADER_SIZE  13    /* LzmaAlone framing: 5 props + LE64 size (LzmaAlone.java:208-217) */

typedef struct lzb_enc lzb_enc;
typedef struct lzb_dec lzb_dec;

/* ---- library ---------------------------------------------------------- */

/* "lzma_b200 <version> sm_100a" */
const char *lzb_version(void);
/* Number of usable CUDA devices (0 if none / no driver). */
int lzb_device_count(void);
/* Description of the last error raised on this thread ("" if none). */
const char *lzb_last_error(void);
/* Total kernels launched by this library in this process (bench.py's
 * gpu_launches is a difference of two readings). */
uint64_t lzb_kernel_launches(void);
/* Pinned host memory for the caller's staging buffers (a Java binding drains
 * its InputStream straight into this). */
void *lzb_host_alloc(size_t bytes);
void lzb_host_free(void *p);

/* ---- Encoder (Encoder.java) -------------------------------------------- */

/* new Encoder() (Encoder.java:207-214) bound to CUDA device `device`.
 * Class defaults: dict 1<<22, fb 32, lc3 lp0 pb2, bt4, no end marker
 * (Encoder.java:26-27,151-158,172).  NULL on failure. */
lzb_enc *lzb_enc_create(int device);
void lzb_enc_destroy(lzb_enc *e);

/* Encoder.SetDictionarySize (Encoder.java:1135-1146): [1, 1<<29]. */
int lzb_enc_set_dictionary_size(lzb_enc *e, int32_t dictionary_size);
/* Encoder.SetNumFastBytes (Encoder.java:1148-1154): [5, 273]. */
int lzb_enc_set_num_fast_bytes(lzb_enc *e, int32_t num_fast_bytes);
/* Encoder.SetMatchFinder (Encoder.java:1156-1167): 0 = bt2, 1 = bt4, 2 = bt4 ("bt4b"). */
int lzb_enc_set_match_finder(lzb_enc *e, int32_t match_finder_index);
/* Encoder.SetLcLpPb (Encoder.java:1169-1180): lc in [0,8], lp in [0,4], pb in [0,4]. */
int lzb_enc_set_lc_lp_pb(lzb_enc *e, int32_t lc, int32_t lp, int32_t pb);
/* Encoder.SetEndMarkerMode (Encoder.java:1182-1184). */
int lzb_enc_set_end_marker_mode(lzb_enc *e, int32_t end_marker_mode);
/* Encoder.SetAlgorithm (Encoder.java:1127-1133): accepted and ignored. */
int lzb_enc_set_algorithm(int32_t algorithm);
/* Encoder.WriteCoderProperties (Encoder.java:1079-1085): exactly 5 bytes. */
int lzb_enc_write_coder_properties(const lzb_enc *e, uint8_t out[5]);

/* Capacity that always suffices for the payload of in_len bytes
 * (in_len + in_len/3 + 128; add LZB_HEADER_SIZE when with_header13). */
uint64_t lzb_enc_bound(uint64_t in_len);

/* The ICodeProgress argument of Encoder.Code (ICodeProgress.java:3-5,
 * Encoder.java:929-933, 1070-1072: SetProgress(processedInSize,
 * processedOutSize) after every CodeOneBlock, i.e. every >= 4096 input
 * bytes).  While a later lzb_enc_code* call of this handle runs, `fn` is
 * called from the calling thread with the call's running totals (input bytes
 * consumed, output bytes produced incl. the range coder's pending bytes, as
 * RangeEncoder.getProcessedSizeAdd), monotonically, about once per 64 KiB
 * consumed by a stream while the parser runs (the match finder's phase
 * reports nothing: it produces no output).  fn = NULL switches it off. */
typedef void (*lzb_progress_fn)(void *user, uint64_t in_size, uint64_t out_size);
int lzb_enc_set_progress(lzb_enc *e, lzb_progress_fn fn, void *user);

/* Encoder.Code (Encoder.java:1064-1077) for one stream held in host memory:
 * `in` is what the reference would have drained from its InputStream, `out`
 * receives what it would have written to its OutputStream (payload only, no
 * header).  *out_len = payload bytes. */
int lzb_enc_code(lzb_enc *e, const uint8_t *in, uint64_t in_len,
                 uint8_t *out, uint64_t out_cap, uint64_t *out_len);

/* n independent Encoder.Code calls with this handle's properties, one stream
 * per block, in one batch of kernels.  Block i reads in[in_off[i] ..
 * +in_len[i]) and writes at out[out_off[i] ..) at most out_cap[i] bytes;
 * out_len[i] = bytes written.  with_header13 != 0 prepends the LzmaAlone
 * header (LzmaAlone.java:208-217) so each block is a standalone .lzma file.
 * All pointers are HOST memory (pinned memory from lzb_host_alloc avoids a
 * staging copy). */
int lzb_enc_code_batch(lzb_enc *e, const uint8_t *in, const uint64_t *in_off,
                       const uint64_t *in_len, uint32_t n,
                       uint8_t *out, const uint64_t *out_off,
                       const uint64_t *out_cap, uint64_t *out_len,
                       int32_t with_header13);

/* Same batch with every pointer in DEVICE memory of the handle's device;
 * work is enqueued on `cuda_stream` (a cudaStream_t passed as void*, NULL =
 * the handle's own stream).  The call is BLOCKING: between the match finder
 * and the parser the host reads back the per-block list sizes (to place the
 * lists and to order the blocks), so it synchronises `cuda_stream` several
 * times and returns when the batch is done; out_len[i] is valid on return.
 * It may allocate device memory (grow-only scratch of the handle) and cannot
 * be captured into a CUDA graph.  A handle serves one call at a time.  A
 * block whose capacity was too small reports out_len[i] = UINT64_MAX.
 * max_in_len >= every in_len[i] sizes the scratch; a longer block makes the
 * call fail with LZB_E_ARG instead of touching memory it was not given. */
int lzb_enc_code_batch_device(lzb_enc *e, const uint8_t *d_in,
                              const uint64_t *d_in_off, const uint64_t *d_in_len,
                              uint32_t n, uint64_t max_in_len,
                              uint8_t *d_out, const uint64_t *d_out_off,
                              const uint64_t *d_out_cap, uint64_t *d_out_len,
                              int32_t with_header13, void *cuda_stream);

/* Trace tap of the match finder (BinTree.fillMatches0 / Skip, BinTree.java:152-356; the
 * machine-readable twin of the reference's FINE log, BinTree.java:139-150): for every position p
 * of one block, counts[p] = number of (length, distance) pairs the reference's match finder
 * yields there -- Skip()ped positions included -- and the pairs themselves, concatenated in
 * position order as pairs[2k] = length, pairs[2k+1] = distance (= back - 1).  *pairs_used =
 * total number of pairs.  HOST pointers.  For parity tests and debugging, not a hot path. */
int lzb_enc_trace_matches(lzb_enc *e, const uint8_t *in, uint64_t in_len,
                          uint32_t *counts, uint32_t *pairs, uint64_t pairs_cap,
                          uint64_t *pairs_used);

/* ---- Decoder (Decoder.java) -------------------------------------------- */

/* new Decoder() (Decoder.java:154-158) bound to CUDA device `device`. */
lzb_dec *lzb_dec_create(int device);
void lzb_dec_destroy(lzb_dec *d);

/* Decoder.SetDecoderProperties (Decoder.java:303-318): n_props >= 5;
 * lc = v%9, lp = (v/9)%5, pb = v/45, dict = LE32. */
int lzb_dec_set_decoder_properties(lzb_dec *d, const uint8_t *props, uint32_t n_props);

/* Decoder.Code (Decoder.java:205-301) for one payload in host memory, using
 * the properties set above.  out_size < 0: decode until the end marker.
 * Returns 1 / 0 exactly where the reference returns true / false.  Like the
 * reference (Decoder.java:292-293) the last match is not clamped to
 * out_size, so out_cap should allow out_size + 272.  *written = bytes
 * produced. */
int lzb_dec_code(lzb_dec *d, const uint8_t *in, uint64_t in_len,
                 uint8_t *out, uint64_t out_cap, int64_t out_size,
                 uint64_t *written);

/* n independent LzmaAlone decodes (LzmaAlone.java:220-239): stream i is the
 * .lzma file in[in_off[i] .. +in_len[i]) whose own 13-byte header supplies
 * properties and size.  status[i] = 1 / 0 (reference true / false, or a
 * header the reference rejects), LZB_E_CAPACITY, or LZB_E_UNSUPPORTED for a
 * stream of 4 GiB or more (positions are 32-bit in the kernel); out_len[i] = bytes
 * written at out[out_off[i] ..).  HOST pointers.  The call returns 1 if it
 * ran (inspect status[] per stream), < 0 on infrastructure errors. */
int lzb_dec_code_batch(lzb_dec *d, const uint8_t *in, const uint64_t *in_off,
                       const uint64_t *in_len, uint32_t n,
                       uint8_t *out, const uint64_t *out_off,
                       const uint64_t *out_cap, uint64_t *out_len,
                       int32_t *status);

/* Same batch with every pointer in DEVICE memory; enqueued on `cuda_stream`
 * (NULL = the handle's own stream).  The call synchronises the stream once
 * (it scans the n stream headers on the device to pick the kernel variant and
 * size the model scratch) and returns after the decode kernel has been
 * enqueued: d_out, d_out_len and d_status are valid once `cuda_stream` has
 * been synchronised.  A handle serves one call at a time (the calls share
 * its ticket counter and scratch); not capturable into a CUDA graph. */
int lzb_dec_code_batch_device(lzb_dec *d, const uint8_t *d_in,
                              const uint64_t *d_in_off, const uint64_t *d_in_len,
                              uint32_t n, uint8_t *d_out,
                              const uint64_t *d_out_off, const uint64_t *d_out_cap,
                              uint64_t *d_out_len, int32_t *d_status,
                              void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* LZMA_B200_H */
