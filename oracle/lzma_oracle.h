/*
 * lzma_oracle.h -- CPU restatement of rfalke/lzma-java (LZMA SDK Java 4.61).
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle: a plain-C restatement
 * of the reference's algorithm.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product
 * (lzma-java_b200/csrc, liblzma_b200.so) never links, loads or calls it.
 *
 * Pinning: the 12 (length, md5) golden vectors of LzmaAloneTest.java:27-38,
 * the range-encoder byte strings of RangeCoder/EncoderLearningTest.java:31-72
 * and the bit-tree prices of RangeCoder/BitTreeEncoderLearningTest.java:24-31
 * are all reproduced (tests/test_oracle_golden.py).
 *
 * File:line citations below are relative to
 * /root/reference/src/main/java/SevenZip/.
 */
#ifndef LZMA_ORACLE_H
#define LZMA_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Encoder configuration == the reference's setters (Encoder.java:1127-1184). */
typedef struct {
    int32_t dict_size;    /* SetDictionarySize: [1, 1<<29]            */
    int32_t lc, lp, pb;   /* SetLcLpPb: lc<=8, lp<=4, pb<=4            */
    int32_t fb;           /* SetNumFastBytes: [5, 273]                 */
    int32_t mf;           /* SetMatchFinder: 0 = bt2, 1 or 2 = bt4      */
    int32_t eos;          /* SetEndMarkerMode                          */
} lzo_props;

/* Optional trace taps (the machine-readable twin of the reference's FINE
 * log, BinTree.java:139-150 and Encoder.java:891-897).  Any pointer may be
 * NULL.  Match lists are recorded for EVERY position that reaches the match
 * finder, Skip()ped ones included. */
typedef struct {
    /* per-position match lists: for position p (0-based) pairs
     * mf_pairs[mf_off[p] .. mf_off[p+1]) as (length, distance) */
    uint32_t *mf_off;      /* n+1 entries, caller allocated              */
    uint32_t *mf_pairs;    /* 2 * mf_pairs_cap u32, caller allocated     */
    uint64_t  mf_pairs_cap;
    uint64_t  mf_pairs_used;  /* out: number of pairs recorded           */
    uint64_t  mf_overflow;    /* out: pairs dropped for lack of capacity */
    /* per-decision list: (offset, back, len); back = -1 literal, 0..3 rep,
     * else distance + 4 (Encoder.PosAndLength, Encoder.java:43-84) */
    int64_t  *dec;         /* 3 * dec_cap int64, caller allocated        */
    uint64_t  dec_cap;
    uint64_t  dec_used;    /* out */
} lzo_trace;

/* Validate props the way the setters do; 1 = accepted, 0 = a setter would
 * have returned false. */
int lzo_props_valid(const lzo_props *p);

/* Encoder.WriteCoderProperties (Encoder.java:1079-1085). */
void lzo_write_props(const lzo_props *p, uint8_t out[5]);

/* Upper bound we use for the payload of n input bytes. */
size_t lzo_encode_bound(size_t n);

/* Encoder.Code (Encoder.java:1064-1077): payload only, no header.
 * Returns payload length, or (size_t)-1 on bad props / allocation failure /
 * insufficient capacity. */
size_t lzo_encode(const lzo_props *p, const uint8_t *in, size_t n,
                  uint8_t *out, size_t out_cap, lzo_trace *trace);

/* LzmaAlone framing (LzmaAlone.java:208-217): 5 props + LE64 size (or -1
 * with eos) + payload.  Returns total length or (size_t)-1. */
size_t lzo_encode_alone(const lzo_props *p, const uint8_t *in, size_t n,
                        uint8_t *out, size_t out_cap);

/* Decoder.SetDecoderProperties + Decoder.Code (Decoder.java:205-318).
 * props: 5 bytes.  out_size < 0: decode until the end marker.
 * Returns 1 (true), 0 (reference returns false: bad props / corrupt data),
 * -1 if out_cap was too small to hold what the reference would have written.
 * *written = bytes produced (on 0 the reference would not have flushed its
 * window; we still report how far decoding got).
 * *consumed (may be NULL) = input bytes read by the range decoder. */
int lzo_decode(const uint8_t props[5], const uint8_t *in, size_t in_len,
               uint8_t *out, size_t out_cap, int64_t out_size,
               size_t *written, size_t *consumed);

/* LzmaAlone decode (LzmaAlone.java:220-239): parses the 13-byte header. */
int lzo_decode_alone(const uint8_t *in, size_t in_len,
                     uint8_t *out, size_t out_cap, size_t *written);

/* ---- known-answer helpers (tests only) ---- */

/* RangeEncoder on ONE adaptive prob (index 4 of a 12-entry model), as in
 * RangeCoder/EncoderLearningTest.java:86-96.  Returns bytes written. */
size_t lzo_kat_rc_bits(const int *bits, int nbits, uint8_t *out, size_t cap);
/* encodeDirectBits calls (v[i], nbits[i]) then flush,
 * EncoderLearningTest.java:55-68. */
size_t lzo_kat_rc_direct(const int *v, const int *nbits, int ncalls,
                         uint8_t *out, size_t cap);
/* BitTreeEncoder(3).encode(3) then getPrice(0..7),
 * BitTreeEncoderLearningTest.java:14-31. */
void lzo_kat_bittree_prices(int prices[8]);
/* ProbPrices table (512 ints, ProbPrices.java:8-18). */
void lzo_kat_prob_prices(int table[512]);

/* ---- multi-threaded batch drivers (cpu_baseline only) ---- */

/* Encode n_blocks blocks with `threads` pthreads, one block per task
 * (BASELINE.md section 3).  in_off/in_len/out_off in bytes; out_len filled.
 * with_header != 0 writes the 13-byte LzmaAlone header before the payload.
 * Returns 0 on success. */
int lzo_encode_batch(const lzo_props *p, const uint8_t *in,
                     const uint64_t *in_off, const uint64_t *in_len,
                     uint32_t n_blocks, uint8_t *out, const uint64_t *out_off,
                     const uint64_t *out_cap, uint64_t *out_len,
                     int with_header, int threads);

/* Decode n_blocks LzmaAlone streams with `threads` pthreads.
 * status[i] receives lzo_decode_alone's return value. */
int lzo_decode_batch(const uint8_t *in, const uint64_t *in_off,
                     const uint64_t *in_len, uint32_t n_blocks, uint8_t *out,
                     const uint64_t *out_off, const uint64_t *out_cap,
                     uint64_t *out_len, int32_t *status, int threads);

#ifdef __cplusplus
}
#endif
#endif
