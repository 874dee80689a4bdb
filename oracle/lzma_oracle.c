/*
 * lzma_oracle.c -- CPU restatement of rfalke/lzma-java (LZMA SDK Java 4.61).
 *
 * TEST INFRASTRUCTURE ONLY (see lzma_oracle.h).  Nothing under
 * lzma-java_b200/ links or loads this file.
 *
 * The layout follows the reference class by class so that every function can
 * be checked against the Java it restates; citations are relative to
 * /root/reference/src/main/java/SevenZip/.  Stream I/O is replaced by whole
 * buffers, which is output-transparent: while the
 * reference is not at EOF its window always holds >= fb + 274 bytes ahead, so
 * the only place buffering could show (InWindow.GetMatchLen's clamp,
 * InWindow.java:121-125) behaves as if the end were always known.
 *
 * Pinned by tests/test_oracle_golden.py (12 firefox.exe md5/length vectors,
 * range-coder KATs, bit-tree prices).
 */
#include "lzma_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* constants: LZMA/Base.java:5-86, RangeCoder/RangeBase.java:4-7       */
/* ------------------------------------------------------------------ */
enum {
    kNumRepDistances = 4,
    kNumStates = 12,
    kNumPosSlotBits = 6,
    kNumLenToPosStatesBits = 2,
    kNumLenToPosStates = 1 << kNumLenToPosStatesBits,
    kMatchMinLen = 2,
    kNumAlignBits = 4,
    kAlignTableSize = 1 << kNumAlignBits,
    kAlignMask = kAlignTableSize - 1,
    kStartPosModelIndex = 4,
    kEndPosModelIndex = 14,
    kNumFullDistances = 1 << (kEndPosModelIndex / 2),
    kNumLitPosStatesBitsEncodingMax = 4,
    kNumLitContextBitsMax = 8,
    kNumPosStatesBitsMax = 4,
    kNumPosStatesMax = 1 << kNumPosStatesBitsMax,
    kNumPosStatesBitsEncodingMax = 4,
    kNumPosStatesEncodingMax = 1 << kNumPosStatesBitsEncodingMax,
    kNumLowLenBits = 3,
    kNumMidLenBits = 3,
    kNumHighLenBits = 8,
    kNumLowLenSymbols = 1 << kNumLowLenBits,
    kNumMidLenSymbols = 1 << kNumMidLenBits,
    kNumLenSymbols = kNumLowLenSymbols + kNumMidLenSymbols + (1 << kNumHighLenBits),
    kMatchMaxLen = kMatchMinLen + kNumLenSymbols - 1, /* 273 */

    kNumBitModelTotalBits = 11,
    kBitModelTotal = 1 << kNumBitModelTotalBits,
    kNumMoveBits = 5,
    kNumMoveReducingBits = 2,   /* ProbPrices.java:4 */
    kNumBitPriceShiftBits = 6,  /* ProbPrices.java:6 */

    kNumOpts = 1 << 12,         /* Encoder.java:19 */
    kIfinityPrice = 0xFFFFFFF,  /* Encoder.java:22 */

    kHash2Size = 1 << 10,       /* BinTree.java:13-19 */
    kHash3Size = 1 << 16,
    kBT2HashSize = 1 << 16,
    kStartMaxLen = 1,
    kHash3Offset = kHash2Size,
    kEmptyHashValue = 0,
    kMaxValForNormalize = (1 << 30) - 1
};
#define kTopValue (1u << 24)

/* Base.java:16-40 */
static inline int st_lit(int s) { return s < 4 ? 0 : (s < 10 ? s - 3 : s - 6); }
static inline int st_match(int s) { return s < 7 ? 7 : 10; }
static inline int st_shortrep(int s) { return s < 7 ? 9 : 11; }
static inline int st_longrep(int s) { return s < 7 ? 8 : 11; }
static inline int st_is_char(int s) { return s < 7; }
/* Base.java:52-58 */
static inline int len_to_pos_state(int len) {
    len -= kMatchMinLen;
    return len < kNumLenToPosStates ? len : kNumLenToPosStates - 1;
}

/* ------------------------------------------------------------------ */
/* static tables                                                       */
/* ------------------------------------------------------------------ */
static uint32_t g_crc[256];          /* CRC.java:6-20 */
static int32_t g_prob_prices[kBitModelTotal >> kNumMoveReducingBits]; /* ProbPrices.java:5-18 */
static uint8_t g_fast_pos[1 << 11];  /* Encoder.java:24-41 */
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static void init_tables(void) {
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t r = i;
        for (int j = 0; j < 8; j++) r = (r & 1) ? (r >> 1) ^ 0xEDB88320u : r >> 1;
        g_crc[i] = r;
    }
    /* ProbPrices.java:8-18; entry 0 is never written and stays 0 */
    const int kNumBits = kNumBitModelTotalBits - kNumMoveReducingBits;
    memset(g_prob_prices, 0, sizeof g_prob_prices);
    for (int i = kNumBits - 1; i >= 0; i--) {
        int start = 1 << (kNumBits - i - 1);
        int end = 1 << (kNumBits - i);
        for (int j = start; j < end; j++)
            g_prob_prices[j] = (i << kNumBitPriceShiftBits) +
                (int32_t)(((uint32_t)(end - j) << kNumBitPriceShiftBits) >> (kNumBits - i - 1));
    }
    /* Encoder.java:30-41 */
    g_fast_pos[0] = 0;
    g_fast_pos[1] = 1;
    int c = 2;
    for (int slot = 2; slot < 22; slot++) {
        int k = 1 << ((slot >> 1) - 1);
        for (int j = 0; j < k; j++, c++) g_fast_pos[c] = (uint8_t)slot;
    }
}

/* ProbPrices.java:23-37 */
static inline int32_t price_bit(int prob, int bit) {
    return g_prob_prices[(((prob - bit) ^ (-bit)) & (kBitModelTotal - 1)) >> kNumMoveReducingBits];
}
static inline int32_t price0(int prob) { return g_prob_prices[prob >> kNumMoveReducingBits]; }
static inline int32_t price1(int prob) { return g_prob_prices[(kBitModelTotal - prob) >> kNumMoveReducingBits]; }

/* Encoder.java:86-104 */
static inline int get_pos_slot(uint32_t pos) {
    if (pos < (1u << 11)) return g_fast_pos[pos];
    if (pos < (1u << 21)) return g_fast_pos[pos >> 10] + 20;
    return g_fast_pos[pos >> 20] + 40;
}
static inline int get_pos_slot2(uint32_t pos) {
    if (pos < (1u << 17)) return g_fast_pos[pos >> 6] + 12;
    if (pos < (1u << 27)) return g_fast_pos[pos >> 16] + 32;
    return g_fast_pos[pos >> 26] + 52;
}

/* ------------------------------------------------------------------ */
/* RangeCoder/RangeEncoder.java                                        */
/* ------------------------------------------------------------------ */
typedef struct {
    uint64_t low;
    uint32_t range;
    int cache_size;
    int cache;
    uint64_t position;
    uint8_t *out;
    size_t out_pos, out_cap;
    int overflow;
} rc_enc;

static void rc_init(rc_enc *rc, uint8_t *out, size_t cap) { /* :23-29 */
    rc->position = 0;
    rc->low = 0;
    rc->range = 0xFFFFFFFFu;
    rc->cache_size = 1;
    rc->cache = 0;
    rc->out = out;
    rc->out_pos = 0;
    rc->out_cap = cap;
    rc->overflow = 0;
}

static inline void rc_put(rc_enc *rc, int b) {
    if (rc->out_pos < rc->out_cap) rc->out[rc->out_pos] = (uint8_t)b;
    else rc->overflow = 1;
    rc->out_pos++;
}

static void rc_shift_low(rc_enc *rc) { /* :73-87 */
    int low_hi = (int)(rc->low >> 32);
    if (low_hi != 0 || rc->low < 0xFF000000ull) {
        rc->position += (uint64_t)rc->cache_size;
        int temp = rc->cache;
        do {
            rc_put(rc, temp + low_hi);
            temp = 0xFF;
        } while (--rc->cache_size != 0);
        rc->cache = (int)(((uint32_t)rc->low) >> 24);
    }
    rc->cache_size++;
    rc->low = (rc->low & 0xFFFFFF) << 8;
}

static void rc_flush(rc_enc *rc) { /* :31-36 */
    for (int i = 0; i < 5; i++) rc_shift_low(rc);
}

static inline void rc_encode(rc_enc *rc, uint16_t *probs, int index, int symbol) { /* :38-54 */
    uint32_t prob = probs[index];
    uint32_t bound = (rc->range >> kNumBitModelTotalBits) * prob;
    if (symbol == 0) {
        rc->range = bound;
        probs[index] = (uint16_t)(prob + ((kBitModelTotal - prob) >> kNumMoveBits));
    } else {
        rc->low += bound;
        rc->range -= bound;
        probs[index] = (uint16_t)(prob - (prob >> kNumMoveBits));
    }
    if (rc->range < kTopValue) {
        rc->range <<= 8;
        rc_shift_low(rc);
    }
}

static void rc_encode_direct(rc_enc *rc, uint32_t v, int nbits) { /* :56-67 */
    for (int i = nbits - 1; i >= 0; i--) {
        rc->range >>= 1;
        if (((v >> i) & 1) == 1) rc->low += rc->range;
        if (rc->range < kTopValue) {
            rc->range <<= 8;
            rc_shift_low(rc);
        }
    }
}

/* ------------------------------------------------------------------ */
/* RangeCoder/BitTreeEncoder.java (probs: 1 << nbits shorts)           */
/* ------------------------------------------------------------------ */
static void bt_encode(rc_enc *rc, uint16_t *probs, int nbits, int symbol) { /* :18-26 */
    int m = 1;
    for (int bi = nbits; bi != 0;) {
        bi--;
        int bit = (symbol >> bi) & 1;
        rc_encode(rc, probs, m, bit);
        m = (m << 1) | bit;
    }
}
static void bt_reverse_encode(rc_enc *rc, uint16_t *probs, int nbits, int symbol) { /* :28-36, Encoder.java:196-205 */
    int m = 1;
    for (int i = 0; i < nbits; i++) {
        int bit = symbol & 1;
        rc_encode(rc, probs, m, bit);
        m = (m << 1) | bit;
        symbol >>= 1;
    }
}
static int32_t bt_price(const uint16_t *probs, int nbits, int symbol) { /* :38-48 */
    int32_t price = 0;
    int m = 1;
    for (int bi = nbits; bi != 0;) {
        bi--;
        int bit = (symbol >> bi) & 1;
        price += price_bit(probs[m], bit);
        m = (m << 1) + bit;
    }
    return price;
}
static int32_t bt_reverse_price(const uint16_t *probs, int nbits, int symbol) { /* :50-60, Encoder.java:183-194 */
    int32_t price = 0;
    int m = 1;
    for (int i = nbits; i != 0; i--) {
        int bit = symbol & 1;
        symbol >>= 1;
        price += price_bit(probs[m], bit);
        m = (m << 1) | bit;
    }
    return price;
}

static void init_probs(uint16_t *p, size_t n) { /* RangeBase.java:9-13 */
    for (size_t i = 0; i < n; i++) p[i] = kBitModelTotal >> 1;
}

/* ------------------------------------------------------------------ */
/* LZMA/LenEncoder.java + LenPriceTableEncoder.java                    */
/* ------------------------------------------------------------------ */
typedef struct {
    uint16_t choice[2];
    uint16_t low[kNumPosStatesEncodingMax][1 << kNumLowLenBits];
    uint16_t mid[kNumPosStatesEncodingMax][1 << kNumMidLenBits];
    uint16_t high[1 << kNumHighLenBits];
    int32_t prices[kNumLenSymbols << kNumPosStatesBitsEncodingMax];
    int32_t counters[kNumPosStatesEncodingMax];
    int table_size;
} len_enc;

static void len_init(len_enc *le, int num_pos_states) { /* LenEncoder.java:23-31 */
    init_probs(le->choice, 2);
    for (int ps = 0; ps < num_pos_states; ps++) {
        init_probs(le->low[ps], 1 << kNumLowLenBits);
        init_probs(le->mid[ps], 1 << kNumMidLenBits);
    }
    init_probs(le->high, 1 << kNumHighLenBits);
}

static void len_set_prices(len_enc *le, int pos_state, int num_symbols, int32_t *prices, int st) { /* LenEncoder.java:50-71 */
    int32_t a0 = price0(le->choice[0]);
    int32_t a1 = price1(le->choice[0]);
    int32_t b0 = a1 + price0(le->choice[1]);
    int32_t b1 = a1 + price1(le->choice[1]);
    int i;
    for (i = 0; i < kNumLowLenSymbols; i++) {
        if (i >= num_symbols) return;
        prices[st + i] = a0 + bt_price(le->low[pos_state], kNumLowLenBits, i);
    }
    for (; i < kNumLowLenSymbols + kNumMidLenSymbols; i++) {
        if (i >= num_symbols) return;
        prices[st + i] = b0 + bt_price(le->mid[pos_state], kNumMidLenBits, i - kNumLowLenSymbols);
    }
    for (; i < num_symbols; i++)
        prices[st + i] = b1 + bt_price(le->high, kNumHighLenBits, i - kNumLowLenSymbols - kNumMidLenSymbols);
}

static void len_update_table(len_enc *le, int pos_state) { /* LenPriceTableEncoder.java:20-23 */
    len_set_prices(le, pos_state, le->table_size, le->prices, pos_state * kNumLenSymbols);
    le->counters[pos_state] = le->table_size;
}
static void len_update_tables(len_enc *le, int num_pos_states) { /* :25-29 */
    for (int ps = 0; ps < num_pos_states; ps++) len_update_table(le, ps);
}
static inline int32_t len_get_price(const len_enc *le, int symbol, int pos_state) { /* :16-18 */
    return le->prices[pos_state * kNumLenSymbols + symbol];
}
static void len_encode(len_enc *le, rc_enc *rc, int symbol, int pos_state) { /* LenEncoder.java:33-48 + LenPriceTableEncoder.java:32-37 */
    if (symbol < kNumLowLenSymbols) {
        rc_encode(rc, le->choice, 0, 0);
        bt_encode(rc, le->low[pos_state], kNumLowLenBits, symbol);
    } else {
        int s = symbol - kNumLowLenSymbols;
        rc_encode(rc, le->choice, 0, 1);
        if (s < kNumMidLenSymbols) {
            rc_encode(rc, le->choice, 1, 0);
            bt_encode(rc, le->mid[pos_state], kNumMidLenBits, s);
        } else {
            rc_encode(rc, le->choice, 1, 1);
            bt_encode(rc, le->high, kNumHighLenBits, s - kNumMidLenSymbols);
        }
    }
    if (--le->counters[pos_state] == 0) len_update_table(le, pos_state);
}

/* ------------------------------------------------------------------ */
/* LZMA/LiteralEncoder.java                                            */
/* ------------------------------------------------------------------ */
static void lit_encode(rc_enc *rc, uint16_t *probs, uint8_t symbol) { /* :17-24 */
    int context = 1;
    for (int i = 7; i >= 0; i--) {
        int bit = (symbol >> i) & 1;
        rc_encode(rc, probs, context, bit);
        context = (context << 1) | bit;
    }
}
static void lit_encode_matched(rc_enc *rc, uint16_t *probs, uint8_t match_byte, uint8_t symbol) { /* :26-40 */
    int context = 1;
    int same = 1;
    for (int i = 7; i >= 0; i--) {
        int bit = (symbol >> i) & 1;
        int state = context;
        if (same) {
            int match_bit = (match_byte >> i) & 1;
            state += (1 + match_bit) << 8;
            same = (match_bit == bit);
        }
        rc_encode(rc, probs, state, bit);
        context = (context << 1) | bit;
    }
}
static int32_t lit_price(const uint16_t *probs, int match_mode, uint8_t match_byte, uint8_t symbol) { /* :42-64 */
    int32_t price = 0;
    int context = 1;
    int i = 7;
    if (match_mode) {
        for (; i >= 0; i--) {
            int match_bit = (match_byte >> i) & 1;
            int bit = (symbol >> i) & 1;
            price += price_bit(probs[((1 + match_bit) << 8) + context], bit);
            context = (context << 1) | bit;
            if (match_bit != bit) {
                i--;
                break;
            }
        }
    }
    for (; i >= 0; i--) {
        int bit = (symbol >> i) & 1;
        price += price_bit(probs[context], bit);
        context = (context << 1) | bit;
    }
    return price;
}

/* ------------------------------------------------------------------ */
/* LZ/BinTree.java over a whole-buffer LZ/InWindow.java                */
/* ------------------------------------------------------------------ */
typedef struct { int32_t length; int32_t distance; } len_dist; /* BinTree.java:22-39 */

typedef struct {
    const uint8_t *data;  /* data[p] == _bufferBase[_bufferOffset + p + 1] */
    int64_t n;            /* stream length                                 */
    int64_t pos1;         /* _pos: 1-based after Init's reduceOffsets(-1)  */
    int64_t stream_pos1;  /* _streamPos = n + 1                            */
    int32_t cyclic_pos, cyclic_size;
    int32_t match_max_len, cut_value;
    uint32_t hash_mask;
    int32_t hash_size_sum;
    int hash_array;       /* HASH_ARRAY */
    int num_hash_direct_bytes, min_match_check, fix_hash_size;
    int32_t *son;
    int32_t *hash;
    lzo_trace *trace;
} bin_tree;

static void bt_set_type(bin_tree *bt, int num_hash_bytes) { /* :59-70 */
    bt->hash_array = num_hash_bytes > 2;
    if (bt->hash_array) {
        bt->num_hash_direct_bytes = 0;
        bt->min_match_check = 4;
        bt->fix_hash_size = kHash2Size + kHash3Size;
    } else {
        bt->num_hash_direct_bytes = 2;
        bt->min_match_check = 2 + 1;
        bt->fix_hash_size = 0;
    }
}

static int bt_create(bin_tree *bt, int32_t history_size, int match_max_len) { /* :93-134 */
    if (history_size > kMaxValForNormalize - 256) return 0;
    bt->cut_value = 16 + (match_max_len >> 1);
    bt->match_max_len = match_max_len;
    bt->cyclic_size = history_size + 1;
    bt->son = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)bt->cyclic_size);
    uint32_t hs = kBT2HashSize;
    if (bt->hash_array) {
        hs = (uint32_t)history_size - 1;
        hs |= hs >> 1;
        hs |= hs >> 2;
        hs |= hs >> 4;
        hs |= hs >> 8;
        hs >>= 1;
        hs |= 0xFFFF;
        if (hs > (1u << 24)) hs >>= 1;
        bt->hash_mask = hs;
        hs++;
        hs += (uint32_t)bt->fix_hash_size;
    }
    bt->hash_size_sum = (int32_t)hs;
    bt->hash = (int32_t *)calloc(hs, sizeof(int32_t)); /* Init :75-77 zero-fill */
    return bt->son != NULL && bt->hash != NULL;
}

static void bt_init(bin_tree *bt, const uint8_t *data, int64_t n) { /* :73-80 + InWindow.Init */
    bt->data = data;
    bt->n = n;
    bt->pos1 = 1;
    bt->stream_pos1 = n + 1;
    bt->cyclic_pos = 0;
}

static inline void bt_increment(bin_tree *bt) { /* :83-91 (Normalize unreachable: n < 2^30) */
    if (++bt->cyclic_pos >= bt->cyclic_size) bt->cyclic_pos = 0;
    bt->pos1++;
}

static void trace_mf(lzo_trace *t, int64_t p, const len_dist *d, int count) {
    if (!t || !t->mf_off) return;
    t->mf_off[p] = (uint32_t)t->mf_pairs_used;
    for (int i = 0; i < count; i++) {
        if (t->mf_pairs && t->mf_pairs_used < t->mf_pairs_cap) {
            t->mf_pairs[2 * t->mf_pairs_used] = (uint32_t)d[i].length;
            t->mf_pairs[2 * t->mf_pairs_used + 1] = (uint32_t)d[i].distance;
            t->mf_pairs_used++;
        } else {
            t->mf_overflow++;
        }
    }
    t->mf_off[p + 1] = (uint32_t)t->mf_pairs_used;
}

/* fillMatches0 (:152-273) when `distances` != NULL, Skip's body (:276-354)
 * when NULL -- the two update the tree identically;
 * with a trace attached Skip()ped positions run the recording variant into a
 * scratch list so that the tap sees every position. */
static int bt_get_matches(bin_tree *bt, len_dist *distances) {
    const uint8_t *buf = bt->data - 1; /* buf[pos1] == data[pos1 - 1] */
    int32_t len_limit;
    if (bt->pos1 + bt->match_max_len <= bt->stream_pos1) {
        len_limit = bt->match_max_len;
    } else {
        len_limit = (int32_t)(bt->stream_pos1 - bt->pos1);
        if (len_limit < bt->min_match_check) {
            if (bt->trace) trace_mf(bt->trace, bt->pos1 - 1, NULL, 0);
            bt_increment(bt);
            return 0;
        }
    }
    const int32_t pos = (int32_t)bt->pos1;
    const int32_t match_min_pos = pos > bt->cyclic_size ? pos - bt->cyclic_size : 0;
    const uint8_t *cur = buf + pos;
    uint32_t hash_value, hash2 = 0, hash3 = 0;
    if (bt->hash_array) {
        uint32_t temp = g_crc[cur[0]] ^ cur[1];
        hash2 = temp & (kHash2Size - 1);
        temp ^= (uint32_t)cur[2] << 8;
        hash3 = temp & (kHash3Size - 1);
        hash_value = (temp ^ (g_crc[cur[3]] << 5)) & bt->hash_mask;
    } else {
        hash_value = cur[0] ^ ((uint32_t)cur[1] << 8);
    }

    int32_t cur_match = bt->hash[bt->fix_hash_size + hash_value];
    int32_t max_len = kStartMaxLen;
    int offset = 0;
    if (bt->hash_array) {
        int32_t cur_match2 = bt->hash[hash2];
        const int32_t cur_match3 = bt->hash[kHash3Offset + hash3];
        bt->hash[hash2] = pos;
        bt->hash[kHash3Offset + hash3] = pos;
        if (cur_match2 > match_min_pos) {
            if (buf[cur_match2] == cur[0]) {
                max_len = 2;
                distances[offset].length = 2;
                distances[offset++].distance = pos - cur_match2 - 1;
            }
        }
        if (cur_match3 > match_min_pos) {
            if (buf[cur_match3] == cur[0]) {
                if (cur_match3 == cur_match2) offset--;
                max_len = 3;
                distances[offset].length = 3;
                distances[offset++].distance = pos - cur_match3 - 1;
                cur_match2 = cur_match3;
            }
        }
        if (offset != 0 && cur_match2 == cur_match) {
            offset--;
            max_len = kStartMaxLen;
        }
    }

    bt->hash[bt->fix_hash_size + hash_value] = pos;

    int32_t ptr0 = (bt->cyclic_pos << 1) + 1;
    int32_t ptr1 = (bt->cyclic_pos << 1);
    int32_t len0 = bt->num_hash_direct_bytes, len1 = bt->num_hash_direct_bytes;

    if (bt->num_hash_direct_bytes != 0) {
        if (cur_match > match_min_pos) {
            if (buf[cur_match + bt->num_hash_direct_bytes] != cur[bt->num_hash_direct_bytes]) {
                max_len = bt->num_hash_direct_bytes;
                distances[offset].length = max_len;
                distances[offset++].distance = pos - cur_match - 1;
            }
        }
    }

    int32_t count = bt->cut_value;
    for (;;) {
        if (cur_match <= match_min_pos || count-- == 0) {
            bt->son[ptr0] = kEmptyHashValue;
            bt->son[ptr1] = kEmptyHashValue;
            break;
        }
        const int32_t delta = pos - cur_match;
        const int32_t cyclic = ((delta <= bt->cyclic_pos) ? (bt->cyclic_pos - delta)
                                                          : (bt->cyclic_pos - delta + bt->cyclic_size)) << 1;
        const uint8_t *pby1 = buf + cur_match;
        int32_t len = len0 < len1 ? len0 : len1;
        if (pby1[len] == cur[len]) {
            while (++len != len_limit)
                if (pby1[len] != cur[len]) break;
            if (max_len < len) {
                max_len = len;
                distances[offset].length = len;
                distances[offset++].distance = delta - 1;
                if (len == len_limit) {
                    bt->son[ptr1] = bt->son[cyclic];
                    bt->son[ptr0] = bt->son[cyclic + 1];
                    break;
                }
            }
        }
        if (pby1[len] < cur[len]) {
            bt->son[ptr1] = cur_match;
            ptr1 = cyclic + 1;
            cur_match = bt->son[ptr1];
            len1 = len;
        } else {
            bt->son[ptr0] = cur_match;
            ptr0 = cyclic;
            cur_match = bt->son[ptr0];
            len0 = len;
        }
    }
    if (bt->trace) trace_mf(bt->trace, bt->pos1 - 1, distances, offset);
    bt_increment(bt);
    return offset;
}

static void bt_skip(bin_tree *bt, int num) { /* :275-356 */
    len_dist scratch[kMatchMaxLen + 1];
    do {
        bt_get_matches(bt, scratch);
    } while (--num != 0);
}

/* InWindow.java:115-138, whole-buffer form; p = pos1 - 1 */
static inline uint8_t bt_index_byte(const bin_tree *bt, int index) {
    return bt->data[bt->pos1 - 1 + index];
}
static int bt_match_len(const bin_tree *bt, int index, int32_t distance, int limit) {
    int64_t s = bt->pos1 + index; /* 1-based */
    if (s + limit > bt->stream_pos1) limit = (int)(bt->stream_pos1 - s);
    distance++;
    const uint8_t *pby = bt->data + (s - 1);
    int i;
    for (i = 0; i < limit && pby[i] == pby[(int64_t)i - distance]; i++) {}
    return i;
}
static inline int32_t bt_avail(const bin_tree *bt) { return (int32_t)(bt->stream_pos1 - bt->pos1); }

/* ------------------------------------------------------------------ */
/* LZMA/Optimal.java, LZMA/Encoder.java                                */
/* ------------------------------------------------------------------ */
typedef struct {
    int32_t state;
    int prev1_is_char, prev2;
    int32_t pos_prev2, back_prev2;
    int32_t price, pos_prev, back_prev;
    int32_t backs0, backs1, backs2, backs3;
} optimal;

static inline void opt_make_char(optimal *o) { o->back_prev = -1; o->prev1_is_char = 0; }
static inline void opt_make_shortrep(optimal *o) { o->back_prev = 0; o->prev1_is_char = 0; }
static inline int opt_is_shortrep(const optimal *o) { return o->back_prev == 0; }

typedef struct {
    rc_enc rc;
    optimal *optimum; /* [kNumOpts] */

    uint16_t is_match[kNumStates << kNumPosStatesBitsMax];
    uint16_t is_rep[kNumStates];
    uint16_t is_rep_g0[kNumStates];
    uint16_t is_rep_g1[kNumStates];
    uint16_t is_rep_g2[kNumStates];
    uint16_t is_rep0_long[kNumStates << kNumPosStatesBitsMax];
    uint16_t pos_slot_enc[kNumLenToPosStates][1 << kNumPosSlotBits];
    uint16_t pos_encoders[kNumFullDistances - kEndPosModelIndex];
    uint16_t pos_align_enc[1 << kNumAlignBits];
    len_enc len_encoder;
    len_enc rep_len_encoder;
    uint16_t *literal; /* 0x300 << (lc + lp) */

    len_dist match_distances[kMatchMaxLen + 1];
    int32_t match_price_count;
    bin_tree mf;

    int32_t num_fast_bytes;
    int32_t longest_match_length;
    int32_t num_distance_pairs;
    int32_t additional_offset;
    int32_t optimum_end_index, optimum_current_index;
    int longest_match_was_found;

    int32_t pos_slot_prices[1 << (kNumPosSlotBits + kNumLenToPosStatesBits)];
    int32_t distances_prices[kNumFullDistances << kNumLenToPosStatesBits];
    int32_t align_prices[kAlignTableSize];
    int32_t align_price_count;
    int32_t dist_table_size;

    int32_t pos_state_bits, pos_state_mask, lp, lc;
    int32_t dictionary_size;
    int write_end_marker;

    int32_t state;
    uint8_t previous_byte;
    int32_t rep_distances[kNumRepDistances];
    int32_t reps[kNumRepDistances];
    int32_t rep_lens[kNumRepDistances];

    int64_t now_pos64;
    lzo_trace *trace;
} encoder;

typedef struct { int32_t pos; int32_t length; } pos_len; /* Encoder.PosAndLength */

static inline uint16_t *lit_sub_coder(encoder *e, int32_t pos, uint8_t prev_byte) { /* LiteralEncoder.java:93-95 */
    int lp_mask = (1 << e->lp) - 1;
    return e->literal + 0x300u * (size_t)(((pos & lp_mask) << e->lc) + (prev_byte >> (8 - e->lc)));
}

static int read_match_distances(encoder *e) { /* Encoder.java:275-287 */
    e->num_distance_pairs = bt_get_matches(&e->mf, e->match_distances);
    int length = 0;
    if (e->num_distance_pairs > 0) {
        const len_dist *big = &e->match_distances[e->num_distance_pairs - 1];
        length = big->length;
        if (length == e->num_fast_bytes)
            length += bt_match_len(&e->mf, length - 1, big->distance, kMatchMaxLen - length);
    }
    e->additional_offset++;
    return length;
}

static void move_pos(encoder *e, int num) { /* :289-294 */
    if (num > 0) {
        bt_skip(&e->mf, num);
        e->additional_offset += num;
    }
}

static int32_t rep_len1_price(encoder *e, int state, int pos_state) { /* :296-299 */
    return price0(e->is_rep_g0[state]) + price0(e->is_rep0_long[(state << kNumPosStatesBitsMax) + pos_state]);
}
static int32_t pure_rep_price(encoder *e, int rep_index, int state, int pos_state) { /* :301-316 */
    int32_t price;
    if (rep_index == 0) {
        price = price0(e->is_rep_g0[state]);
        price += price1(e->is_rep0_long[(state << kNumPosStatesBitsMax) + pos_state]);
    } else {
        price = price1(e->is_rep_g0[state]);
        if (rep_index == 1) {
            price += price0(e->is_rep_g1[state]);
        } else {
            price += price1(e->is_rep_g1[state]);
            price += price_bit(e->is_rep_g2[state], rep_index - 2);
        }
    }
    return price;
}
static int32_t rep_price(encoder *e, int rep_index, int len, int state, int pos_state) { /* :318-321 */
    return len_get_price(&e->rep_len_encoder, len - kMatchMinLen, pos_state) + pure_rep_price(e, rep_index, state, pos_state);
}
static int32_t pos_len_price(encoder *e, int32_t pos, int len, int pos_state) { /* :323-333 */
    int32_t price;
    int lps = len_to_pos_state(len);
    if (pos < kNumFullDistances)
        price = e->distances_prices[lps * kNumFullDistances + pos];
    else
        price = e->pos_slot_prices[(lps << kNumPosSlotBits) + get_pos_slot2((uint32_t)pos)] + e->align_prices[pos & kAlignMask];
    return price + len_get_price(&e->len_encoder, len - kMatchMinLen, pos_state);
}

static pos_len backward(encoder *e, int cur) { /* :335-362 */
    optimal *opt = e->optimum;
    e->optimum_end_index = cur;
    int pos_mem = opt[cur].pos_prev;
    int back_mem = opt[cur].back_prev;
    do {
        if (opt[cur].prev1_is_char) {
            opt_make_char(&opt[pos_mem]);
            opt[pos_mem].pos_prev = pos_mem - 1;
            if (opt[cur].prev2) {
                opt[pos_mem - 1].prev1_is_char = 0;
                opt[pos_mem - 1].pos_prev = opt[cur].pos_prev2;
                opt[pos_mem - 1].back_prev = opt[cur].back_prev2;
            }
        }
        int pos_prev = pos_mem;
        int back_cur = back_mem;
        back_mem = opt[pos_prev].back_prev;
        pos_mem = opt[pos_prev].pos_prev;
        opt[pos_prev].back_prev = back_cur;
        opt[pos_prev].pos_prev = cur;
        cur = pos_prev;
    } while (cur > 0);
    e->optimum_current_index = opt[0].pos_prev;
    pos_len r = { opt[0].back_prev, e->optimum_current_index };
    return r;
}

static pos_len get_optimum(encoder *e, int32_t position) { /* :364-811 */
    optimal *opt = e->optimum;
    pos_len r;
    if (e->optimum_end_index != e->optimum_current_index) { /* :365-370 */
        r.length = opt[e->optimum_current_index].pos_prev - e->optimum_current_index;
        r.pos = opt[e->optimum_current_index].back_prev;
        e->optimum_current_index = opt[e->optimum_current_index].pos_prev;
        return r;
    }
    e->optimum_current_index = 0;
    e->optimum_end_index = 0;

    int len_main;
    if (e->longest_match_was_found) {
        len_main = e->longest_match_length;
        e->longest_match_was_found = 0;
    } else {
        len_main = read_match_distances(e);
    }
    int num_distance_pairs = e->num_distance_pairs;

    int num_avail = bt_avail(&e->mf) + 1;
    if (num_avail < 2) { r.pos = -1; r.length = 1; return r; }
    if (num_avail > kMatchMaxLen) num_avail = kMatchMaxLen;

    int rep_max_index = 0;
    int i;
    for (i = 0; i < kNumRepDistances; i++) { /* :393-399 */
        e->reps[i] = e->rep_distances[i];
        e->rep_lens[i] = bt_match_len(&e->mf, 0 - 1, e->reps[i], kMatchMaxLen);
        if (e->rep_lens[i] > e->rep_lens[rep_max_index]) rep_max_index = i;
    }
    if (e->rep_lens[rep_max_index] >= e->num_fast_bytes) { /* :400-404 */
        r.length = e->rep_lens[rep_max_index];
        r.pos = rep_max_index;
        move_pos(e, r.length - 1);
        return r;
    }
    if (len_main >= e->num_fast_bytes) { /* :406-410 */
        r.pos = e->match_distances[num_distance_pairs - 1].distance + kNumRepDistances;
        r.length = len_main;
        move_pos(e, len_main - 1);
        return r;
    }

    uint8_t current_byte = bt_index_byte(&e->mf, 0 - 1);
    uint8_t match_byte = bt_index_byte(&e->mf, 0 - e->rep_distances[0] - 1 - 1);

    if (len_main < 2 && current_byte != match_byte && e->rep_lens[rep_max_index] < 2) { /* :415-417 */
        r.pos = -1; r.length = 1; return r;
    }

    opt[0].state = e->state;
    int pos_state = position & e->pos_state_mask;

    opt[1].price = price0(e->is_match[(e->state << kNumPosStatesBitsMax) + pos_state]) +
        lit_price(lit_sub_coder(e, position, e->previous_byte), !st_is_char(e->state), match_byte, current_byte);
    opt_make_char(&opt[1]);

    int32_t match_price = price1(e->is_match[(e->state << kNumPosStatesBitsMax) + pos_state]);
    int32_t rep_match_price = match_price + price1(e->is_rep[e->state]);

    if (match_byte == current_byte) { /* :430-436 */
        int32_t short_rep_price = rep_match_price + rep_len1_price(e, e->state, pos_state);
        if (short_rep_price < opt[1].price) {
            opt[1].price = short_rep_price;
            opt_make_shortrep(&opt[1]);
        }
    }

    int len_end = len_main >= e->rep_lens[rep_max_index] ? len_main : e->rep_lens[rep_max_index];
    if (len_end < 2) { r.pos = opt[1].back_prev; r.length = 1; return r; }

    opt[1].pos_prev = 0;
    opt[0].backs0 = e->reps[0];
    opt[0].backs1 = e->reps[1];
    opt[0].backs2 = e->reps[2];
    opt[0].backs3 = e->reps[3];

    int len = len_end;
    do {
        opt[len--].price = kIfinityPrice;
    } while (len >= 2);

    for (i = 0; i < kNumRepDistances; i++) { /* :457-474 */
        int rep_len = e->rep_lens[i];
        if (rep_len < 2) continue;
        int32_t price = rep_match_price + pure_rep_price(e, i, e->state, pos_state);
        do {
            int32_t cur_and_len_price = price + len_get_price(&e->rep_len_encoder, rep_len - 2, pos_state);
            optimal *o = &opt[rep_len];
            if (cur_and_len_price < o->price) {
                o->price = cur_and_len_price;
                o->pos_prev = 0;
                o->back_prev = i;
                o->prev1_is_char = 0;
            }
        } while (--rep_len >= 2);
    }

    int32_t normal_match_price = match_price + price0(e->is_rep[e->state]);

    len = e->rep_lens[0] >= 2 ? e->rep_lens[0] + 1 : 2; /* :478-501 */
    if (len <= len_main) {
        int offs = 0;
        while (len > e->match_distances[offs].length) offs++;
        for (;; len++) {
            int32_t distance = e->match_distances[offs].distance;
            int32_t cur_and_len_price = normal_match_price + pos_len_price(e, distance, len, pos_state);
            optimal *o = &opt[len];
            if (cur_and_len_price < o->price) {
                o->price = cur_and_len_price;
                o->pos_prev = 0;
                o->back_prev = distance + kNumRepDistances;
                o->prev1_is_char = 0;
            }
            if (len == e->match_distances[offs].length) {
                offs++;
                if (offs == num_distance_pairs) break;
            }
        }
    }

    int cur = 0;
    for (;;) { /* :505-810 */
        cur++;
        if (cur == len_end) return backward(e, cur);
        int new_len = read_match_distances(e);
        num_distance_pairs = e->num_distance_pairs;
        if (new_len >= e->num_fast_bytes) {
            e->longest_match_length = new_len;
            e->longest_match_was_found = 1;
            return backward(e, cur);
        }
        position++;
        int pos_prev = opt[cur].pos_prev;
        int state;
        if (opt[cur].prev1_is_char) { /* :520-535 */
            pos_prev--;
            if (opt[cur].prev2) {
                state = opt[opt[cur].pos_prev2].state;
                if (opt[cur].back_prev2 < kNumRepDistances) state = st_longrep(state);
                else state = st_match(state);
            } else {
                state = opt[pos_prev].state;
            }
            state = st_lit(state);
        } else {
            state = opt[pos_prev].state;
        }
        if (pos_prev == cur - 1) { /* :536-585 */
            if (opt_is_shortrep(&opt[cur])) state = st_shortrep(state);
            else state = st_lit(state);
        } else {
            int pos;
            if (opt[cur].prev1_is_char && opt[cur].prev2) {
                pos_prev = opt[cur].pos_prev2;
                pos = opt[cur].back_prev2;
                state = st_longrep(state);
            } else {
                pos = opt[cur].back_prev;
                if (pos < kNumRepDistances) state = st_longrep(state);
                else state = st_match(state);
            }
            const optimal *o = &opt[pos_prev];
            if (pos < kNumRepDistances) {
                if (pos == 0) {
                    e->reps[0] = o->backs0; e->reps[1] = o->backs1; e->reps[2] = o->backs2; e->reps[3] = o->backs3;
                } else if (pos == 1) {
                    e->reps[0] = o->backs1; e->reps[1] = o->backs0; e->reps[2] = o->backs2; e->reps[3] = o->backs3;
                } else if (pos == 2) {
                    e->reps[0] = o->backs2; e->reps[1] = o->backs0; e->reps[2] = o->backs1; e->reps[3] = o->backs3;
                } else {
                    e->reps[0] = o->backs3; e->reps[1] = o->backs0; e->reps[2] = o->backs1; e->reps[3] = o->backs2;
                }
            } else {
                e->reps[0] = pos - kNumRepDistances;
                e->reps[1] = o->backs0; e->reps[2] = o->backs1; e->reps[3] = o->backs2;
            }
        }
        opt[cur].state = state;
        opt[cur].backs0 = e->reps[0];
        opt[cur].backs1 = e->reps[1];
        opt[cur].backs2 = e->reps[2];
        opt[cur].backs3 = e->reps[3];
        const int32_t cur_price = opt[cur].price;

        current_byte = bt_index_byte(&e->mf, 0 - 1);
        match_byte = bt_index_byte(&e->mf, 0 - e->reps[0] - 1 - 1);
        pos_state = position & e->pos_state_mask;

        const int32_t cur_and1_price = cur_price +
            price0(e->is_match[(state << kNumPosStatesBitsMax) + pos_state]) +
            lit_price(lit_sub_coder(e, position, bt_index_byte(&e->mf, 0 - 2)), !st_is_char(state), match_byte, current_byte);

        optimal *next = &opt[cur + 1];
        int next_is_char = 0;
        if (cur_and1_price < next->price) { /* :606-611 */
            next->price = cur_and1_price;
            next->pos_prev = cur;
            opt_make_char(next);
            next_is_char = 1;
        }

        match_price = cur_price + price1(e->is_match[(state << kNumPosStatesBitsMax) + pos_state]);
        rep_match_price = match_price + price1(e->is_rep[state]);

        if (match_byte == current_byte && !(next->pos_prev < cur && next->back_prev == 0)) { /* :616-625 */
            int32_t short_rep_price = rep_match_price + rep_len1_price(e, state, pos_state);
            if (short_rep_price <= next->price) {
                next->price = short_rep_price;
                next->pos_prev = cur;
                opt_make_shortrep(next);
                next_is_char = 1;
            }
        }

        int num_avail_full = bt_avail(&e->mf) + 1; /* :627-636 */
        if (kNumOpts - 1 - cur < num_avail_full) num_avail_full = kNumOpts - 1 - cur;
        num_avail = num_avail_full;
        if (num_avail < 2) continue;
        if (num_avail > e->num_fast_bytes) num_avail = e->num_fast_bytes;

        if (!next_is_char && match_byte != current_byte) { /* :637-665 literal + rep0 */
            int t = num_avail_full - 1 < e->num_fast_bytes ? num_avail_full - 1 : e->num_fast_bytes;
            int len_test2 = bt_match_len(&e->mf, 0, e->reps[0], t);
            if (len_test2 >= 2) {
                int state2 = st_lit(state);
                int pos_state_next = (position + 1) & e->pos_state_mask;
                int32_t next_rep_match_price = cur_and1_price +
                    price1(e->is_match[(state2 << kNumPosStatesBitsMax) + pos_state_next]) +
                    price1(e->is_rep[state2]);
                int offset = cur + 1 + len_test2;
                while (len_end < offset) opt[++len_end].price = kIfinityPrice;
                int32_t cur_and_len_price = next_rep_match_price + rep_price(e, 0, len_test2, state2, pos_state_next);
                optimal *o = &opt[offset];
                if (cur_and_len_price < o->price) {
                    o->price = cur_and_len_price;
                    o->pos_prev = cur + 1;
                    o->back_prev = 0;
                    o->prev1_is_char = 1;
                    o->prev2 = 0;
                }
            }
        }

        int start_len = 2;

        for (int rep_index = 0; rep_index < kNumRepDistances; rep_index++) { /* :669-735 */
            int len_test = bt_match_len(&e->mf, 0 - 1, e->reps[rep_index], num_avail);
            if (len_test < 2) continue;
            const int len_test_temp = len_test;
            do {
                while (len_end < cur + len_test) opt[++len_end].price = kIfinityPrice;
                int32_t cur_and_len_price = rep_match_price + rep_price(e, rep_index, len_test, state, pos_state);
                optimal *o = &opt[cur + len_test];
                if (cur_and_len_price < o->price) {
                    o->price = cur_and_len_price;
                    o->pos_prev = cur;
                    o->back_prev = rep_index;
                    o->prev1_is_char = 0;
                }
            } while (--len_test >= 2);
            len_test = len_test_temp;

            if (rep_index == 0) start_len = len_test + 1;

            if (len_test < num_avail_full) { /* :696-734 rep + literal + rep0 */
                int t = num_avail_full - 1 - len_test < e->num_fast_bytes ? num_avail_full - 1 - len_test : e->num_fast_bytes;
                int len_test2 = bt_match_len(&e->mf, len_test, e->reps[rep_index], t);
                if (len_test2 >= 2) {
                    int state2 = st_longrep(state);
                    int pos_state_next = (position + len_test) & e->pos_state_mask;
                    int32_t cur_and_len_char_price =
                        rep_match_price + rep_price(e, rep_index, len_test, state, pos_state) +
                        price0(e->is_match[(state2 << kNumPosStatesBitsMax) + pos_state_next]) +
                        lit_price(lit_sub_coder(e, position + len_test, bt_index_byte(&e->mf, len_test - 1 - 1)), 1,
                                  bt_index_byte(&e->mf, len_test - 1 - (e->reps[rep_index] + 1)),
                                  bt_index_byte(&e->mf, len_test - 1));
                    state2 = st_lit(state2);
                    pos_state_next = (position + len_test + 1) & e->pos_state_mask;
                    int32_t next_match_price = cur_and_len_char_price + price1(e->is_match[(state2 << kNumPosStatesBitsMax) + pos_state_next]);
                    int32_t next_rep_match_price = next_match_price + price1(e->is_rep[state2]);

                    int offset = len_test + 1 + len_test2;
                    while (len_end < cur + offset) opt[++len_end].price = kIfinityPrice;
                    int32_t cur_and_len_price = next_rep_match_price + rep_price(e, 0, len_test2, state2, pos_state_next);
                    optimal *o = &opt[cur + offset];
                    if (cur_and_len_price < o->price) {
                        o->price = cur_and_len_price;
                        o->pos_prev = cur + len_test + 1;
                        o->back_prev = 0;
                        o->prev1_is_char = 1;
                        o->prev2 = 1;
                        o->pos_prev2 = cur;
                        o->back_prev2 = rep_index;
                    }
                }
            }
        }

        if (new_len > num_avail) { /* :737-743 */
            new_len = num_avail;
            for (num_distance_pairs = 0; new_len > e->match_distances[num_distance_pairs].length; num_distance_pairs++) {}
            e->match_distances[num_distance_pairs].length = new_len;
            num_distance_pairs++;
        }
        if (new_len >= start_len) { /* :744-809 */
            normal_match_price = match_price + price0(e->is_rep[state]);
            while (len_end < cur + new_len) opt[++len_end].price = kIfinityPrice;

            int offs = 0;
            while (start_len > e->match_distances[offs].length) offs++;

            for (int len_test = start_len;; len_test++) {
                int32_t cur_back = e->match_distances[offs].distance;
                int32_t cur_and_len_price = normal_match_price + pos_len_price(e, cur_back, len_test, pos_state);
                optimal *o = &opt[cur + len_test];
                if (cur_and_len_price < o->price) {
                    o->price = cur_and_len_price;
                    o->pos_prev = cur;
                    o->back_prev = cur_back + kNumRepDistances;
                    o->prev1_is_char = 0;
                }

                if (len_test == e->match_distances[offs].length) {
                    if (len_test < num_avail_full) { /* match + literal + rep0 */
                        int t = num_avail_full - 1 - len_test < e->num_fast_bytes ? num_avail_full - 1 - len_test : e->num_fast_bytes;
                        int len_test2 = bt_match_len(&e->mf, len_test, cur_back, t);
                        if (len_test2 >= 2) {
                            int state2 = st_match(state);
                            int pos_state_next = (position + len_test) & e->pos_state_mask;
                            int32_t cur_and_len_char_price = cur_and_len_price +
                                price0(e->is_match[(state2 << kNumPosStatesBitsMax) + pos_state_next]) +
                                lit_price(lit_sub_coder(e, position + len_test, bt_index_byte(&e->mf, len_test - 1 - 1)), 1,
                                          bt_index_byte(&e->mf, len_test - (cur_back + 1) - 1),
                                          bt_index_byte(&e->mf, len_test - 1));
                            state2 = st_lit(state2);
                            pos_state_next = (position + len_test + 1) & e->pos_state_mask;
                            int32_t next_match_price = cur_and_len_char_price + price1(e->is_match[(state2 << kNumPosStatesBitsMax) + pos_state_next]);
                            int32_t next_rep_match_price = next_match_price + price1(e->is_rep[state2]);

                            int offset = len_test + 1 + len_test2;
                            while (len_end < cur + offset) opt[++len_end].price = kIfinityPrice;
                            cur_and_len_price = next_rep_match_price + rep_price(e, 0, len_test2, state2, pos_state_next);
                            o = &opt[cur + offset];
                            if (cur_and_len_price < o->price) {
                                o->price = cur_and_len_price;
                                o->pos_prev = cur + len_test + 1;
                                o->back_prev = 0;
                                o->prev1_is_char = 1;
                                o->prev2 = 1;
                                o->pos_prev2 = cur;
                                o->back_prev2 = cur_back + kNumRepDistances;
                            }
                        }
                    }
                    offs++;
                    if (offs == num_distance_pairs) break;
                }
            }
        }
    }
}

static void fill_distances_prices(encoder *e) { /* :1087-1118 */
    int32_t temp_prices[kNumFullDistances];
    for (int i = kStartPosModelIndex; i < kNumFullDistances; i++) {
        int pos_slot = get_pos_slot((uint32_t)i);
        int footer_bits = (pos_slot >> 1) - 1;
        int base_val = (2 | (pos_slot & 1)) << footer_bits;
        temp_prices[i] = bt_reverse_price(e->pos_encoders + (base_val - pos_slot - 1), footer_bits, i - base_val);
    }
    for (int lps = 0; lps < kNumLenToPosStates; lps++) {
        int pos_slot;
        const uint16_t *enc = e->pos_slot_enc[lps];
        int st = lps << kNumPosSlotBits;
        for (pos_slot = 0; pos_slot < e->dist_table_size; pos_slot++)
            e->pos_slot_prices[st + pos_slot] = bt_price(enc, kNumPosSlotBits, pos_slot);
        for (pos_slot = kEndPosModelIndex; pos_slot < e->dist_table_size; pos_slot++)
            e->pos_slot_prices[st + pos_slot] += (((pos_slot >> 1) - 1) - kNumAlignBits) << kNumBitPriceShiftBits;
        int st2 = lps * kNumFullDistances;
        int i;
        for (i = 0; i < kStartPosModelIndex; i++) e->distances_prices[st2 + i] = e->pos_slot_prices[st + i];
        for (; i < kNumFullDistances; i++)
            e->distances_prices[st2 + i] = e->pos_slot_prices[st + get_pos_slot((uint32_t)i)] + temp_prices[i];
    }
    e->match_price_count = 0;
}

static void fill_align_prices(encoder *e) { /* :1120-1125 */
    for (int i = 0; i < kAlignTableSize; i++)
        e->align_prices[i] = bt_reverse_price(e->pos_align_enc, kNumAlignBits, i);
    e->align_price_count = 0;
}

static void write_end_marker(encoder *e, int pos_state) { /* :818-835 */
    if (!e->write_end_marker) return;
    rc_encode(&e->rc, e->is_match, (e->state << kNumPosStatesBitsMax) + pos_state, 1);
    rc_encode(&e->rc, e->is_rep, e->state, 0);
    e->state = st_match(e->state);
    int len = kMatchMinLen;
    len_encode(&e->len_encoder, &e->rc, len - kMatchMinLen, pos_state);
    int pos_slot = (1 << kNumPosSlotBits) - 1;
    int lps = len_to_pos_state(len);
    bt_encode(&e->rc, e->pos_slot_enc[lps], kNumPosSlotBits, pos_slot);
    int footer_bits = 30;
    uint32_t pos_reduced = (1u << footer_bits) - 1;
    rc_encode_direct(&e->rc, pos_reduced >> kNumAlignBits, footer_bits - kNumAlignBits);
    bt_reverse_encode(&e->rc, e->pos_align_enc, kNumAlignBits, (int)(pos_reduced & kAlignMask));
}

static void enc_flush(encoder *e, int32_t now_pos) { /* :837-841 */
    write_end_marker(e, now_pos & e->pos_state_mask);
    rc_flush(&e->rc);
}

static void trace_decision(lzo_trace *t, int64_t off, pos_len d) {
    if (!t || !t->dec) return;
    if (t->dec_used < t->dec_cap) {
        t->dec[3 * t->dec_used] = off;
        t->dec[3 * t->dec_used + 1] = d.pos;
        t->dec[3 * t->dec_used + 2] = d.length;
    }
    t->dec_used++;
}

/* encodeOne (:890-936) + its three emitters (:938-1024).  Returns 0 when the
 * stream has been flushed.  The 4096-byte slicing of CodeOneBlock (:929-933)
 * only serves the progress callback and is not modelled. */
static int encode_one(encoder *e) {
    pos_len d = get_optimum(e, (int32_t)e->now_pos64);
    trace_decision(e->trace, e->now_pos64, d);
    int pos_state = (int32_t)e->now_pos64 & e->pos_state_mask;
    int complex_state = (e->state << kNumPosStatesBitsMax) + pos_state;
    if (d.length == 1 && d.pos == -1) {
        rc_encode(&e->rc, e->is_match, complex_state, 0);
        /* encodeSingleByteLiteral :1007-1024 */
        uint8_t cur_byte = bt_index_byte(&e->mf, 0 - e->additional_offset);
        uint16_t *sub = lit_sub_coder(e, (int32_t)e->now_pos64, e->previous_byte);
        if (st_is_char(e->state)) {
            lit_encode(&e->rc, sub, cur_byte);
        } else {
            uint8_t match_byte = bt_index_byte(&e->mf, 0 - e->rep_distances[0] - 1 - e->additional_offset);
            lit_encode_matched(&e->rc, sub, match_byte, cur_byte);
        }
        e->previous_byte = cur_byte;
        e->state = st_lit(e->state);
    } else {
        rc_encode(&e->rc, e->is_match, complex_state, 1);
        if (d.pos < kNumRepDistances) { /* encodeARepetition :938-974 */
            int pos = d.pos;
            rc_encode(&e->rc, e->is_rep, e->state, 1);
            if (pos == 0) {
                rc_encode(&e->rc, e->is_rep_g0, e->state, 0);
                rc_encode(&e->rc, e->is_rep0_long, complex_state, d.length == 1 ? 0 : 1);
            } else {
                rc_encode(&e->rc, e->is_rep_g0, e->state, 1);
                if (pos == 1) {
                    rc_encode(&e->rc, e->is_rep_g1, e->state, 0);
                } else {
                    rc_encode(&e->rc, e->is_rep_g1, e->state, 1);
                    rc_encode(&e->rc, e->is_rep_g2, e->state, pos - 2);
                }
            }
            if (d.length == 1) {
                e->state = st_shortrep(e->state);
            } else {
                len_encode(&e->rep_len_encoder, &e->rc, d.length - kMatchMinLen, pos_state);
                e->state = st_longrep(e->state);
            }
            int32_t distance = e->rep_distances[pos];
            if (pos != 0) {
                for (int k = pos; k >= 1; k--) e->rep_distances[k] = e->rep_distances[k - 1];
                e->rep_distances[0] = distance;
            }
        } else { /* encodeAMatch :976-1005 */
            rc_encode(&e->rc, e->is_rep, e->state, 0);
            e->state = st_match(e->state);
            len_encode(&e->len_encoder, &e->rc, d.length - kMatchMinLen, pos_state);
            int32_t pos = d.pos - kNumRepDistances;
            int pos_slot = get_pos_slot((uint32_t)pos);
            int lps = len_to_pos_state(d.length);
            bt_encode(&e->rc, e->pos_slot_enc[lps], kNumPosSlotBits, pos_slot);
            if (pos_slot >= kStartPosModelIndex) {
                int footer_bits = (pos_slot >> 1) - 1;
                int32_t base_val = (2 | (pos_slot & 1)) << footer_bits;
                int32_t pos_reduced = pos - base_val;
                if (pos_slot < kEndPosModelIndex) {
                    bt_reverse_encode(&e->rc, e->pos_encoders + (base_val - pos_slot - 1), footer_bits, pos_reduced);
                } else {
                    rc_encode_direct(&e->rc, (uint32_t)pos_reduced >> kNumAlignBits, footer_bits - kNumAlignBits);
                    bt_reverse_encode(&e->rc, e->pos_align_enc, kNumAlignBits, pos_reduced & kAlignMask);
                    e->align_price_count++;
                }
            }
            for (int k = kNumRepDistances - 1; k >= 1; k--) e->rep_distances[k] = e->rep_distances[k - 1];
            e->rep_distances[0] = pos;
            e->match_price_count++;
        }
        e->previous_byte = bt_index_byte(&e->mf, d.length - 1 - e->additional_offset);
    }
    e->additional_offset -= d.length;
    e->now_pos64 += d.length;
    if (e->additional_offset == 0) {
        if (e->match_price_count >= (1 << 7)) fill_distances_prices(e);
        if (e->align_price_count >= kAlignTableSize) fill_align_prices(e);
        if (bt_avail(&e->mf) == 0) {
            enc_flush(e, (int32_t)e->now_pos64);
            return 0;
        }
    }
    return 1;
}

int lzo_props_valid(const lzo_props *p) { /* Encoder.java:1135-1180 */
    if (p->dict_size < 1 || p->dict_size > (1 << 29)) return 0;
    if (p->fb < 5 || p->fb > kMatchMaxLen) return 0;
    if (p->mf < 0 || p->mf > 2) return 0;
    if (p->lp < 0 || p->lp > kNumLitPosStatesBitsEncodingMax || p->lc < 0 || p->lc > kNumLitContextBitsMax ||
        p->pb < 0 || p->pb > kNumPosStatesBitsEncodingMax)
        return 0;
    return 1;
}

void lzo_write_props(const lzo_props *p, uint8_t out[5]) { /* :1079-1085 */
    out[0] = (uint8_t)((p->pb * 5 + p->lp) * 9 + p->lc);
    for (int i = 0; i < 4; i++) out[1 + i] = (uint8_t)((uint32_t)p->dict_size >> (8 * i));
}

size_t lzo_encode_bound(size_t n) { return n + n / 3 + 128; }

size_t lzo_encode(const lzo_props *p, const uint8_t *in, size_t n, uint8_t *out, size_t out_cap, lzo_trace *trace) {
    pthread_once(&g_once, init_tables);
    if (!lzo_props_valid(p) || n >= (size_t)kMaxValForNormalize - 1) return (size_t)-1;
    encoder *e = (encoder *)calloc(1, sizeof(encoder)); /* Java zero-init (App. A #14) */
    if (!e) return (size_t)-1;
    size_t result = (size_t)-1;
    e->optimum = (optimal *)calloc(kNumOpts, sizeof(optimal));
    e->literal = (uint16_t *)malloc(sizeof(uint16_t) * (0x300u << (p->lc + p->lp)));
    /* setters :1135-1184 */
    e->dictionary_size = p->dict_size;
    int dic_log;
    for (dic_log = 0; (uint32_t)p->dict_size > (1u << dic_log); dic_log++) {}
    e->dist_table_size = dic_log * 2;
    e->num_fast_bytes = p->fb;
    e->lp = p->lp;
    e->lc = p->lc;
    e->pos_state_bits = p->pb;
    e->pos_state_mask = (1 << p->pb) - 1;
    e->write_end_marker = p->eos;
    e->trace = trace;
    if (trace) {
        trace->mf_pairs_used = 0;
        trace->mf_overflow = 0;
        trace->dec_used = 0;
        if (trace->mf_off) memset(trace->mf_off, 0, sizeof(uint32_t) * (n + 1));
    }
    /* Create :224-241 */
    bt_set_type(&e->mf, p->mf == 0 ? 2 : 4);
    e->mf.trace = trace;
    if (!e->optimum || !e->literal || !bt_create(&e->mf, p->dict_size, p->fb)) goto done;
    /* Init :247-273 */
    e->state = 0;
    e->previous_byte = 0;
    rc_init(&e->rc, out, out_cap);
    init_probs(e->is_match, sizeof e->is_match / 2);
    init_probs(e->is_rep, kNumStates);
    init_probs(e->is_rep_g0, kNumStates);
    init_probs(e->is_rep_g1, kNumStates);
    init_probs(e->is_rep_g2, kNumStates);
    init_probs(e->is_rep0_long, sizeof e->is_rep0_long / 2);
    init_probs(e->pos_encoders, sizeof e->pos_encoders / 2);
    init_probs(e->literal, 0x300u << (p->lc + p->lp));
    for (int i = 0; i < kNumLenToPosStates; i++) init_probs(e->pos_slot_enc[i], 1 << kNumPosSlotBits);
    len_init(&e->len_encoder, 1 << p->pb);
    len_init(&e->rep_len_encoder, 1 << p->pb);
    init_probs(e->pos_align_enc, 1 << kNumAlignBits);
    /* SetStreams :1053-1061 */
    fill_distances_prices(e);
    fill_align_prices(e);
    e->len_encoder.table_size = p->fb + 1 - kMatchMinLen;
    len_update_tables(&e->len_encoder, 1 << p->pb);
    e->rep_len_encoder.table_size = p->fb + 1 - kMatchMinLen;
    len_update_tables(&e->rep_len_encoder, 1 << p->pb);

    /* CodeOneBlock :843-888 */
    bt_init(&e->mf, in, (int64_t)n);
    e->now_pos64 = 0;
    if (bt_avail(&e->mf) == 0) {
        enc_flush(e, 0);
    } else {
        read_match_distances(e);
        int pos_state = 0 & e->pos_state_mask;
        rc_encode(&e->rc, e->is_match, (e->state << kNumPosStatesBitsMax) + pos_state, 0);
        e->state = st_lit(e->state);
        uint8_t cur_byte = bt_index_byte(&e->mf, 0 - e->additional_offset);
        lit_encode(&e->rc, lit_sub_coder(e, 0, e->previous_byte), cur_byte);
        e->previous_byte = cur_byte;
        e->additional_offset--;
        e->now_pos64++;
        if (bt_avail(&e->mf) == 0) {
            enc_flush(e, (int32_t)e->now_pos64);
        } else {
            while (encode_one(e)) {}
        }
    }
    if (!e->rc.overflow) result = e->rc.out_pos;
done:
    free(e->mf.son);
    free(e->mf.hash);
    free(e->literal);
    free(e->optimum);
    free(e);
    return result;
}

size_t lzo_encode_alone(const lzo_props *p, const uint8_t *in, size_t n, uint8_t *out, size_t out_cap) { /* LzmaAlone.java:208-217 */
    if (out_cap < 13) return (size_t)-1;
    lzo_write_props(p, out);
    uint64_t size = p->eos ? ~0ull : (uint64_t)n;
    for (int i = 0; i < 8; i++) out[5 + i] = (uint8_t)(size >> (8 * i));
    size_t r = lzo_encode(p, in, n, out + 13, out_cap - 13, NULL);
    return r == (size_t)-1 ? r : r + 13;
}

/* ------------------------------------------------------------------ */
/* RangeCoder/RangeDecoder.java, BitTreeDecoder.java, LZMA/Decoder.java */
/* ------------------------------------------------------------------ */
typedef struct {
    uint32_t range, code;
    const uint8_t *in;
    size_t in_pos, in_len;
} rc_dec;

static inline uint32_t rd_read(rc_dec *rd) { /* InputStream.read(): -1 at EOF is OR-ed in as all ones */
    if (rd->in_pos < rd->in_len) return rd->in[rd->in_pos++];
    rd->in_pos++;
    return 0xFFFFFFFFu;
}
static void rd_init(rc_dec *rd, const uint8_t *in, size_t len) { /* RangeDecoder.java:19-25 */
    rd->in = in;
    rd->in_len = len;
    rd->in_pos = 0;
    rd->code = 0;
    rd->range = 0xFFFFFFFFu;
    for (int i = 0; i < 5; i++) rd->code = (rd->code << 8) | rd_read(rd);
}
static inline int rd_bit(rc_dec *rd, uint16_t *probs, int index) { /* :43-64 */
    uint32_t prob = probs[index];
    uint32_t bound = (rd->range >> kNumBitModelTotalBits) * prob;
    int bit;
    if (rd->code < bound) {
        rd->range = bound;
        probs[index] = (uint16_t)(prob + ((kBitModelTotal - prob) >> kNumMoveBits));
        bit = 0;
    } else {
        rd->range -= bound;
        rd->code -= bound;
        probs[index] = (uint16_t)(prob - (prob >> kNumMoveBits));
        bit = 1;
    }
    if (rd->range < kTopValue) {
        rd->code = (rd->code << 8) | rd_read(rd);
        rd->range <<= 8;
    }
    return bit;
}
static uint32_t rd_direct(rc_dec *rd, int nbits) { /* :27-41 */
    uint32_t result = 0;
    for (int i = nbits; i != 0; i--) {
        rd->range >>= 1;
        uint32_t t = (rd->code - rd->range) >> 31;
        rd->code -= rd->range & (t - 1);
        result = (result << 1) | (1 - t);
        if (rd->range < kTopValue) {
            rd->code = (rd->code << 8) | rd_read(rd);
            rd->range <<= 8;
        }
    }
    return result;
}
static int btd_decode(rc_dec *rd, uint16_t *probs, int nbits) { /* BitTreeDecoder.java:19-25 */
    int m = 1;
    for (int bi = nbits; bi != 0; bi--) m = (m << 1) + rd_bit(rd, probs, m);
    return m - (1 << nbits);
}
static int btd_reverse(rc_dec *rd, uint16_t *probs, int nbits) { /* BitTreeDecoder.java:27-37, Decoder.java:13-23 */
    int m = 1, symbol = 0;
    for (int bi = 0; bi < nbits; bi++) {
        int bit = rd_bit(rd, probs, m);
        m = (m << 1) + bit;
        symbol |= bit << bi;
    }
    return symbol;
}

typedef struct {
    uint16_t choice[2];
    uint16_t low[kNumPosStatesMax][1 << kNumLowLenBits];
    uint16_t mid[kNumPosStatesMax][1 << kNumMidLenBits];
    uint16_t high[1 << kNumHighLenBits];
} len_dec;

static int len_decode(len_dec *ld, rc_dec *rd, int pos_state) { /* Decoder.java:48-59 */
    if (rd_bit(rd, ld->choice, 0) == 0) return btd_decode(rd, ld->low[pos_state], kNumLowLenBits);
    int symbol = kNumLowLenSymbols;
    if (rd_bit(rd, ld->choice, 1) == 0) symbol += btd_decode(rd, ld->mid[pos_state], kNumMidLenBits);
    else symbol += kNumMidLenSymbols + btd_decode(rd, ld->high, kNumHighLenBits);
    return symbol;
}

typedef struct {
    uint16_t is_match[kNumStates << kNumPosStatesBitsMax];
    uint16_t is_rep[kNumStates];
    uint16_t is_rep_g0[kNumStates];
    uint16_t is_rep_g1[kNumStates];
    uint16_t is_rep_g2[kNumStates];
    uint16_t is_rep0_long[kNumStates << kNumPosStatesBitsMax];
    uint16_t pos_slot[kNumLenToPosStates][1 << kNumPosSlotBits];
    uint16_t pos_decoders[kNumFullDistances - kEndPosModelIndex];
    uint16_t pos_align[1 << kNumAlignBits];
    len_dec len, rep_len;
} dec_model;

int lzo_decode(const uint8_t props[5], const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap,
               int64_t out_size, size_t *written, size_t *consumed) {
    if (written) *written = 0;
    if (consumed) *consumed = 0;
    /* SetDecoderProperties :303-318 */
    int val = props[0];
    int lc = val % 9;
    int remainder = val / 9;
    int lp = remainder % 5;
    int pb = remainder / 5;
    int32_t dictionary_size = 0;
    for (int i = 0; i < 4; i++) dictionary_size += (int32_t)((uint32_t)props[1 + i] << (i * 8));
    if (lc > kNumLitContextBitsMax || lp > 4 || pb > kNumPosStatesBitsMax) return 0; /* SetLcLpPb :172-182 */
    if (dictionary_size < 0) return 0;                                                /* SetDictionarySize :160-170 */
    int32_t dict_check = dictionary_size > 1 ? dictionary_size : 1;
    int pos_state_mask = (1 << pb) - 1;
    int lp_mask = (1 << lp) - 1;

    dec_model *m = (dec_model *)malloc(sizeof(dec_model));
    size_t nlit = 0x300u << (lc + lp);
    uint16_t *literal = (uint16_t *)malloc(sizeof(uint16_t) * nlit);
    if (!m || !literal) { free(m); free(literal); return -1; }
    init_probs((uint16_t *)m, sizeof(dec_model) / 2); /* Init :184-203 */
    init_probs(literal, nlit);
    rc_dec rd;
    rd_init(&rd, in, in_len);

    int state = 0;
    int32_t rep0 = 0, rep1 = 0, rep2 = 0, rep3 = 0;
    int64_t now_pos = 0;
    uint8_t prev_byte = 0;
    int ret = 1;
    while (out_size < 0 || now_pos < out_size) { /* :219-299 */
        int pos_state = (int)now_pos & pos_state_mask;
        if (rd_bit(&rd, m->is_match, (state << kNumPosStatesBitsMax) + pos_state) == 0) {
            uint16_t *probs = literal + 0x300u * (size_t)((((int)now_pos & lp_mask) << lc) + (prev_byte >> (8 - lc)));
            if (st_is_char(state)) { /* DecodeNormal :70-77 */
                int symbol = 1;
                do symbol = (symbol << 1) | rd_bit(&rd, probs, symbol);
                while (symbol < 0x100);
                prev_byte = (uint8_t)symbol;
            } else { /* DecodeWithMatchByte :79-95 */
                uint8_t match_byte = out[now_pos - rep0 - 1];
                int symbol = 1;
                do {
                    int match_bit = (match_byte >> 7) & 1;
                    match_byte = (uint8_t)(match_byte << 1);
                    int bit = rd_bit(&rd, probs, ((1 + match_bit) << 8) + symbol);
                    symbol = (symbol << 1) | bit;
                    if (match_bit != bit) {
                        while (symbol < 0x100) symbol = (symbol << 1) | rd_bit(&rd, probs, symbol);
                        break;
                    }
                } while (symbol < 0x100);
                prev_byte = (uint8_t)symbol;
            }
            if ((size_t)now_pos >= out_cap) { ret = -1; break; }
            out[now_pos] = prev_byte;
            state = st_lit(state);
            now_pos++;
        } else {
            int len;
            if (rd_bit(&rd, m->is_rep, state) == 1) { /* :233-259 */
                len = 0;
                if (rd_bit(&rd, m->is_rep_g0, state) == 0) {
                    if (rd_bit(&rd, m->is_rep0_long, (state << kNumPosStatesBitsMax) + pos_state) == 0) {
                        state = st_shortrep(state);
                        len = 1;
                    }
                } else {
                    int32_t distance;
                    if (rd_bit(&rd, m->is_rep_g1, state) == 0) {
                        distance = rep1;
                    } else {
                        if (rd_bit(&rd, m->is_rep_g2, state) == 0) {
                            distance = rep2;
                        } else {
                            distance = rep3;
                            rep3 = rep2;
                        }
                        rep2 = rep1;
                    }
                    rep1 = rep0;
                    rep0 = distance;
                }
                if (len == 0) {
                    len = len_decode(&m->rep_len, &rd, pos_state) + kMatchMinLen;
                    state = st_longrep(state);
                }
            } else { /* :260-286 */
                rep3 = rep2;
                rep2 = rep1;
                rep1 = rep0;
                len = kMatchMinLen + len_decode(&m->len, &rd, pos_state);
                state = st_match(state);
                int pos_slot = btd_decode(&rd, m->pos_slot[len_to_pos_state(len)], kNumPosSlotBits);
                if (pos_slot >= kStartPosModelIndex) {
                    int num_direct_bits = (pos_slot >> 1) - 1;
                    rep0 = (int32_t)((uint32_t)(2 | (pos_slot & 1)) << num_direct_bits);
                    if (pos_slot < kEndPosModelIndex) {
                        rep0 += btd_reverse(&rd, m->pos_decoders + (rep0 - pos_slot - 1), num_direct_bits);
                    } else {
                        rep0 = (int32_t)((uint32_t)rep0 + (rd_direct(&rd, num_direct_bits - kNumAlignBits) << kNumAlignBits));
                        rep0 = (int32_t)((uint32_t)rep0 + (uint32_t)btd_reverse(&rd, m->pos_align, kNumAlignBits));
                        if (rep0 < 0) {
                            if (rep0 == -1) break; /* end marker */
                            ret = 0;
                            break;
                        }
                    }
                } else {
                    rep0 = pos_slot;
                }
            }
            if ((int64_t)rep0 >= now_pos || rep0 >= dict_check) { /* :288-291 */
                ret = 0;
                break;
            }
            if ((size_t)(now_pos + len) > out_cap) { ret = -1; break; }
            /* OutWindow.CopyBlock :53-67, linear buffer (out doubles as the window) */
            const uint8_t *src = out + (now_pos - rep0 - 1);
            uint8_t *dst = out + now_pos;
            for (int k = 0; k < len; k++) dst[k] = src[k];
            now_pos += len;
            prev_byte = out[now_pos - 1];
        }
    }
    if (written) *written = (size_t)now_pos;
    if (consumed) *consumed = rd.in_pos;
    free(m);
    free(literal);
    return ret;
}

int lzo_decode_alone(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap, size_t *written) { /* LzmaAlone.java:220-239 */
    if (written) *written = 0;
    if (in_len < 13) return 0; /* "input .lzma file is too short" / "Can't read stream size" */
    uint64_t out_size = 0;
    for (int i = 0; i < 8; i++) out_size |= (uint64_t)in[5 + i] << (8 * i);
    return lzo_decode(in, in + 13, in_len - 13, out, out_cap, (int64_t)out_size, written, NULL);
}

/* ------------------------------------------------------------------ */
/* known-answer helpers                                                */
/* ------------------------------------------------------------------ */
size_t lzo_kat_rc_bits(const int *bits, int nbits, uint8_t *out, size_t cap) {
    pthread_once(&g_once, init_tables);
    rc_enc rc;
    uint16_t probs[kNumStates];
    rc_init(&rc, out, cap);
    init_probs(probs, kNumStates);
    for (int i = 0; i < nbits; i++) rc_encode(&rc, probs, 4, bits[i]);
    rc_flush(&rc);
    return rc.out_pos;
}
size_t lzo_kat_rc_direct(const int *v, const int *nbits, int ncalls, uint8_t *out, size_t cap) {
    rc_enc rc;
    rc_init(&rc, out, cap);
    for (int i = 0; i < ncalls; i++) rc_encode_direct(&rc, (uint32_t)v[i], nbits[i]);
    rc_flush(&rc);
    return rc.out_pos;
}
void lzo_kat_bittree_prices(int prices[8]) {
    pthread_once(&g_once, init_tables);
    uint8_t sink[16];
    rc_enc rc;
    uint16_t probs[8];
    rc_init(&rc, sink, sizeof sink);
    init_probs(probs, 8);
    bt_encode(&rc, probs, 3, 3);
    for (int i = 0; i < 8; i++) prices[i] = bt_price(probs, 3, i);
}
void lzo_kat_prob_prices(int table[512]) {
    pthread_once(&g_once, init_tables);
    for (int i = 0; i < 512; i++) table[i] = g_prob_prices[i];
}

/* ------------------------------------------------------------------ */
/* batch drivers: one block per task over a fixed pthread pool          */
/* ------------------------------------------------------------------ */
typedef struct {
    int is_encode;
    const lzo_props *props;
    const uint8_t *in;
    const uint64_t *in_off, *in_len;
    uint32_t n_blocks;
    uint8_t *out;
    const uint64_t *out_off, *out_cap;
    uint64_t *out_len;
    int32_t *status;
    int with_header;
    volatile uint32_t next;
    volatile int failed;
} batch_job;

static void *batch_worker(void *arg) {
    batch_job *j = (batch_job *)arg;
    for (;;) {
        uint32_t b = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (b >= j->n_blocks) break;
        if (j->is_encode) {
            size_t r = j->with_header
                ? lzo_encode_alone(j->props, j->in + j->in_off[b], j->in_len[b], j->out + j->out_off[b], j->out_cap[b])
                : lzo_encode(j->props, j->in + j->in_off[b], j->in_len[b], j->out + j->out_off[b], j->out_cap[b], NULL);
            if (r == (size_t)-1) { j->failed = 1; j->out_len[b] = 0; }
            else j->out_len[b] = r;
        } else {
            size_t w = 0;
            int s = lzo_decode_alone(j->in + j->in_off[b], j->in_len[b], j->out + j->out_off[b], j->out_cap[b], &w);
            j->out_len[b] = w;
            if (j->status) j->status[b] = s;
            if (s != 1) j->failed = 1;
        }
    }
    return NULL;
}

static int run_batch(batch_job *j, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 1024) threads = 1024;
    pthread_t *tids = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    int started = 0;
    for (int t = 0; t < threads; t++)
        if (pthread_create(&tids[t], NULL, batch_worker, j) == 0) started++; else break;
    if (started == 0) batch_worker(j);
    for (int t = 0; t < started; t++) pthread_join(tids[t], NULL);
    free(tids);
    return j->failed ? 1 : 0;
}

int lzo_encode_batch(const lzo_props *p, const uint8_t *in, const uint64_t *in_off, const uint64_t *in_len,
                     uint32_t n_blocks, uint8_t *out, const uint64_t *out_off, const uint64_t *out_cap,
                     uint64_t *out_len, int with_header, int threads) {
    batch_job j;
    memset(&j, 0, sizeof j);
    j.is_encode = 1; j.props = p; j.in = in; j.in_off = in_off; j.in_len = in_len; j.n_blocks = n_blocks;
    j.out = out; j.out_off = out_off; j.out_cap = out_cap; j.out_len = out_len; j.with_header = with_header;
    return run_batch(&j, threads);
}

int lzo_decode_batch(const uint8_t *in, const uint64_t *in_off, const uint64_t *in_len, uint32_t n_blocks,
                     uint8_t *out, const uint64_t *out_off, const uint64_t *out_cap, uint64_t *out_len,
                     int32_t *status, int threads) {
    batch_job j;
    memset(&j, 0, sizeof j);
    j.is_encode = 0; j.in = in; j.in_off = in_off; j.in_len = in_len; j.n_blocks = n_blocks;
    j.out = out; j.out_off = out_off; j.out_cap = out_cap; j.out_len = out_len; j.status = status;
    return run_batch(&j, threads);
}
