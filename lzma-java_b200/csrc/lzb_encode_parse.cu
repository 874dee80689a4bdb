// lzb_encode_parse.cu -- optimal parse + price tables + range encoder, one
// warp per block (LZMA/Encoder.java:275-1125, LenEncoder.java,
// LenPriceTableEncoder.java, LiteralEncoder.java, RangeCoder/RangeEncoder.java,
// BitTreeEncoder.java, ProbPrices.java of rfalke/lzma-java).
//
// The match finder has already run (lzb_encode_mf.cu): ReadMatchDistances
// reads the position's list from global memory and Skip is a cursor bump.
// What remains is the strictly serial chain  parse chunk -> emit chunk ->
// parse next chunk with the adapted probabilities  (SURVEY.md section 3.1),
// kept per warp with the whole model and every price table in the warp's
// private slice of shared memory:
//   probabilities  pb-strided layout of lzb_common.cuh (7 320 u16 at lc3 lp0 pb2)
//   prices         u16 tables (a price never exceeds 10 * 576)
//   match list     the current position's pairs (the parser truncates them in
//                  place, Encoder.java:737-743)
//   _optimum[]     32-byte packed nodes in global memory (L1/L2 resident)
// Every relaxation keeps the reference's order and comparison (strict <, one
// <= for the short rep, App. A #8) so ties resolve identically.
#include "lzb_encode.cuh"

namespace lzb {

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr int kNumOpts = 1 << 12;           // Encoder.java:19
constexpr uint32_t kInfinityPrice = 0xFFFFFFF;  // Encoder.java:22
constexpr int kNumBitPriceShiftBits = 6;    // ProbPrices.java:6

// ---- tables shared by the CTA ----------------------------------------------
struct CtaTables {
    uint16_t prob_prices[512];  // ProbPrices.java:8-18
    uint8_t fast_pos[2048];     // Encoder.java:24-41
};

__device__ void init_cta_tables(CtaTables* t) {
    for (int j = threadIdx.x; j < 512; j += blockDim.x) {
        uint32_t v = 0;  // entry 0 is never written by the reference and stays 0
        if (j > 0) {
            const int hb = 31 - __clz(j);  // j in [2^hb, 2^(hb+1)): i = 8 - hb, end = 2^(hb+1)
            const int i = 8 - hb;
            v = ((uint32_t)i << kNumBitPriceShiftBits) + ((((1u << (hb + 1)) - (uint32_t)j) << kNumBitPriceShiftBits) >> hb);
        }
        t->prob_prices[j] = (uint16_t)v;
    }
    for (int c = threadIdx.x; c < 2048; c += blockDim.x) {
        // slot s >= 2 covers k = 2^((s>>1)-1) consecutive values starting at c0(s)
        uint32_t s;
        if (c < 2) {
            s = (uint32_t)c;
        } else {
            const int hb = 31 - __clz(c);
            s = (uint32_t)(2 * hb) + (((uint32_t)c >> (hb - 1)) & 1u);
        }
        t->fast_pos[c] = (uint8_t)s;
    }
    __syncthreads();
}

// ---- packed _optimum node (Optimal.java:4-33) ------------------------------
struct __align__(16) OptNode {
    uint32_t price;
    uint32_t back_prev;   // 0xFFFFFFFF = literal ("-1")
    uint32_t back_prev2;
    uint32_t link;        // pos_prev | pos_prev2 << 12 | state << 24 | prev1_is_char << 28 | prev2 << 29
    uint32_t backs[4];
};
__device__ __forceinline__ uint32_t mk_link(uint32_t pos_prev, uint32_t pos_prev2, bool p1, bool p2) {
    return pos_prev | (pos_prev2 << 12) | ((uint32_t)p1 << 28) | ((uint32_t)p2 << 29);
}
__device__ __forceinline__ uint32_t ln_pos_prev(uint32_t l) { return l & 0xFFF; }
__device__ __forceinline__ uint32_t ln_pos_prev2(uint32_t l) { return (l >> 12) & 0xFFF; }
__device__ __forceinline__ int ln_state(uint32_t l) { return (int)((l >> 24) & 0xF); }
__device__ __forceinline__ bool ln_prev1(uint32_t l) { return (l >> 28) & 1; }
__device__ __forceinline__ bool ln_prev2(uint32_t l) { return (l >> 29) & 1; }

// ---- range encoder (RangeEncoder.java:23-87) -------------------------------
struct RangeEnc {
    uint64_t low;
    uint32_t range;
    uint32_t cache_size;
    uint32_t cache;
    uint8_t* out;
    uint64_t pos, cap;

    __device__ __forceinline__ void init(uint8_t* o, uint64_t c) {
        low = 0;
        range = 0xFFFFFFFFu;
        cache_size = 1;
        cache = 0;
        out = o;
        pos = 0;
        cap = c;
    }
    __device__ __forceinline__ void put(uint32_t b) {
        if (pos < cap) out[pos] = (uint8_t)b;
        pos++;
    }
    __device__ __noinline__ void shift_low() {
        const uint32_t low_hi = (uint32_t)(low >> 32);
        if (low_hi != 0 || low < 0xFF000000ull) {
            uint32_t temp = cache;
            do {
                put(temp + low_hi);
                temp = 0xFF;
            } while (--cache_size != 0);
            cache = ((uint32_t)low) >> 24;
        }
        cache_size++;
        low = (low & 0xFFFFFF) << 8;
    }
    __device__ __forceinline__ void encode(uint16_t* prob, uint32_t bit) {
        const uint32_t p = *prob;
        const uint32_t bound = (range >> kNumBitModelTotalBits) * p;
        if (bit == 0) {
            range = bound;
            *prob = (uint16_t)(p + ((kBitModelTotal - p) >> kNumMoveBits));
        } else {
            low += bound;
            range -= bound;
            *prob = (uint16_t)(p - (p >> kNumMoveBits));
        }
        if (range < kTopValue) {
            range <<= 8;
            shift_low();
        }
    }
    __device__ __forceinline__ void direct(uint32_t v, int nbits) {
        for (int i = nbits - 1; i >= 0; i--) {
            range >>= 1;
            if ((v >> i) & 1) low += range;
            if (range < kTopValue) {
                range <<= 8;
                shift_low();
            }
        }
    }
    __device__ __forceinline__ void flush() {
        for (int i = 0; i < 5; i++) shift_low();
    }
    // BitTreeEncoder.encode / ReverseEncode (BitTreeEncoder.java:18-36, Encoder.java:196-205)
    __device__ __forceinline__ void tree(uint16_t* probs, int nbits, uint32_t symbol) {
        uint32_t m = 1;
        for (int bi = nbits; bi != 0;) {
            bi--;
            const uint32_t bit = (symbol >> bi) & 1;
            encode(probs + m, bit);
            m = (m << 1) | bit;
        }
    }
    __device__ __forceinline__ void reverse(uint16_t* probs, int nbits, uint32_t symbol) {
        uint32_t m = 1;
        for (int i = 0; i < nbits; i++) {
            const uint32_t bit = symbol & 1;
            encode(probs + m, bit);
            m = (m << 1) | bit;
            symbol >>= 1;
        }
    }
};

// ---- everything one stream needs -------------------------------------------
struct Enc {
    // tables / shared memory
    const CtaTables* T;
    uint16_t* model;        // fixed part of the probability model (shared)
    uint16_t* lit;          // literal coders (shared, or global when they do not fit)
    uint16_t* dist_prices;  // [4][128]  _distancesPrices
    uint16_t* slot_prices;  // [4][64]   _posSlotPrices
    uint16_t* align_prices; // [16]
    uint16_t* len_prices;   // [2][1<<pb][table_size]
    int32_t* len_counters;  // [2][16]
    uint32_t* md;           // current match list, len << 23 | distance
    OptNode* opt;           // [kNumOpts]
    ModelLayout L;
    // stream
    const uint8_t* data;
    uint32_t n;
    const uint32_t* idx;    // 1-based
    const uint32_t* pairs;
    // parameters
    int lc, lp, pb, fb, table_size, dist_table_size;
    uint32_t pos_mask, lp_mask;
    bool eos;
    // state (Encoder.java:132-181)
    RangeEnc rc;
    uint32_t m;             // match-finder cursor, 0-based (== _pos - 1 of the reference's InWindow)
    int state;
    uint32_t prev_byte;
    uint32_t rep_dist[4];
    uint32_t reps[4];
    uint32_t rep_lens[4];
    int num_pairs;
    int additional_offset;
    int opt_end, opt_cur;
    bool longest_found;
    int longest_len;
    int match_price_count, align_price_count;
    uint32_t now_pos;

    // ---- prices (ProbPrices.java:23-37) ----
    __device__ __forceinline__ uint32_t price_bit(uint32_t prob, uint32_t bit) const {
        return T->prob_prices[(((prob - bit) ^ (0u - bit)) & (kBitModelTotal - 1)) >> 2];
    }
    __device__ __forceinline__ uint32_t price0(uint32_t prob) const { return T->prob_prices[prob >> 2]; }
    __device__ __forceinline__ uint32_t price1(uint32_t prob) const { return T->prob_prices[(kBitModelTotal - prob) >> 2]; }
    __device__ __forceinline__ uint32_t tree_price(const uint16_t* probs, int nbits, uint32_t symbol) const {  // BitTreeEncoder.java:38-48
        uint32_t price = 0, mm = 1;
        for (int bi = nbits; bi != 0;) {
            bi--;
            const uint32_t bit = (symbol >> bi) & 1;
            price += price_bit(probs[mm], bit);
            mm = (mm << 1) + bit;
        }
        return price;
    }
    __device__ __forceinline__ uint32_t reverse_price(const uint16_t* probs, int nbits, uint32_t symbol) const {  // :50-60
        uint32_t price = 0, mm = 1;
        for (int i = nbits; i != 0; i--) {
            const uint32_t bit = symbol & 1;
            symbol >>= 1;
            price += price_bit(probs[mm], bit);
            mm = (mm << 1) | bit;
        }
        return price;
    }
    __device__ __forceinline__ int pos_slot(uint32_t pos) const {  // Encoder.java:86-94
        if (pos < (1u << 11)) return T->fast_pos[pos];
        if (pos < (1u << 21)) return T->fast_pos[pos >> 10] + 20;
        return T->fast_pos[pos >> 20] + 40;
    }
    __device__ __forceinline__ int pos_slot2(uint32_t pos) const {  // :96-104
        if (pos < (1u << 17)) return T->fast_pos[pos >> 6] + 12;
        if (pos < (1u << 27)) return T->fast_pos[pos >> 16] + 32;
        return T->fast_pos[pos >> 26] + 52;
    }

    // ---- probability addressing ----
    __device__ __forceinline__ uint16_t* p_is_match(int st, uint32_t ps) const { return model + L.is_match + (st << pb) + ps; }
    __device__ __forceinline__ uint16_t* p_is_rep0_long(int st, uint32_t ps) const { return model + L.is_rep0_long + (st << pb) + ps; }
    __device__ __forceinline__ uint16_t* p_is_rep(int st) const { return model + L.is_rep + st; }
    __device__ __forceinline__ uint16_t* p_is_rep_g0(int st) const { return model + L.is_rep_g0 + st; }
    __device__ __forceinline__ uint16_t* p_is_rep_g1(int st) const { return model + L.is_rep_g1 + st; }
    __device__ __forceinline__ uint16_t* p_is_rep_g2(int st) const { return model + L.is_rep_g2 + st; }
    __device__ __forceinline__ uint16_t* lit_coder(uint32_t pos, uint32_t prev) const {  // LiteralEncoder.java:93-95
        return lit + 0x300u * (((pos & lp_mask) << lc) + (prev >> (8 - lc)));
    }

    // ---- window (InWindow.java:115-138, whole block resident) ----
    __device__ __forceinline__ uint32_t byte_at(int index) const { return data[m + index]; }
    __device__ __forceinline__ int avail() const { return (int)(n - m); }
    __device__ int match_len(int index, uint32_t distance, int limit) const {
        const uint32_t s = m + index;
        if (s + limit > n) limit = (int)(n - s);
        const uint8_t* a = data + s;
        const uint8_t* b = a - distance - 1;
        int i = 0;
        while (i < limit && a[i] == b[i]) i++;
        return i;
    }

    // ---- match list ----
    __device__ __forceinline__ int md_len(int i) const { return (int)(md[i] >> kPairDistBits); }
    __device__ __forceinline__ uint32_t md_dist(int i) const { return md[i] & kPairDistMask; }

    __device__ int read_match_distances() {  // Encoder.java:275-287
        const uint32_t off = idx[m + 1];
        int cnt = 0;
        if (off != kMfEmpty) {
            cnt = (int)pairs[off];
            for (int i = 0; i < cnt; i++) md[i] = pairs[off + 1 + i];
        }
        num_pairs = cnt;
        m++;  // fillMatches advanced the window
        int length = 0;
        if (cnt > 0) {
            length = md_len(cnt - 1);
            if (length == fb) length += match_len(length - 1, md_dist(cnt - 1), kMatchMaxLen - length);
        }
        additional_offset++;
        return length;
    }
    __device__ __forceinline__ void move_pos(int num) {  // :289-294; Skip is free, the trees are already built
        if (num > 0) {
            m += num;
            additional_offset += num;
        }
    }

    // ---- length coder (LenEncoder.java, LenPriceTableEncoder.java) ----
    __device__ __forceinline__ uint32_t len_price(int which, int symbol, uint32_t ps) const {
        return len_prices[((which << pb) + ps) * table_size + symbol];
    }
    __device__ void len_update_table(int which, uint32_t ps) {  // LenEncoder.SetPrices :50-71 + UpdateTable :20-23
        const uint16_t* lp_ = model + (which ? L.rep_len : L.len);
        uint16_t* prices = len_prices + ((which << pb) + ps) * table_size;
        const uint32_t a0 = price0(lp_[0]), a1 = price1(lp_[0]);
        const uint32_t b0 = a1 + price0(lp_[1]), b1 = a1 + price1(lp_[1]);
        int i = 0;
        for (; i < kNumLowLenSymbols && i < table_size; i++) prices[i] = (uint16_t)(a0 + tree_price(lp_ + len_low(pb, ps), kNumLowLenBits, i));
        for (; i < kNumLowLenSymbols + kNumMidLenSymbols && i < table_size; i++)
            prices[i] = (uint16_t)(b0 + tree_price(lp_ + len_mid(pb, ps), kNumMidLenBits, i - kNumLowLenSymbols));
        for (; i < table_size; i++)
            prices[i] = (uint16_t)(b1 + tree_price(lp_ + len_high(pb), kNumHighLenBits, i - kNumLowLenSymbols - kNumMidLenSymbols));
        len_counters[which * 16 + ps] = table_size;
    }
    __device__ void len_encode(int which, uint32_t symbol, uint32_t ps) {  // LenEncoder.encode :33-48 + LenPriceTableEncoder.encode :32-37
        uint16_t* lp_ = model + (which ? L.rep_len : L.len);
        if (symbol < kNumLowLenSymbols) {
            rc.encode(lp_ + 0, 0);
            rc.tree(lp_ + len_low(pb, ps), kNumLowLenBits, symbol);
        } else {
            rc.encode(lp_ + 0, 1);
            if (symbol < kNumLowLenSymbols + kNumMidLenSymbols) {
                rc.encode(lp_ + 1, 0);
                rc.tree(lp_ + len_mid(pb, ps), kNumMidLenBits, symbol - kNumLowLenSymbols);
            } else {
                rc.encode(lp_ + 1, 1);
                rc.tree(lp_ + len_high(pb), kNumHighLenBits, symbol - kNumLowLenSymbols - kNumMidLenSymbols);
            }
        }
        if (--len_counters[which * 16 + ps] == 0) len_update_table(which, ps);
    }

    // ---- literal coder (LiteralEncoder.java:17-64) ----
    __device__ void lit_encode(uint16_t* probs, uint32_t symbol) {
        uint32_t context = 1;
        for (int i = 7; i >= 0; i--) {
            const uint32_t bit = (symbol >> i) & 1;
            rc.encode(probs + context, bit);
            context = (context << 1) | bit;
        }
    }
    __device__ void lit_encode_matched(uint16_t* probs, uint32_t match_byte, uint32_t symbol) {
        uint32_t context = 1;
        bool same = true;
        for (int i = 7; i >= 0; i--) {
            const uint32_t bit = (symbol >> i) & 1;
            uint32_t st = context;
            if (same) {
                const uint32_t match_bit = (match_byte >> i) & 1;
                st += (1 + match_bit) << 8;
                same = (match_bit == bit);
            }
            rc.encode(probs + st, bit);
            context = (context << 1) | bit;
        }
    }
    __device__ uint32_t lit_price(const uint16_t* probs, bool match_mode, uint32_t match_byte, uint32_t symbol) const {
        uint32_t price = 0, context = 1;
        int i = 7;
        if (match_mode) {
            for (; i >= 0; i--) {
                const uint32_t match_bit = (match_byte >> i) & 1;
                const uint32_t bit = (symbol >> i) & 1;
                price += price_bit(probs[((1 + match_bit) << 8) + context], bit);
                context = (context << 1) | bit;
                if (match_bit != bit) {
                    i--;
                    break;
                }
            }
        }
        for (; i >= 0; i--) {
            const uint32_t bit = (symbol >> i) & 1;
            price += price_bit(probs[context], bit);
            context = (context << 1) | bit;
        }
        return price;
    }

    // ---- rep / match prices (Encoder.java:296-333) ----
    __device__ __forceinline__ uint32_t rep_len1_price(int st, uint32_t ps) const {
        return price0(*p_is_rep_g0(st)) + price0(*p_is_rep0_long(st, ps));
    }
    __device__ uint32_t pure_rep_price(int rep_index, int st, uint32_t ps) const {
        uint32_t price;
        if (rep_index == 0) {
            price = price0(*p_is_rep_g0(st));
            price += price1(*p_is_rep0_long(st, ps));
        } else {
            price = price1(*p_is_rep_g0(st));
            if (rep_index == 1) {
                price += price0(*p_is_rep_g1(st));
            } else {
                price += price1(*p_is_rep_g1(st));
                price += price_bit(*p_is_rep_g2(st), rep_index - 2);
            }
        }
        return price;
    }
    __device__ __forceinline__ uint32_t rep_price(int rep_index, int len, int st, uint32_t ps) const {
        return len_price(1, len - kMatchMinLen, ps) + pure_rep_price(rep_index, st, ps);
    }
    __device__ __forceinline__ uint32_t pos_len_price(uint32_t pos, int len, uint32_t ps) const {
        uint32_t price;
        const int lps = len_to_pos_state(len);
        if (pos < kNumFullDistances)
            price = dist_prices[lps * kNumFullDistances + pos];
        else
            price = (uint32_t)slot_prices[(lps << kNumPosSlotBits) + pos_slot2(pos)] + align_prices[pos & kAlignMask];
        return price + len_price(0, len - kMatchMinLen, ps);
    }

    // ---- price table refresh (Encoder.java:1087-1125) ----
    __device__ void fill_distances_prices() {
        uint32_t temp[kNumFullDistances];
        for (int i = kStartPosModelIndex; i < kNumFullDistances; i++) {
            const int slot = pos_slot(i);
            const int footer = (slot >> 1) - 1;
            const int base = (2 | (slot & 1)) << footer;
            temp[i] = reverse_price(model + L.pos_dec + base - slot - 1, footer, i - base);
        }
        for (int lps = 0; lps < kNumLenToPosStates; lps++) {
            const uint16_t* enc = model + L.pos_slot + (lps << kNumPosSlotBits);
            const int st = lps << kNumPosSlotBits;
            int slot;
            for (slot = 0; slot < dist_table_size; slot++) slot_prices[st + slot] = (uint16_t)tree_price(enc, kNumPosSlotBits, slot);
            for (slot = kEndPosModelIndex; slot < dist_table_size; slot++)
                slot_prices[st + slot] += (uint16_t)((((slot >> 1) - 1) - kNumAlignBits) << kNumBitPriceShiftBits);
            const int st2 = lps * kNumFullDistances;
            int i;
            for (i = 0; i < kStartPosModelIndex; i++) dist_prices[st2 + i] = slot_prices[st + i];
            for (; i < kNumFullDistances; i++) dist_prices[st2 + i] = (uint16_t)(slot_prices[st + pos_slot(i)] + temp[i]);
        }
        match_price_count = 0;
    }
    __device__ void fill_align_prices() {
        for (int i = 0; i < kAlignTableSize; i++) align_prices[i] = (uint16_t)reverse_price(model + L.pos_align, kNumAlignBits, i);
        align_price_count = 0;
    }

    // ---- Backward (Encoder.java:335-362): returns back in *back_out, length as result ----
    __device__ int backward(int cur, uint32_t* back_out) {
        opt_end = cur;
        uint32_t pos_mem = ln_pos_prev(opt[cur].link);
        uint32_t back_mem = opt[cur].back_prev;
        do {
            const uint32_t lk = opt[cur].link;
            if (ln_prev1(lk)) {
                // MakeAsChar + PosPrev = posMem - 1
                opt[pos_mem].back_prev = 0xFFFFFFFFu;
                opt[pos_mem].link = mk_link(pos_mem - 1, 0, false, false);
                if (ln_prev2(lk)) {
                    opt[pos_mem - 1].link = mk_link(ln_pos_prev2(lk), 0, false, false);
                    opt[pos_mem - 1].back_prev = opt[cur].back_prev2;
                }
            }
            const uint32_t pos_prev = pos_mem;
            const uint32_t back_cur = back_mem;
            back_mem = opt[pos_prev].back_prev;
            pos_mem = ln_pos_prev(opt[pos_prev].link);
            opt[pos_prev].back_prev = back_cur;
            opt[pos_prev].link = (opt[pos_prev].link & ~0xFFFu) | (uint32_t)cur;  // only PosPrev changes; Prev1IsChar/Prev2 are read next round
            cur = (int)pos_prev;
        } while (cur > 0);
        opt_cur = (int)ln_pos_prev(opt[0].link);
        *back_out = opt[0].back_prev;
        return opt_cur;
    }

    // relaxation helper: strict '<' keeps the first candidate on ties (App. A #8)
    __device__ __forceinline__ void relax(int at, uint32_t price, uint32_t pos_prev, uint32_t back, bool p1, bool p2,
                                          uint32_t pos_prev2, uint32_t back2) {
        OptNode* o = &opt[at];
        if (price < o->price) {
            o->price = price;
            o->back_prev = back;
            o->back_prev2 = back2;
            o->link = mk_link(pos_prev, pos_prev2, p1, p2);
        }
    }

    __device__ int get_optimum(uint32_t position, uint32_t* back_out);
    __device__ void write_end_marker(uint32_t ps);
    __device__ void flush_stream(uint32_t now);
    __device__ bool encode_one();
    __device__ void run();
};

// getOptimum (Encoder.java:364-811).  Returns the length, *back_out = "pos" of PosAndLength
// (0xFFFFFFFF literal, 0..3 rep index, else distance + 4).
__device__ int Enc::get_optimum(uint32_t position, uint32_t* back_out) {
    if (opt_end != opt_cur) {  // :365-370
        const uint32_t lk = opt[opt_cur].link;
        const int len_res = (int)ln_pos_prev(lk) - opt_cur;
        *back_out = opt[opt_cur].back_prev;
        opt_cur = (int)ln_pos_prev(lk);
        return len_res;
    }
    opt_cur = 0;
    opt_end = 0;

    int len_main;
    if (longest_found) {
        len_main = longest_len;
        longest_found = false;
    } else {
        len_main = read_match_distances();
    }
    int num_distance_pairs = num_pairs;

    int num_avail = avail() + 1;
    if (num_avail < 2) {
        *back_out = 0xFFFFFFFFu;
        return 1;
    }
    if (num_avail > kMatchMaxLen) num_avail = kMatchMaxLen;

    int rep_max_index = 0;
    for (int i = 0; i < kNumRepDistances; i++) {  // :393-399
        reps[i] = rep_dist[i];
        rep_lens[i] = (uint32_t)match_len(-1, reps[i], kMatchMaxLen);
        if (rep_lens[i] > rep_lens[rep_max_index]) rep_max_index = i;
    }
    if ((int)rep_lens[rep_max_index] >= fb) {  // :400-404
        const int len_res = (int)rep_lens[rep_max_index];
        *back_out = (uint32_t)rep_max_index;
        move_pos(len_res - 1);
        return len_res;
    }
    if (len_main >= fb) {  // :406-410
        *back_out = md_dist(num_distance_pairs - 1) + kNumRepDistances;
        move_pos(len_main - 1);
        return len_main;
    }

    uint32_t current_byte = byte_at(-1);
    uint32_t match_byte = byte_at(0 - (int)rep_dist[0] - 1 - 1);

    if (len_main < 2 && current_byte != match_byte && rep_lens[rep_max_index] < 2) {  // :415-417
        *back_out = 0xFFFFFFFFu;
        return 1;
    }

    uint32_t pos_state = position & pos_mask;
    {
        OptNode* o0 = &opt[0];
        o0->link = mk_link(0, 0, false, false) | ((uint32_t)state << 24);
        o0->backs[0] = reps[0];
        o0->backs[1] = reps[1];
        o0->backs[2] = reps[2];
        o0->backs[3] = reps[3];
    }
    uint32_t price1_ = price0(*p_is_match(state, pos_state)) +
                       lit_price(lit_coder(position, prev_byte), !st_is_char(state), match_byte, current_byte);
    uint32_t back1 = 0xFFFFFFFFu;  // MakeAsChar

    uint32_t match_price = price1(*p_is_match(state, pos_state));
    uint32_t rep_match_price = match_price + price1(*p_is_rep(state));

    if (match_byte == current_byte) {  // :430-436
        const uint32_t short_rep_price = rep_match_price + rep_len1_price(state, pos_state);
        if (short_rep_price < price1_) {
            price1_ = short_rep_price;
            back1 = 0;  // MakeAsShortRep
        }
    }

    int len_end = len_main >= (int)rep_lens[rep_max_index] ? len_main : (int)rep_lens[rep_max_index];
    if (len_end < 2) {
        *back_out = back1;
        return 1;
    }
    opt[1].price = price1_;
    opt[1].back_prev = back1;
    opt[1].link = mk_link(0, 0, false, false);

    for (int len = len_end; len >= 2; len--) opt[len].price = kInfinityPrice;  // :451-455

    for (int i = 0; i < kNumRepDistances; i++) {  // :457-474
        int rep_len = (int)rep_lens[i];
        if (rep_len < 2) continue;
        const uint32_t price = rep_match_price + pure_rep_price(i, state, pos_state);
        do {
            relax(rep_len, price + len_price(1, rep_len - 2, pos_state), 0, (uint32_t)i, false, false, 0, 0);
        } while (--rep_len >= 2);
    }

    uint32_t normal_match_price = match_price + price0(*p_is_rep(state));

    {
        int len = rep_lens[0] >= 2 ? (int)rep_lens[0] + 1 : 2;  // :478-501
        if (len <= len_main) {
            int offs = 0;
            while (len > md_len(offs)) offs++;
            for (;; len++) {
                const uint32_t distance = md_dist(offs);
                relax(len, normal_match_price + pos_len_price(distance, len, pos_state), 0, distance + kNumRepDistances,
                      false, false, 0, 0);
                if (len == md_len(offs)) {
                    offs++;
                    if (offs == num_distance_pairs) break;
                }
            }
        }
    }

    int cur = 0;
    for (;;) {  // :505-810
        cur++;
        if (cur == len_end) return backward(cur, back_out);
        int new_len = read_match_distances();
        num_distance_pairs = num_pairs;
        if (new_len >= fb) {
            longest_len = new_len;
            longest_found = true;
            return backward(cur, back_out);
        }
        position++;
        OptNode* oc = &opt[cur];
        const uint32_t clink = oc->link;
        uint32_t pos_prev = ln_pos_prev(clink);
        int st;
        if (ln_prev1(clink)) {  // :520-535
            pos_prev--;
            if (ln_prev2(clink)) {
                st = ln_state(opt[ln_pos_prev2(clink)].link);
                if (oc->back_prev2 < kNumRepDistances) st = st_longrep(st);
                else st = st_match(st);
            } else {
                st = ln_state(opt[pos_prev].link);
            }
            st = st_lit(st);
        } else {
            st = ln_state(opt[pos_prev].link);
        }
        if (pos_prev == (uint32_t)cur - 1) {  // :536-541
            if (oc->back_prev == 0) st = st_shortrep(st);
            else st = st_lit(st);
        } else {  // :542-585
            uint32_t pos;
            if (ln_prev1(clink) && ln_prev2(clink)) {
                pos_prev = ln_pos_prev2(clink);
                pos = oc->back_prev2;
                st = st_longrep(st);
            } else {
                pos = oc->back_prev;
                if (pos < kNumRepDistances) st = st_longrep(st);
                else st = st_match(st);
            }
            const OptNode* o = &opt[pos_prev];
            const uint32_t b0 = o->backs[0], b1 = o->backs[1], b2 = o->backs[2], b3 = o->backs[3];
            if (pos < kNumRepDistances) {
                if (pos == 0) { reps[0] = b0; reps[1] = b1; reps[2] = b2; reps[3] = b3; }
                else if (pos == 1) { reps[0] = b1; reps[1] = b0; reps[2] = b2; reps[3] = b3; }
                else if (pos == 2) { reps[0] = b2; reps[1] = b0; reps[2] = b1; reps[3] = b3; }
                else { reps[0] = b3; reps[1] = b0; reps[2] = b1; reps[3] = b2; }
            } else {
                reps[0] = pos - kNumRepDistances;
                reps[1] = b0;
                reps[2] = b1;
                reps[3] = b2;
            }
        }
        oc->link = (clink & ~(0xFu << 24)) | ((uint32_t)st << 24);
        oc->backs[0] = reps[0];
        oc->backs[1] = reps[1];
        oc->backs[2] = reps[2];
        oc->backs[3] = reps[3];
        const uint32_t cur_price = oc->price;

        current_byte = byte_at(-1);
        match_byte = byte_at(0 - (int)reps[0] - 1 - 1);
        pos_state = position & pos_mask;

        const uint32_t cur_and1_price = cur_price + price0(*p_is_match(st, pos_state)) +
                                        lit_price(lit_coder(position, byte_at(-2)), !st_is_char(st), match_byte, current_byte);

        OptNode* next = &opt[cur + 1];
        bool next_is_char = false;
        if (cur_and1_price < next->price) {  // :606-611
            next->price = cur_and1_price;
            next->back_prev = 0xFFFFFFFFu;
            next->link = mk_link((uint32_t)cur, 0, false, false);
            next_is_char = true;
        }

        match_price = cur_price + price1(*p_is_match(st, pos_state));
        rep_match_price = match_price + price1(*p_is_rep(st));

        if (match_byte == current_byte && !(ln_pos_prev(next->link) < (uint32_t)cur && next->back_prev == 0)) {  // :616-625
            const uint32_t short_rep_price = rep_match_price + rep_len1_price(st, pos_state);
            if (short_rep_price <= next->price) {
                next->price = short_rep_price;
                next->back_prev = 0;
                next->link = mk_link((uint32_t)cur, 0, false, false);
                next_is_char = true;
            }
        }

        int num_avail_full = avail() + 1;  // :627-636
        if (kNumOpts - 1 - cur < num_avail_full) num_avail_full = kNumOpts - 1 - cur;
        num_avail = num_avail_full;
        if (num_avail < 2) continue;
        if (num_avail > fb) num_avail = fb;

        if (!next_is_char && match_byte != current_byte) {  // :637-665  literal + rep0
            const int t = num_avail_full - 1 < fb ? num_avail_full - 1 : fb;
            const int len_test2 = match_len(0, reps[0], t);
            if (len_test2 >= 2) {
                const int state2 = st_lit(st);
                const uint32_t ps_next = (position + 1) & pos_mask;
                const uint32_t next_rep_match_price = cur_and1_price + price1(*p_is_match(state2, ps_next)) + price1(*p_is_rep(state2));
                const int offset = cur + 1 + len_test2;
                while (len_end < offset) opt[++len_end].price = kInfinityPrice;
                relax(offset, next_rep_match_price + rep_price(0, len_test2, state2, ps_next), (uint32_t)cur + 1, 0, true,
                      false, 0, 0);
            }
        }

        int start_len = 2;

        for (int rep_index = 0; rep_index < kNumRepDistances; rep_index++) {  // :669-735
            int len_test = match_len(-1, reps[rep_index], num_avail);
            if (len_test < 2) continue;
            const int len_test_temp = len_test;
            const uint32_t rp = rep_match_price + pure_rep_price(rep_index, st, pos_state);
            do {
                while (len_end < cur + len_test) opt[++len_end].price = kInfinityPrice;
                relax(cur + len_test, rp + len_price(1, len_test - 2, pos_state), (uint32_t)cur, (uint32_t)rep_index, false,
                      false, 0, 0);
            } while (--len_test >= 2);
            len_test = len_test_temp;

            if (rep_index == 0) start_len = len_test + 1;

            if (len_test < num_avail_full) {  // :696-734  rep + literal + rep0
                const int t = num_avail_full - 1 - len_test < fb ? num_avail_full - 1 - len_test : fb;
                const int len_test2 = match_len(len_test, reps[rep_index], t);
                if (len_test2 >= 2) {
                    int state2 = st_longrep(st);
                    uint32_t ps_next = (position + len_test) & pos_mask;
                    const uint32_t cur_and_len_char_price =
                        rp + len_price(1, len_test - 2, pos_state) + price0(*p_is_match(state2, ps_next)) +
                        lit_price(lit_coder(position + len_test, byte_at(len_test - 1 - 1)), true,
                                  byte_at(len_test - 1 - ((int)reps[rep_index] + 1)), byte_at(len_test - 1));
                    state2 = st_lit(state2);
                    ps_next = (position + len_test + 1) & pos_mask;
                    const uint32_t next_match_price = cur_and_len_char_price + price1(*p_is_match(state2, ps_next));
                    const uint32_t next_rep_match_price = next_match_price + price1(*p_is_rep(state2));
                    const int offset = len_test + 1 + len_test2;
                    while (len_end < cur + offset) opt[++len_end].price = kInfinityPrice;
                    relax(cur + offset, next_rep_match_price + rep_price(0, len_test2, state2, ps_next),
                          (uint32_t)(cur + len_test + 1), 0, true, true, (uint32_t)cur, (uint32_t)rep_index);
                }
            }
        }

        if (new_len > num_avail) {  // :737-743
            new_len = num_avail;
            for (num_distance_pairs = 0; new_len > md_len(num_distance_pairs); num_distance_pairs++) {}
            md[num_distance_pairs] = ((uint32_t)new_len << kPairDistBits) | md_dist(num_distance_pairs);
            num_distance_pairs++;
        }
        if (new_len >= start_len) {  // :744-809
            normal_match_price = match_price + price0(*p_is_rep(st));
            while (len_end < cur + new_len) opt[++len_end].price = kInfinityPrice;

            int offs = 0;
            while (start_len > md_len(offs)) offs++;

            for (int len_test = start_len;; len_test++) {
                const uint32_t cur_back = md_dist(offs);
                uint32_t cur_and_len_price = normal_match_price + pos_len_price(cur_back, len_test, pos_state);
                relax(cur + len_test, cur_and_len_price, (uint32_t)cur, cur_back + kNumRepDistances, false, false, 0, 0);

                if (len_test == md_len(offs)) {
                    if (len_test < num_avail_full) {  // match + literal + rep0
                        const int t = num_avail_full - 1 - len_test < fb ? num_avail_full - 1 - len_test : fb;
                        const int len_test2 = match_len(len_test, cur_back, t);
                        if (len_test2 >= 2) {
                            int state2 = st_match(st);
                            uint32_t ps_next = (position + len_test) & pos_mask;
                            const uint32_t cur_and_len_char_price =
                                cur_and_len_price + price0(*p_is_match(state2, ps_next)) +
                                lit_price(lit_coder(position + len_test, byte_at(len_test - 1 - 1)), true,
                                          byte_at(len_test - ((int)cur_back + 1) - 1), byte_at(len_test - 1));
                            state2 = st_lit(state2);
                            ps_next = (position + len_test + 1) & pos_mask;
                            const uint32_t next_match_price = cur_and_len_char_price + price1(*p_is_match(state2, ps_next));
                            const uint32_t next_rep_match_price = next_match_price + price1(*p_is_rep(state2));
                            const int offset = len_test + 1 + len_test2;
                            while (len_end < cur + offset) opt[++len_end].price = kInfinityPrice;
                            cur_and_len_price = next_rep_match_price + rep_price(0, len_test2, state2, ps_next);
                            relax(cur + offset, cur_and_len_price, (uint32_t)(cur + len_test + 1), 0, true, true, (uint32_t)cur,
                                  cur_back + kNumRepDistances);
                        }
                    }
                    offs++;
                    if (offs == num_distance_pairs) break;
                }
            }
        }
    }
}

__device__ void Enc::write_end_marker(uint32_t ps) {  // Encoder.java:818-835
    if (!eos) return;
    rc.encode(p_is_match(state, ps), 1);
    rc.encode(p_is_rep(state), 0);
    state = st_match(state);
    len_encode(0, 0, ps);
    const uint32_t slot = (1u << kNumPosSlotBits) - 1;
    rc.tree(model + L.pos_slot + (len_to_pos_state(kMatchMinLen) << kNumPosSlotBits), kNumPosSlotBits, slot);
    const int footer_bits = 30;
    const uint32_t pos_reduced = (1u << footer_bits) - 1;
    rc.direct(pos_reduced >> kNumAlignBits, footer_bits - kNumAlignBits);
    rc.reverse(model + L.pos_align, kNumAlignBits, pos_reduced & kAlignMask);
}

__device__ void Enc::flush_stream(uint32_t now) {  // :837-841
    write_end_marker(now & pos_mask);
    rc.flush();
}

// encodeOne (:890-936) with its emitters (:938-1024); false once the stream is flushed
__device__ bool Enc::encode_one() {
    uint32_t back;
    const int len = get_optimum(now_pos, &back);
    const uint32_t ps = now_pos & pos_mask;
    if (len == 1 && back == 0xFFFFFFFFu) {
        rc.encode(p_is_match(state, ps), 0);
        const uint32_t cur_byte = byte_at(0 - additional_offset);  // encodeSingleByteLiteral :1007-1024
        uint16_t* sub = lit_coder(now_pos, prev_byte);
        if (st_is_char(state)) {
            lit_encode(sub, cur_byte);
        } else {
            const uint32_t mb = byte_at(0 - (int)rep_dist[0] - 1 - additional_offset);
            lit_encode_matched(sub, mb, cur_byte);
        }
        prev_byte = cur_byte;
        state = st_lit(state);
    } else {
        rc.encode(p_is_match(state, ps), 1);
        if (back < kNumRepDistances) {  // encodeARepetition :938-974
            rc.encode(p_is_rep(state), 1);
            if (back == 0) {
                rc.encode(p_is_rep_g0(state), 0);
                rc.encode(p_is_rep0_long(state, ps), len == 1 ? 0 : 1);
            } else {
                rc.encode(p_is_rep_g0(state), 1);
                if (back == 1) {
                    rc.encode(p_is_rep_g1(state), 0);
                } else {
                    rc.encode(p_is_rep_g1(state), 1);
                    rc.encode(p_is_rep_g2(state), back - 2);
                }
            }
            if (len == 1) {
                state = st_shortrep(state);
            } else {
                len_encode(1, len - kMatchMinLen, ps);
                state = st_longrep(state);
            }
            const uint32_t distance = rep_dist[back];
            if (back != 0) {
                for (int k = (int)back; k >= 1; k--) rep_dist[k] = rep_dist[k - 1];
                rep_dist[0] = distance;
            }
        } else {  // encodeAMatch :976-1005
            rc.encode(p_is_rep(state), 0);
            state = st_match(state);
            len_encode(0, len - kMatchMinLen, ps);
            const uint32_t pos = back - kNumRepDistances;
            const int slot = pos_slot(pos);
            rc.tree(model + L.pos_slot + (len_to_pos_state(len) << kNumPosSlotBits), kNumPosSlotBits, slot);
            if (slot >= kStartPosModelIndex) {
                const int footer_bits = (slot >> 1) - 1;
                const uint32_t base = (2u | (slot & 1)) << footer_bits;
                const uint32_t pos_reduced = pos - base;
                if (slot < kEndPosModelIndex) {
                    rc.reverse(model + L.pos_dec + base - slot - 1, footer_bits, pos_reduced);
                } else {
                    rc.direct(pos_reduced >> kNumAlignBits, footer_bits - kNumAlignBits);
                    rc.reverse(model + L.pos_align, kNumAlignBits, pos_reduced & kAlignMask);
                    align_price_count++;
                }
            }
            rep_dist[3] = rep_dist[2];
            rep_dist[2] = rep_dist[1];
            rep_dist[1] = rep_dist[0];
            rep_dist[0] = pos;
            match_price_count++;
        }
        prev_byte = byte_at(len - 1 - additional_offset);
    }
    additional_offset -= len;
    now_pos += len;
    if (additional_offset == 0) {
        if (match_price_count >= (1 << 7)) fill_distances_prices();
        if (align_price_count >= kAlignTableSize) fill_align_prices();
        if (avail() == 0) {
            flush_stream(now_pos);
            return false;
        }
    }
    return true;
}

// SetStreams + CodeOneBlock loop (Encoder.java:1046-1077, 843-888); probabilities already initialised
__device__ void Enc::run() {
    state = 0;
    prev_byte = 0;
    for (int i = 0; i < 4; i++) rep_dist[i] = 0;
    longest_found = false;
    opt_end = opt_cur = 0;
    additional_offset = 0;
    m = 0;
    now_pos = 0;
    num_pairs = 0;
    fill_distances_prices();
    fill_align_prices();
    for (int which = 0; which < 2; which++)
        for (uint32_t ps = 0; ps < (1u << pb); ps++) len_update_table(which, ps);

    if (avail() == 0) {
        flush_stream(0);
        return;
    }
    read_match_distances();  // first byte is always a plain literal (:860-878)
    rc.encode(p_is_match(state, 0), 0);
    state = st_lit(state);
    const uint32_t cur_byte = byte_at(0 - additional_offset);
    lit_encode(lit_coder(0, prev_byte), cur_byte);
    prev_byte = cur_byte;
    additional_offset--;
    now_pos++;
    if (avail() == 0) {
        flush_stream(now_pos);
        return;
    }
    while (encode_one()) {}
}

// ---- shared-memory slice of one warp ----------------------------------------
struct SliceLayout {
    uint32_t model, dist_prices, slot_prices, align_prices, len_prices, len_counters, md, total;  // byte offsets
    bool lit_in_smem;
};
__host__ __device__ inline SliceLayout make_slice(int lc, int lp, int pb, int fb) {
    SliceLayout s;
    const ModelLayout L = make_layout(lc, lp, pb);
    const uint32_t table = (uint32_t)(fb - 1);
    auto fixed_after = [&](uint32_t o) {
        SliceLayout r;
        r.dist_prices = o;  o += 512 * 2;
        r.slot_prices = o;  o += 256 * 2;
        r.align_prices = o; o += 16 * 2;
        r.len_prices = o;   o += ((2u << pb) * table * 2 + 3) & ~3u;
        r.len_counters = o; o += 32 * 4;
        r.md = o;           o += 276 * 4;
        r.total = o;
        return r;
    };
    uint32_t with_lit = (uint32_t)(L.n_fixed + L.n_literal) * 2;
    SliceLayout a = fixed_after((with_lit + 15) & ~15u);
    if (a.total <= kEncSliceBytes) {
        s = a;
        s.lit_in_smem = true;
    } else {
        s = fixed_after(((uint32_t)L.n_fixed * 2 + 15) & ~15u);
        s.lit_in_smem = false;
    }
    s.model = 0;
    return s;
}

__global__ void __launch_bounds__(kEncMaxWarps * 32, 1) lzb_parse_kernel(ParseArgs a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    CtaTables* tables = reinterpret_cast<CtaTables*>(smem_raw);
    init_cta_tables(tables);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warps = blockDim.x >> 5;
    uint8_t* slice = smem_raw + sizeof(CtaTables) + (size_t)warp * kEncSliceBytes;
    const size_t slot = (size_t)blockIdx.x * warps + warp;
    const ModelLayout L = make_layout(a.lc, a.lp, a.pb);
    const SliceLayout S = make_slice(a.lc, a.lp, a.pb, a.fb);
    uint16_t* model = reinterpret_cast<uint16_t*>(slice + S.model);
    uint16_t* lit = S.lit_in_smem ? model + L.literal : a.lit_scratch + slot * (size_t)L.n_literal;

    for (;;) {
        uint32_t b = 0;
        if (lane == 0) b = atomicAdd(a.ticket, 1u);
        b = __shfl_sync(kFull, b, 0);
        if (b >= a.mf.n_blocks) break;
        const uint32_t n = (uint32_t)a.mf.in_len[b];
        uint8_t* out = a.out + a.out_off[b];
        uint64_t cap = a.out_cap[b];
        uint64_t header = 0;
        if (a.with_header) {  // LzmaAlone.java:208-217
            if (cap >= LZB_KERNEL_HEADER) {
                if (lane < LZB_KERNEL_HEADER) {
                    uint32_t v;
                    if (lane == 0) v = (uint32_t)((a.pb * 5 + a.lp) * 9 + a.lc);
                    else if (lane < 5) v = ((uint32_t)a.dict_size >> (8 * (lane - 1))) & 0xFF;
                    else v = a.eos ? 0xFF : (uint32_t)(((uint64_t)n >> (8 * (lane - 5))) & 0xFF);
                    out[lane] = (uint8_t)v;
                }
                header = LZB_KERNEL_HEADER;
                out += LZB_KERNEL_HEADER;
                cap -= LZB_KERNEL_HEADER;
            } else {
                cap = 0;
                header = LZB_KERNEL_HEADER;
            }
        }
        // Encoder.Init (:247-273): every probability = 1024; price tables start from Java's zero-init (App. A #14)
        for (int i = lane; i < L.n_fixed; i += 32) model[i] = kProbInit;
        for (int i = lane; i < L.n_literal; i += 32) lit[i] = kProbInit;
        {
            uint32_t* z = reinterpret_cast<uint32_t*>(slice + S.dist_prices);
            const uint32_t words = (S.total - S.dist_prices) / 4;
            for (uint32_t i = lane; i < words; i += 32) z[i] = 0;
        }
        __syncwarp();
        if (lane == 0) {
            Enc e;
            e.T = tables;
            e.model = model;
            e.lit = lit;
            e.dist_prices = reinterpret_cast<uint16_t*>(slice + S.dist_prices);
            e.slot_prices = reinterpret_cast<uint16_t*>(slice + S.slot_prices);
            e.align_prices = reinterpret_cast<uint16_t*>(slice + S.align_prices);
            e.len_prices = reinterpret_cast<uint16_t*>(slice + S.len_prices);
            e.len_counters = reinterpret_cast<int32_t*>(slice + S.len_counters);
            e.md = reinterpret_cast<uint32_t*>(slice + S.md);
            e.opt = reinterpret_cast<OptNode*>(a.opt_scratch) + slot * (size_t)kNumOpts;
            e.L = L;
            e.data = a.mf.in + a.mf.in_off[b];
            e.n = n;
            e.idx = a.mf.idx + (size_t)b * a.mf.np;
            e.pairs = a.mf.pairs + (size_t)b * a.mf.pair_cap;
            e.lc = a.lc;
            e.lp = a.lp;
            e.pb = a.pb;
            e.fb = a.fb;
            e.table_size = a.fb + 1 - kMatchMinLen;
            e.dist_table_size = a.dist_table_size;
            e.pos_mask = (1u << a.pb) - 1;
            e.lp_mask = (1u << a.lp) - 1;
            e.eos = a.eos;
            e.match_price_count = 0;
            e.align_price_count = 0;
            e.rc.init(out, cap);
            e.run();
            a.out_len[b] = e.rc.pos > cap ? ~0ull : e.rc.pos + header;
        }
        __syncwarp();
    }
}

size_t parse_smem_bytes(int warps) { return sizeof(CtaTables) + (size_t)warps * kEncSliceBytes; }

bool parse_lit_in_smem(int lc, int lp, int pb, int fb) { return make_slice(lc, lp, pb, fb).lit_in_smem; }

size_t parse_opt_bytes_per_slot() { return sizeof(OptNode) * (size_t)kNumOpts; }

cudaError_t launch_parse(const ParseArgs& a, int grid, int warps, cudaStream_t st) {
    const size_t smem = parse_smem_bytes(warps);
    cudaError_t e = cudaFuncSetAttribute(lzb_parse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)parse_smem_bytes(kEncMaxWarps));
    if (e != cudaSuccess) return e;
    lzb_parse_kernel<<<grid, warps * 32, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace lzb
