// lzb_encode_parse.cu -- optimal parse + price tables + range encoder, one
// warp per block, warp-cooperative (LZMA/Encoder.java:275-1125,
// LenEncoder.java, LenPriceTableEncoder.java, LiteralEncoder.java,
// RangeCoder/RangeEncoder.java, BitTreeEncoder.java, ProbPrices.java of
// rfalke/lzma-java).
//
// The match finder has already run (lzb_encode_mf.cu): ReadMatchDistances
// reads the position's list from global memory and Skip is a cursor bump.
// What remains is the strictly serial chain  parse chunk -> emit chunk ->
// parse next chunk with the adapted probabilities  (SURVEY.md section 3.1).
// Within that chain the warp works as one:
//   * every lane holds the same copy of the scalar parser state (state, reps,
//     cursor, prices of the current node), so control flow is uniform;
//   * byte comparisons (InWindow.GetMatchLen) are done 32 bytes per step:
//     each lane loads one byte of the window and one byte per rep distance,
//     a ballot gives the equality bitmap, and both the rep length and the
//     "literal + rep0" continuation (Encoder.java:640,698,769) are run
//     lengths in that bitmap;
//   * the relaxation loops over lengths (Encoder.java:463-473, 484-500,
//     675-688, 755-808) give one length to each lane; candidates that can
//     reach the same node are applied in the reference's order with the
//     reference's comparison (strict <, one <= for the short rep), so ties
//     resolve identically (App. A #8);
//   * the four rep distances of a position belong to lanes 0..3, one each:
//     rep length, "rep + literal + rep0" length and the node range they need
//     come from ONE copy of the bitmap arithmetic (the kernel's hot path has
//     to fit a 32 KB instruction cache: DESIGN.md, parser);
//   * literal prices use eight lanes, one per bit; price-table refreshes use
//     one lane per table entry;
//   * a symbol is emitted as a batch: lane k adapts the k-th probability, the
//     range-coder fold (a serial chain) runs over shuffled (probability, bit)
//     pairs in every lane and lane 0 stores the bytes.
// Shared memory per warp: the probability model (pb-strided layout), all
// price tables as u16, the current match list, and a ring of >= 4 fb + 3
// packed 32-byte _optimum nodes.  A parse chunk may span 4095 positions, but a
// step only touches nodes [cur - 2fb - 1, cur + 2fb + 1]; older nodes are
// written back to global memory and Backward runs there when a chunk wrapped.
#include "lzb_encode.cuh"

namespace lzb {

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr int kNumOpts = 1 << 12;               // Encoder.java:19
constexpr uint32_t kInfinityPrice = 0xFFFFFFF;  // Encoder.java:22
constexpr int kNumBitPriceShiftBits = 6;        // ProbPrices.java:6
constexpr uint32_t kLit = 0xFFFFFFFFu;          // back value of a literal ("-1")

// ---- tables shared by the CTA ----------------------------------------------
struct CtaTables {
    uint16_t prob_prices[512];  // ProbPrices.java:8-18
    uint8_t fast_pos[2048];     // Encoder.java:24-41
};

__device__ void init_cta_tables(CtaTables* t) {
    #pragma unroll 1
    for (int j = threadIdx.x; j < 512; j += blockDim.x) {
        uint32_t v = 0;  // entry 0 is never written by the reference and stays 0
        if (j > 0) {
            const int hb = 31 - __clz(j);  // j in [2^hb, 2^(hb+1)): i = 8 - hb, end = 2^(hb+1)
            const int i = 8 - hb;
            v = ((uint32_t)i << kNumBitPriceShiftBits) + ((((1u << (hb + 1)) - (uint32_t)j) << kNumBitPriceShiftBits) >> hb);
        }
        t->prob_prices[j] = (uint16_t)v;
    }
    #pragma unroll 1
    for (int c = threadIdx.x; c < 2048; c += blockDim.x) {
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
    uint32_t back_prev;   // kLit = literal
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

// ---- range encoder (RangeEncoder.java:23-87) --------------------------------
// The coder state lives in the warp's shared-memory slice and the emission routine is a real
// (out-of-line) function: the parser reaches it from seven places, and seven inlined copies were
// 16 KB of a kernel whose hot path must fit a 32 KB instruction cache (parser).  It takes
// values and a shared-memory pointer only -- nothing of the caller's register state has its address
// taken.  A symbol is emitted as a batch: lane k passes the k-th binary decision (probability
// address + bit, or a direct bit); the routine adapts the probabilities -- all addresses of one
// symbol are distinct -- and folds the decisions into low/range in order, fetching each
// (old probability, bit) by shuffle.  Every lane runs the same fold; lane 0 stores the bytes.
struct __align__(16) RcState {
    uint64_t low;
    uint32_t range, cache_size;
    uint32_t cache, pos;  // pos = bytes produced so far (may run past cap: then nothing is stored)
    uint32_t cap, pad;
    uint8_t* out;
};

__device__ __forceinline__ void rc_init(RcState* rs, uint8_t* out, uint64_t cap, int lane) {  // RangeEncoder.java:23-29
    if (lane == 0) {
        rs->low = 0;
        rs->range = 0xFFFFFFFFu;
        rs->cache_size = 1;
        rs->cache = 0;
        rs->pos = 0;
        rs->cap = cap > 0xFFFFFFF0ull ? 0xFFFFFFF0u : (uint32_t)cap;
        rs->out = out;
    }
    __syncwarp();
}

// RangeEncoder.shiftLow (:73-87)
#define LZB_SHIFT_LOW()                                                            \
    do {                                                                           \
        const uint32_t low_hi = (uint32_t)(low >> 32);                             \
        if (low_hi != 0 || low < 0xFF000000ull) {                                  \
            uint32_t temp = cache;                                                 \
            _Pragma("unroll 1") do {                                               \
                if (lane == 0 && pos < cap) out[pos] = (uint8_t)(temp + low_hi);   \
                pos++;                                                             \
                temp = 0xFF;                                                       \
            } while (--cache_size != 0);                                           \
            cache = ((uint32_t)low) >> 24;                                         \
        }                                                                          \
        cache_size++;                                                              \
        low = (low & 0xFFFFFF) << 8;                                               \
    } while (0)

// Encode `cnt` (<= 32) binary decisions.  Lane k < cnt passes its decision: the probability and the bit.
// RangeEncoder.encode :38-54.  (Direct bits go through rc_direct: a test for them in this loop cost every decision
// six issue slots.)
__device__ __noinline__ void rc_batch(RcState* rs, int cnt, uint16_t* prob, uint32_t bit, int lane) {
    uint32_t p = 0;
    if (lane < cnt) {
        p = *prob;
        *prob = (uint16_t)(bit ? p - (p >> kNumMoveBits) : p + ((kBitModelTotal - p) >> kNumMoveBits));
    }
    const uint32_t w = p | (bit << 31);  // old probability | bit << 31
    uint64_t low = rs->low;
    uint32_t range = rs->range, cache_size = rs->cache_size, cache = rs->cache, pos = rs->pos;
    const uint32_t cap = rs->cap;
    uint8_t* out = rs->out;
#pragma unroll 1
    for (int k = 0; k < cnt; k++) {
        const uint32_t e = __shfl_sync(kFull, w, k);
        const uint32_t bound = (range >> kNumBitModelTotalBits) * (e & 0xFFFFu);
        if ((int32_t)e < 0) {
            low += bound;
            range -= bound;
        } else {
            range = bound;
        }
        if (range < kTopValue) {
            range <<= 8;
            LZB_SHIFT_LOW();
        }
    }
    if (lane == 0) {
        rs->low = low;
        rs->range = range;
        rs->cache_size = cache_size;
        rs->cache = cache;
        rs->pos = pos;
    }
    __syncwarp();
}

// RangeEncoder.encodeDirectBits (:56-67): the low `nbits` bits of `value`, most significant first; uniform arguments.
__device__ __noinline__ void rc_direct(RcState* rs, uint32_t value, int nbits, int lane) {
    uint64_t low = rs->low;
    uint32_t range = rs->range, cache_size = rs->cache_size, cache = rs->cache, pos = rs->pos;
    const uint32_t cap = rs->cap;
    uint8_t* out = rs->out;
#pragma unroll 1
    for (int i = nbits - 1; i >= 0; i--) {
        range >>= 1;
        if ((value >> i) & 1u) low += range;
        if (range < kTopValue) {
            range <<= 8;
            LZB_SHIFT_LOW();
        }
    }
    if (lane == 0) {
        rs->low = low;
        rs->range = range;
        rs->cache_size = cache_size;
        rs->cache = cache;
        rs->pos = pos;
    }
    __syncwarp();
}

__device__ __noinline__ void rc_flush(RcState* rs, int lane) {  // RangeEncoder.flush :31-36
    uint64_t low = rs->low;
    uint32_t cache_size = rs->cache_size, cache = rs->cache, pos = rs->pos;
    const uint32_t cap = rs->cap;
    uint8_t* out = rs->out;
#pragma unroll 1
    for (int i = 0; i < 5; i++) LZB_SHIFT_LOW();
    if (lane == 0) {
        rs->low = low;
        rs->cache_size = cache_size;
        rs->cache = cache;
        rs->pos = pos;
    }
    __syncwarp();
}

// What the out-of-line helpers (price-table refreshes) need to know about a stream; written once
// per stream into the warp's slice so that they take one pointer instead of the parser's registers.
struct __align__(16) WarpCtx {
    RcState rc;
    const uint16_t* prob_prices;  // CtaTables::prob_prices
    const uint8_t* fast_pos;
    uint16_t* model;
    uint16_t* dist_prices;
    uint16_t* slot_prices;
    uint16_t* align_prices;
    uint16_t* len_prices;
    int32_t* len_counters;
    int32_t pb, table_size, dist_table_size;
    int32_t off_len, off_rep_len, off_pos_slot, off_pos_dec, off_pos_align;
    unsigned long long* progress;         // ParseArgs::progress (null: nobody listens)
    uint32_t progress_in, progress_out;   // what this stream has reported so far
};

// ---- shared-memory slice of one warp ----------------------------------------
struct SliceLayout {
    uint32_t ctx, dist_prices, slot_prices, align_prices, len_prices, len_counters, md, md2, ring, total;  // byte offsets
    uint32_t ring_nodes;
    bool lit_in_smem;
};
// `with_lit`: the literal coders (0x300 << (lc + lp) probabilities) follow the fixed part of the model in
// shared memory; otherwise they live in global memory (ParseArgs::lit_scratch) and the slice shrinks
// by 12 KB at lc + lp = 3, which is what bounds the streams per SM.
__host__ __device__ inline SliceLayout make_slice(int lc, int lp, int pb, int fb, bool with_lit) {
    const ModelLayout L = make_layout(lc, lp, pb);
    const uint32_t table = (uint32_t)(fb - 1);
    // a DP step touches nodes [cur - 2fb - 1, cur + 2fb + 1]
    const uint32_t ring_nodes = ((uint32_t)(4 * fb + 3) + 7) & ~7u;
    SliceLayout s;
    uint32_t o = (uint32_t)(L.n_fixed + (with_lit ? L.n_literal : 0)) * 2;
    o = (o + 15) & ~15u;
    s.ctx = o;          o += (uint32_t)sizeof(WarpCtx);
    o = (o + 15) & ~15u;
    s.dist_prices = o;  o += 512 * 2;
    s.slot_prices = o;  o += 256 * 2;
    s.align_prices = o; o += 16 * 2;
    s.len_prices = o;   o += ((2u << pb) * table * 2 + 15) & ~15u;
    s.len_counters = o; o += 32 * 4;  // [2][16]
    // a position has at most fb - 1 pairs (lengths 2 .. fb, strictly increasing) and the parser may append one
    const uint32_t md_entries = ((uint32_t)fb + 1 + 31) & ~31u;
    s.md = o;           o += md_entries * 4;
    s.md2 = o;          o += md_entries * 4;
    o = (o + 15) & ~15u;
    s.ring = o;         o += ring_nodes * 32;
    s.total = o;
    s.ring_nodes = ring_nodes;
    s.lit_in_smem = with_lit;
    return s;
}

// InWindow.GetMatchLen (InWindow.java:120-134) from absolute position `s`, 32 bytes per round;
// every lane returns the same value.  Out of line: it is the rare continuation of a bitmap run.
__device__ __noinline__ int warp_match_len(const uint8_t* data, uint32_t n, int lane, uint32_t s, uint32_t distance, int limit) {
    if (limit > 0 && s + (uint32_t)limit > n) limit = (int)(n - s);
    const uint8_t* a = data + s;
    const uint8_t* b = a - distance - 1;
    int len = 0;
    #pragma unroll 1
    while (len < limit) {
        const int i = len + lane;
        const bool ok = i < limit;
        uint32_t x = 0, y = 1;
        if (ok) {
            x = a[i];
            y = b[i];
        }
        const unsigned neq = __ballot_sync(kFull, x != y);
        if (neq) return len + (__ffs(neq) - 1);
        len += 32;
    }
    return limit > 0 ? limit : 0;
}

// element i of a 4-array without dynamic indexing (which would push the array -- and the struct
// around it -- into local memory)
template <typename T>
__device__ __forceinline__ T sel4(const T (&a)[4], int i) {
    return i == 0 ? a[0] : (i == 1 ? a[1] : (i == 2 ? a[2] : a[3]));
}
template <typename T>
__device__ __forceinline__ void set4(T (&a)[4], int i, T v) {
    if (i == 0) a[0] = v;
    if (i == 1) a[1] = v;
    if (i == 2) a[2] = v;
    if (i == 3) a[3] = v;
}

// ---- price-table refreshes, out of line (rare: every fb-1 length symbols / 128 matches / 16 aligned
// distances) -- LenEncoder.SetPrices (LenEncoder.java:50-71) + LenPriceTableEncoder.UpdateTable (:20-23),
// Encoder.FillDistancesPrices / FillAlignPrices (Encoder.java:1087-1125).  One lane per table entry.
__device__ __forceinline__ uint32_t ctx_price_bit(const uint16_t* pp, uint32_t prob, uint32_t bit) {  // ProbPrices.java:23-37
    return pp[(((prob - bit) ^ (0u - bit)) & (kBitModelTotal - 1)) >> 2];
}
__device__ __forceinline__ uint32_t ctx_tree_price(const uint16_t* pp, const uint16_t* probs, int nbits, uint32_t symbol) {  // BitTreeEncoder.java:38-48
    uint32_t price = 0, mm = 1;
    #pragma unroll 1
    for (int bi = nbits; bi != 0;) {
        bi--;
        const uint32_t bit = (symbol >> bi) & 1;
        price += ctx_price_bit(pp, probs[mm], bit);
        mm = (mm << 1) + bit;
    }
    return price;
}
__device__ __forceinline__ uint32_t ctx_reverse_price(const uint16_t* pp, const uint16_t* probs, int nbits, uint32_t symbol) {  // :50-60
    uint32_t price = 0, mm = 1;
    #pragma unroll 1
    for (int i = nbits; i != 0; i--) {
        const uint32_t bit = symbol & 1;
        symbol >>= 1;
        price += ctx_price_bit(pp, probs[mm], bit);
        mm = (mm << 1) | bit;
    }
    return price;
}

__device__ __noinline__ void ctx_len_update_table(const WarpCtx* c, int which, uint32_t ps, int lane) {
    __syncwarp();
    const uint16_t* pp = c->prob_prices;
    const int pb = c->pb, table_size = c->table_size;
    const uint16_t* lp_ = c->model + (which ? c->off_rep_len : c->off_len);
    uint16_t* prices = c->len_prices + ((which << pb) + ps) * table_size;
    const uint32_t a0 = pp[lp_[0] >> 2], a1 = pp[(kBitModelTotal - lp_[0]) >> 2];
    const uint32_t b0 = a1 + pp[lp_[1] >> 2], b1 = a1 + pp[(kBitModelTotal - lp_[1]) >> 2];
    #pragma unroll 1
    for (int i = lane; i < table_size; i += 32) {
        uint32_t v;
        if (i < kNumLowLenSymbols) v = a0 + ctx_tree_price(pp, lp_ + len_low(pb, ps), kNumLowLenBits, i);
        else if (i < kNumLowLenSymbols + kNumMidLenSymbols) v = b0 + ctx_tree_price(pp, lp_ + len_mid(pb, ps), kNumMidLenBits, i - kNumLowLenSymbols);
        else v = b1 + ctx_tree_price(pp, lp_ + len_high(pb), kNumHighLenBits, i - kNumLowLenSymbols - kNumMidLenSymbols);
        prices[i] = (uint16_t)v;
    }
    if (lane == 0) c->len_counters[which * 16 + ps] = table_size;
    __syncwarp();
}

__device__ __noinline__ void ctx_fill_distances_prices(const WarpCtx* c, int lane) {
    __syncwarp();
    const uint16_t* pp = c->prob_prices;
    const uint16_t* model = c->model;
    uint16_t* slot_prices = c->slot_prices;
    const int dist_table_size = c->dist_table_size;
    // slot prices first (they do not depend on tempPrices)
    #pragma unroll 1
    for (int k = lane; k < kNumLenToPosStates * dist_table_size; k += 32) {
        const int lps = k / dist_table_size, slot = k - lps * dist_table_size;
        uint32_t v = ctx_tree_price(pp, model + c->off_pos_slot + (lps << kNumPosSlotBits), kNumPosSlotBits, slot);
        if (slot >= kEndPosModelIndex) v += (uint32_t)((((slot >> 1) - 1) - kNumAlignBits) << kNumBitPriceShiftBits);
        slot_prices[(lps << kNumPosSlotBits) + slot] = (uint16_t)v;
    }
    __syncwarp();
    #pragma unroll 1
    for (int k = lane; k < kNumLenToPosStates * kNumFullDistances; k += 32) {
        const int lps = k >> 7, i = k & (kNumFullDistances - 1);
        uint32_t v;
        if (i < kStartPosModelIndex) {
            v = slot_prices[(lps << kNumPosSlotBits) + i];
        } else {
            const int slot = c->fast_pos[i];  // getPosSlot, i < 2^11 (Encoder.java:86-94)
            const int footer = (slot >> 1) - 1;
            const int base = (2 | (slot & 1)) << footer;
            v = (uint32_t)slot_prices[(lps << kNumPosSlotBits) + slot] +
                ctx_reverse_price(pp, model + c->off_pos_dec + base - slot - 1, footer, i - base);
        }
        c->dist_prices[lps * kNumFullDistances + i] = (uint16_t)v;
    }
    __syncwarp();
}

__device__ __noinline__ void ctx_fill_align_prices(const WarpCtx* c, int lane) {
    __syncwarp();
    if (lane < kAlignTableSize)
        c->align_prices[lane] = (uint16_t)ctx_reverse_price(c->prob_prices, c->model + c->off_pos_align, kNumAlignBits, lane);
    __syncwarp();
}

// ICodeProgress.SetProgress(nowPos64, rangeEncoder.getProcessedSizeAdd()) of Encoder.java:922-923, 1070-1072, as deltas
// added to the call's totals in pinned host memory; informational, so no fence.  Out of line and on values only: it runs
// once per kProgressStep input bytes.  Returns the position of the next report.
__device__ __noinline__ uint32_t ctx_report_progress(WarpCtx* c, uint32_t now_pos, int lane) {
    __syncwarp();
    if (lane == 0) {
        const uint32_t produced = c->rc.pos + c->rc.cache_size + 4;  // RangeEncoder.java:69-71
        atomicAdd_system(c->progress, (unsigned long long)(now_pos - c->progress_in));
        atomicAdd_system(c->progress + 1, (unsigned long long)(produced - c->progress_out));
        c->progress_in = now_pos;
        c->progress_out = produced;
    }
    __syncwarp();
    return now_pos + kProgressStep;
}

// ---- everything one stream needs (identical in every lane unless noted) -----
struct Enc {
    const CtaTables* T;
    uint16_t* model;        // fixed part of the probability model (shared)
    uint16_t* lit;          // literal coders (shared, or global when they do not fit)
    uint16_t* dist_prices;  // [4][128]  _distancesPrices
    uint16_t* slot_prices;  // [4][64]   _posSlotPrices
    uint16_t* align_prices; // [16]
    uint16_t* len_prices;   // [2][1<<pb][table_size]
    int32_t* len_counters;  // [2][16]
    uint32_t* md;           // current match list: distances
    uint32_t* md2;          // per pair: length | continuation << 16  (continuation = rep0 length after "match + literal")
    OptNode* ring;          // [R] shared
    OptNode* gopt;          // [kNumOpts] global spill
    OptNode* qbase;         // where Backward left the decision queue (ring or gopt)
    uint32_t rsize, rinv;   // ring size (a multiple of 32) and ceil(2^24 / rsize)
    ModelLayout L;
    const uint8_t* data;
    uint32_t n;
    const uint32_t* idx;    // 1-based
    const uint32_t* pairs;
    int lane;
    int lc, lp, pb, fb, table_size, dist_table_size;
    uint32_t pos_mask, lp_mask;
    bool eos;
    WarpCtx* ctx;           // shared: range-coder state + what the out-of-line helpers need
    uint32_t m;             // match-finder cursor, 0-based (== _pos - 1 of the reference's InWindow)
    const uint16_t* pairs2;
    uint32_t pf_pos, pf_cnt, pf_pair, pf_l2, pf_from;  // prefetched list (pf_pair / pf_l2 differ per lane)
    uint32_t pf_off_pos, pf_off;                        // prefetched idx[] entry
    int state;
    uint32_t prev_byte;
    uint32_t rep_dist[4];
    int num_pairs;
    int additional_offset;
    int opt_end, opt_cur;
    bool longest_found;
    int longest_len;
    int match_price_count, align_price_count;
    uint32_t now_pos;
    uint32_t progress_next;  // input position of the next ICodeProgress report (0xFFFFFFFF: nobody listens)

    // ---- prices (ProbPrices.java:23-37) ----
    __device__ __forceinline__ uint32_t price_bit(uint32_t prob, uint32_t bit) const {
        return T->prob_prices[(((prob - bit) ^ (0u - bit)) & (kBitModelTotal - 1)) >> 2];
    }
    __device__ __forceinline__ uint32_t price0(uint32_t prob) const { return T->prob_prices[prob >> 2]; }
    __device__ __forceinline__ uint32_t price1(uint32_t prob) const { return T->prob_prices[(kBitModelTotal - prob) >> 2]; }
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

    // i mod rsize for i < 4096: the 2^24 reciprocal is exact in that range (rsize >= 128)
    __device__ __forceinline__ uint32_t slot(uint32_t i) const { return i - ((i * rinv) >> 24) * rsize; }
    __device__ __forceinline__ OptNode* node(int i) const { return ring + slot((uint32_t)i); }

    // ---- window (InWindow.java:115-138, whole block resident) ----
    __device__ __forceinline__ uint32_t byte_at(int index) const { return data[m + index]; }
    __device__ __forceinline__ int avail() const { return (int)(n - m); }

    // GetMatchLen for an absolute start `s`, 32 bytes per round (all lanes, uniform result)
    __device__ __forceinline__ int match_len_abs(uint32_t s, uint32_t distance, int limit) const {
        return warp_match_len(data, n, lane, s, distance, limit);
    }

    // Run of set bits in a 32-byte equality bitmap starting at bit `start`, capped at `limit`.
    // Pure bit arithmetic: when the run reaches the end of the bitmap with the limit not exhausted,
    // `more` is set and the caller continues in memory (eq_more) -- rare, and kept out of the hot path
    // so that the DP step stays small in the instruction cache.
    __device__ __forceinline__ int eq_fast(unsigned e, int start, int limit, bool& more) const {
        more = false;
        if (limit <= 0) return 0;
        if (start >= 32) {
            more = true;
            return 0;
        }
        const int room = 32 - start;
        const unsigned w = ~(e >> start);
        const int run = w ? __ffs(w) - 1 : 32;  // for start > 0 the top `start` bits of w are set, so run <= room
        if (run >= room && limit > room) {
            more = true;
            return room;
        }
        return run < limit ? run : limit;
    }
    // continuation of eq_fast: `base` = absolute position of bit 0 of the bitmap
    __device__ __forceinline__ int eq_more(int fast, int start, uint32_t base, uint32_t distance, int limit) const {
        return fast + match_len_abs(base + start + fast, distance, limit - fast);
    }

    // ---- match list ----
    __device__ __forceinline__ int md_len(int i) const { return (int)reinterpret_cast<const uint16_t*>(md2)[2 * i]; }
    __device__ __forceinline__ int md_cont(int i) const { return (int)reinterpret_cast<const uint16_t*>(md2)[2 * i + 1]; }
    __device__ __forceinline__ uint32_t md_dist(int i) const { return md[i]; }

    // List of 0-based position p: idx[p + 1] -> pairs[off] = count, pairs[off + 1 ..] = pairs.
    // The lists are read one step ahead into registers (lane i holds pair i), and idx[] two steps
    // ahead, so that a DP step never waits for global memory on the list.
    __device__ __forceinline__ void prefetch_list(uint32_t p) {  // p = position whose list to fetch
        pf_pos = p;
        pf_cnt = 0;
        pf_pair = 0;
        pf_l2 = 0;
        uint32_t off = kMfEmpty;
        if (p < n) off = (pf_off_pos == p) ? pf_off : idx[p + 1];
        if (off != kMfEmpty) {
            pf_cnt = pairs[off];
            pf_pair = pairs[off + 1 + lane];   // blind: lanes >= count read slack
            pf_l2 = pairs2[off + 1 + lane];
        }
        pf_from = off;
        pf_off_pos = p + 1;
        pf_off = kMfEmpty;
        if (p + 1 < n) pf_off = idx[p + 2];
    }

    __device__ __forceinline__ int read_match_distances() {  // Encoder.java:275-287
        __syncwarp();  // everyone is done with the previous list
        int cnt = 0;
        // one inlined copy of prefetch_list: round 1 fetches the wanted list if the prefetch missed
        // (after a Skip), the last round prefetches the next position
#pragma unroll 1
        for (bool consumed = false;;) {
            if (!consumed && pf_pos == m) {
                cnt = (int)pf_cnt;
                if (lane < cnt) {
                    md[lane] = pair_dist(pf_pair);
                    md2[lane] = pair_len(pf_pair, pf_l2) | (pair_cont(pf_l2) << 16);
                }
#pragma unroll 1
                for (int i = 32 + lane; i < cnt; i += 32) {
                    const uint32_t w = pairs[pf_from + 1 + i], w2 = pairs2[pf_from + 1 + i];
                    md[i] = pair_dist(w);
                    md2[i] = pair_len(w, w2) | (pair_cont(w2) << 16);
                }
                m++;  // fillMatches advanced the window
                consumed = true;
            }
            prefetch_list(m);
            if (consumed) break;
        }
        __syncwarp();
        num_pairs = cnt;
        int length = 0;
        if (cnt > 0) {
            length = md_len(cnt - 1);
            if (length == fb) length += match_len_abs(m + length - 1, md_dist(cnt - 1), kMatchMaxLen - length);
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
    __device__ __forceinline__ void len_update_table(int which, uint32_t ps) { ctx_len_update_table(ctx, which, ps, lane); }
    // all lanes, after the bits were emitted: LenPriceTableEncoder.encode :32-37
    __device__ __forceinline__ void len_count(int which, uint32_t ps) {
        const int c = len_counters[which * 16 + ps] - 1;
        __syncwarp();
        if (c == 0) {
            len_update_table(which, ps);
        } else {
            if (lane == 0) len_counters[which * 16 + ps] = c;
            __syncwarp();
        }
    }

    // ---- literal coder (LiteralEncoder.java:17-64) ----
    // ---- decision lists: lane `j` of a group describes one binary decision of the symbol ----
    // MSB-first bit tree of `nbits` (BitTreeEncoder.encode :18-26): decision j codes bit nbits-1-j at node m_j
    __device__ __forceinline__ void tree_decision(int j, uint16_t* probs, int nbits, uint32_t symbol, uint16_t*& ptr, uint32_t& bit) const {
        ptr = probs + ((1u << j) | (symbol >> (nbits - j)));
        bit = (symbol >> (nbits - 1 - j)) & 1;
    }
    // LSB-first bit tree (ReverseEncode :28-36, Encoder.java:196-205)
    __device__ __forceinline__ void reverse_decision(int j, uint16_t* probs, uint32_t symbol, uint16_t*& ptr, uint32_t& bit) const {
        ptr = probs + ((1u << j) | (j ? __brev(symbol) >> (32 - j) : 0u));
        bit = (symbol >> j) & 1;
    }
    // LenEncoder.encode (:33-48): returns the number of decisions; lane offset j within the symbol
    __device__ __forceinline__ int len_decisions(int j, int which, uint32_t symbol, uint32_t ps, uint16_t*& ptr, uint32_t& bit) const {
        uint16_t* lp_ = model + (which ? L.rep_len : L.len);
        int nchoice, nbits;
        uint16_t* tree;
        uint32_t sym;
        if (symbol < kNumLowLenSymbols) {
            nchoice = 1; nbits = kNumLowLenBits; tree = lp_ + len_low(pb, ps); sym = symbol;
        } else if (symbol < kNumLowLenSymbols + kNumMidLenSymbols) {
            nchoice = 2; nbits = kNumMidLenBits; tree = lp_ + len_mid(pb, ps); sym = symbol - kNumLowLenSymbols;
        } else {
            nchoice = 2; nbits = kNumHighLenBits; tree = lp_ + len_high(pb); sym = symbol - kNumLowLenSymbols - kNumMidLenSymbols;
        }
        if (j < 0) {
            // a lane that belongs to an earlier part of the symbol
        } else if (j < nchoice) {
            ptr = lp_ + j;
            bit = (j == 0) ? (symbol >= kNumLowLenSymbols) : (symbol >= kNumLowLenSymbols + kNumMidLenSymbols);
        } else if (j < nchoice + nbits) {
            tree_decision(j - nchoice, tree, nbits, sym, ptr, bit);
        }
        return nchoice + nbits;
    }
    // literal (LiteralEncoder.Encoder2.encode / encodeMatched :17-40): decision j codes bit 7-j
    __device__ __forceinline__ void literal_decision(int j, uint16_t* probs, bool matched, uint32_t match_byte, uint32_t symbol,
                                                     uint16_t*& ptr, uint32_t& bit) const {
        const int i = 7 - j;
        uint32_t index = (0x100u | symbol) >> (i + 1);
        bit = (symbol >> i) & 1;
        if (matched && (((match_byte ^ symbol) & 0xFF) >> (i + 1)) == 0) index += (1 + ((match_byte >> i) & 1)) << 8;
        ptr = probs + index;
    }

    // Encoder2.GetPrice (:42-64): lanes 0..7 price one bit each; bit i uses the matched
    // context while every higher bit of symbol and match_byte agrees.  Uniform result.
    __device__ __forceinline__ uint32_t lit_price(const uint16_t* probs, bool match_mode, uint32_t match_byte, uint32_t symbol) const {
        uint32_t price = 0;
        if (lane < 8) {
            const int i = 7 - lane;
            const uint32_t ctx = (0x100u | symbol) >> (i + 1);
            const uint32_t bit = (symbol >> i) & 1;
            uint32_t index = ctx;
            if (match_mode && (((match_byte ^ symbol) & 0xFF) >> (i + 1)) == 0) index += (1 + ((match_byte >> i) & 1)) << 8;
            price = price_bit(probs[index], bit);
        }
        return __reduce_add_sync(kFull, price);  // one REDUX instead of a four-shuffle tree
    }

    // ---- rep / match prices (Encoder.java:296-333) ----
    __device__ __forceinline__ uint32_t rep_len1_price(int st, uint32_t ps) const {
        return price0(*p_is_rep_g0(st)) + price0(*p_is_rep0_long(st, ps));
    }
    __device__ __forceinline__ uint32_t pure_rep_price(int rep_index, int st, uint32_t ps) const {
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

    // ---- price table refresh (Encoder.java:1087-1125): out of line, see ctx_fill_* ----
    __device__ __forceinline__ void fill_distances_prices() {
        ctx_fill_distances_prices(ctx, lane);
        match_price_count = 0;
    }
    __device__ __forceinline__ void fill_align_prices() {
        ctx_fill_align_prices(ctx, lane);
        align_price_count = 0;
    }

    // ---- node ring management ----
    // make nodes (len_end, need] usable: spill the nodes their slots still hold, then price = infinity
    __device__ __forceinline__ void extend(int& len_end, int& wb, int need) {
        if (need <= len_end) return;
        const int new_wb = need - (int)rsize + 1;
        if (new_wb > wb) {
            __syncwarp();
            const uint4* src = reinterpret_cast<const uint4*>(ring);
            uint4* dst = reinterpret_cast<uint4*>(gopt);
            #pragma unroll 1
            for (int k = 2 * wb + lane; k < 2 * new_wb; k += 32) dst[k] = src[2 * slot((uint32_t)k >> 1) + (k & 1)];
            wb = new_wb;
            __syncwarp();
        }
        #pragma unroll 1
        for (int t = len_end + 1 + lane; t <= need; t += 32) node(t)->price = kInfinityPrice;
        len_end = need;
    }

    // strict '<' keeps the first candidate on ties (App. A #8); called by the lane that owns node `at`
    __device__ __forceinline__ void relax(int at, uint32_t price, uint32_t pos_prev, uint32_t back, bool p1, bool p2,
                                          uint32_t pos_prev2, uint32_t back2) {
        OptNode* o = node(at);
        if (price < o->price) *reinterpret_cast<uint4*>(o) = make_uint4(price, back, back2, mk_link(pos_prev, pos_prev2, p1, p2));
    }

    // Backward (Encoder.java:335-362).  Runs on lane 0 over `base` (ring when the chunk never
    // wrapped, else the global spill area after the live part of the ring has been flushed).
    __device__ __forceinline__ int backward(int cur, int wb, uint32_t* back_out) {
        __syncwarp();
        OptNode* base = ring;
        if (wb > 0) {
            const uint4* src = reinterpret_cast<const uint4*>(ring);
            uint4* dst = reinterpret_cast<uint4*>(gopt);
            #pragma unroll 1
            for (int k = 2 * wb + lane; k < 2 * (cur + 1); k += 32) dst[k] = src[2 * slot((uint32_t)k >> 1) + (k & 1)];
            base = gopt;
            __syncwarp();
        }
        qbase = base;
        opt_end = cur;
        uint32_t res_back = 0, res_cur = 0;
        if (lane == 0) {
            OptNode* opt = base;
            uint32_t pos_mem = ln_pos_prev(opt[cur].link);
            uint32_t back_mem = opt[cur].back_prev;
            #pragma unroll 1
            do {
                const uint32_t lk = opt[cur].link;
                if (ln_prev1(lk)) {
                    opt[pos_mem].back_prev = kLit;  // MakeAsChar + PosPrev = posMem - 1
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
                opt[pos_prev].link = (opt[pos_prev].link & ~0xFFFu) | (uint32_t)cur;  // only PosPrev changes
                cur = (int)pos_prev;
            } while (cur > 0);
            res_cur = ln_pos_prev(opt[0].link);
            res_back = opt[0].back_prev;
        }
        __syncwarp();
        opt_cur = (int)__shfl_sync(kFull, res_cur, 0);
        *back_out = __shfl_sync(kFull, res_back, 0);
        return opt_cur;
    }

    __device__ __forceinline__ int get_optimum(uint32_t position, uint32_t* back_out);
    __device__ __forceinline__ void emit_match(uint32_t ps, int len, uint32_t pos, int slot);
    template <bool PROGRESS> __device__ __forceinline__ bool encode_one(bool finish);
    template <bool PROGRESS> __device__ __forceinline__ void run();
};

// getOptimum (Encoder.java:364-811).  Returns the length, *back_out = "pos" of PosAndLength
// (kLit literal, 0..3 rep index, else distance + 4).  Uniform across the warp.
__device__ __forceinline__ int Enc::get_optimum(uint32_t position, uint32_t* back_out) {
    if (opt_end != opt_cur) {  // :365-370
        const OptNode* q = qbase + opt_cur;
        const uint32_t lk = q->link;
        const int len_res = (int)ln_pos_prev(lk) - opt_cur;
        *back_out = q->back_prev;
        opt_cur = (int)ln_pos_prev(lk);
        return len_res;
    }
    opt_cur = 0;
    opt_end = 0;

    // One loop for the first position of the chunk (:371-503) and the following ones (:505-810): they share the list
    // read, the window gather, the rep comparisons and the rep lengths, and the parser is bound by instruction fetch, so
    // one copy of those beats two.  `cur` starts from a zero the compiler cannot see, or it would peel the first
    // iteration off again.
    int num_distance_pairs = 0, num_avail = 0;
    uint32_t c = 0;
    uint32_t reps[4];
    uint32_t a_byte = 0, b_byte[4];
    unsigned eq[4];
    const int ri = lane & 3;  // the four reps are handled by lanes 0..3, one rep each: `my_*` differ per lane
    uint32_t my_rep = 0;
    int my_len = 0;
    uint32_t current_byte = 0, match_byte = 0, pos_state = 0, last_byte = 0;
    uint32_t match_price = 0, rep_match_price = 0, normal_match_price = 0;
    int st = state;
    int len_end = 0;
    int wb = 0;  // nodes [0, wb) live in gopt, the rest in the ring
    int cur;
    asm volatile("mov.u32 %0, 0;" : "=r"(cur));
    #pragma unroll 1
    for (;; cur++) {
        const bool first = cur == 0;
        if (!first && cur == len_end) break;
        int new_len;
        if (first && longest_found) {
            new_len = longest_len;
            longest_found = false;
        } else {
            new_len = read_match_distances();
        }
        num_distance_pairs = num_pairs;
        if (first) {
            if (avail() + 1 < 2) {
                *back_out = kLit;
                return 1;
            }
        } else {
            if (new_len >= fb) {
                longest_len = new_len;
                longest_found = true;
                break;
            }
            position++;
        }
        c = m - 1;  // position of the current byte

        // ---- gather: window bytes + rep comparisons; issued before the node logic to overlap latency
        const bool inr = c + lane < n;
        a_byte = 0;
        if (inr) a_byte = data[c + lane];

        uint32_t cur_price = 0;
        if (first) {
            st = state;
#pragma unroll
            for (int i = 0; i < 4; i++) reps[i] = rep_dist[i];
        } else {
            // ---- node cur -> state, reps (:518-590)
            OptNode* oc = node(cur);
            const uint4 ca = *reinterpret_cast<const uint4*>(oc);  // price, back_prev, back_prev2, link
            const uint32_t clink = ca.w;
            uint32_t pos_prev = ln_pos_prev(clink);
            if (ln_prev1(clink)) {
                pos_prev--;
                if (ln_prev2(clink)) {
                    st = ln_state(node((int)ln_pos_prev2(clink))->link);
                    if (ca.z < kNumRepDistances) st = st_longrep(st);
                    else st = st_match(st);
                } else {
                    st = ln_state(node((int)pos_prev)->link);
                }
                st = st_lit(st);
            } else {
                st = ln_state(node((int)pos_prev)->link);
            }
            if (pos_prev == (uint32_t)cur - 1) {
                if (ca.y == 0) st = st_shortrep(st);
                else st = st_lit(st);
                // reps stay those of the previous step only if that step was cur - 1's node; reload to be exact
                const uint4 pb_ = *reinterpret_cast<const uint4*>(node((int)pos_prev)->backs);
                reps[0] = pb_.x; reps[1] = pb_.y; reps[2] = pb_.z; reps[3] = pb_.w;
            } else {
                uint32_t pos;
                if (ln_prev1(clink) && ln_prev2(clink)) {
                    pos_prev = ln_pos_prev2(clink);
                    pos = ca.z;
                    st = st_longrep(st);
                } else {
                    pos = ca.y;
                    if (pos < kNumRepDistances) st = st_longrep(st);
                    else st = st_match(st);
                }
                const uint4 pb_ = *reinterpret_cast<const uint4*>(node((int)pos_prev)->backs);
                const uint32_t b0 = pb_.x, b1 = pb_.y, b2 = pb_.z, b3 = pb_.w;
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
            if (lane == 0) {
                oc->link = (clink & ~(0xFu << 24)) | ((uint32_t)st << 24);
                *reinterpret_cast<uint4*>(oc->backs) = make_uint4(reps[0], reps[1], reps[2], reps[3]);
            }
            cur_price = ca.x;
        }

#pragma unroll
        for (int i = 0; i < 4; i++) {
            b_byte[i] = 1;
            if (inr) b_byte[i] = data[c + lane - reps[i] - 1];
        }

        int num_avail_full = avail() + 1;  // :627-636
        if (kNumOpts - 1 - cur < num_avail_full) num_avail_full = kNumOpts - 1 - cur;
        num_avail = num_avail_full < fb ? num_avail_full : fb;

#pragma unroll
        for (int i = 0; i < 4; i++) eq[i] = __ballot_sync(kFull, inr && a_byte == b_byte[i]);
        current_byte = __shfl_sync(kFull, a_byte, 0);
        match_byte = __shfl_sync(kFull, b_byte[0], 0);
        pos_state = position & pos_mask;

        // ---- rep lengths (:393-399 with limit 273 at the first position, :669-690 with numAvailableBytes after it)
        my_rep = sel4(reps, ri);
        const unsigned my_eq = sel4(eq, ri);
        {
            const int rep_limit = first ? kMatchMaxLen : num_avail;
            bool more;
            my_len = eq_fast(my_eq, 0, rep_limit, more);
#pragma unroll 1
            for (unsigned mm = __ballot_sync(kFull, more && lane < 4); mm; mm &= mm - 1) {
                const int i = __ffs(mm) - 1;
                const int v = eq_more(__shfl_sync(kFull, my_len, i), 0, c, __shfl_sync(kFull, my_rep, i), rep_limit);
                if (lane == i) my_len = v;
            }
        }

        if (first) {
            const int len_main = new_len;
            // first index of the maximum (:396-398)
            const uint32_t rep_max_len = __reduce_max_sync(kFull, lane < 4 ? (uint32_t)my_len : 0u);
            const int rep_max_index = __ffs(__ballot_sync(kFull, lane < 4 && (uint32_t)my_len == rep_max_len)) - 1;
            if ((int)rep_max_len >= fb) {  // :400-404
                const int len_res = (int)rep_max_len;
                *back_out = (uint32_t)rep_max_index;
                move_pos(len_res - 1);
                return len_res;
            }
            if (len_main >= fb) {  // :406-410
                *back_out = md_dist(num_distance_pairs - 1) + kNumRepDistances;
                move_pos(len_main - 1);
                return len_main;
            }
            if (len_main < 2 && current_byte != match_byte && rep_max_len < 2) {  // :415-417
                *back_out = kLit;
                return 1;
            }
            uint32_t price1_ = price0(*p_is_match(st, pos_state)) +
                               lit_price(lit_coder(position, prev_byte), !st_is_char(st), match_byte, current_byte);
            uint32_t back1 = kLit;  // MakeAsChar
            match_price = price1(*p_is_match(st, pos_state));
            rep_match_price = match_price + price1(*p_is_rep(st));
            if (match_byte == current_byte) {  // :430-436
                const uint32_t short_rep_price = rep_match_price + rep_len1_price(st, pos_state);
                if (short_rep_price < price1_) {
                    price1_ = short_rep_price;
                    back1 = 0;  // MakeAsShortRep
                }
            }
            len_end = len_main >= (int)rep_max_len ? len_main : (int)rep_max_len;
            if (len_end < 2) {
                *back_out = back1;
                return 1;
            }
            __syncwarp();
            if (lane == 0) {
                OptNode* o0 = node(0);
                o0->link = (uint32_t)st << 24;
                o0->backs[0] = reps[0];
                o0->backs[1] = reps[1];
                o0->backs[2] = reps[2];
                o0->backs[3] = reps[3];
                OptNode* o1 = node(1);
                o1->price = price1_;
                o1->back_prev = back1;
                o1->link = mk_link(0, 0, false, false);
            }
            #pragma unroll 1
            for (int len = 2 + lane; len <= len_end; len += 32) node(len)->price = kInfinityPrice;  // :451-455
            __syncwarp();

            const int rep0_len = __shfl_sync(kFull, my_len, 0);
            #pragma unroll 1
            for (unsigned rm = __ballot_sync(kFull, lane < 4 && my_len >= 2); rm; rm &= rm - 1) {  // :457-474, one length per lane
                const int i = __ffs(rm) - 1;
                const int rep_len = __shfl_sync(kFull, my_len, i);
                const uint32_t price = rep_match_price + pure_rep_price(i, st, pos_state);
                #pragma unroll 1
                for (int len = 2 + lane; len <= rep_len; len += 32)
                    relax(len, price + len_price(1, len - 2, pos_state), 0, (uint32_t)i, false, false, 0, 0);
                __syncwarp();
            }

            normal_match_price = match_price + price0(*p_is_rep(st));
            const int start = rep0_len >= 2 ? rep0_len + 1 : 2;  // :478-501
            #pragma unroll 1
            for (int len = start + lane; len <= len_main; len += 32) {
                int offs = 0;
                #pragma unroll 1
                while (len > md_len(offs)) offs++;
                const uint32_t distance = md_dist(offs);
                relax(len, normal_match_price + pos_len_price(distance, len, pos_state), 0, distance + kNumRepDistances, false, false,
                      0, 0);
            }
            __syncwarp();
            last_byte = current_byte;  // data[c] of the previous step
            continue;
        }

        // ---- literal and short rep into node cur + 1 (:598-625)
        const uint32_t cur_and1_price = cur_price + price0(*p_is_match(st, pos_state)) +
                                        lit_price(lit_coder(position, last_byte), !st_is_char(st), match_byte, current_byte);
        last_byte = current_byte;
        OptNode* next = node(cur + 1);
        uint32_t n_price = next->price, n_back = next->back_prev, n_link = next->link;
        bool next_is_char = false, n_dirty = false;
        if (cur_and1_price < n_price) {
            n_price = cur_and1_price;
            n_back = kLit;
            n_link = mk_link((uint32_t)cur, 0, false, false);
            next_is_char = true;
            n_dirty = true;
        }
        match_price = cur_price + price1(*p_is_match(st, pos_state));
        rep_match_price = match_price + price1(*p_is_rep(st));
        if (match_byte == current_byte && !(ln_pos_prev(n_link) < (uint32_t)cur && n_back == 0)) {
            const uint32_t short_rep_price = rep_match_price + rep_len1_price(st, pos_state);
            if (short_rep_price <= n_price) {
                n_price = short_rep_price;
                n_back = 0;
                n_link = mk_link((uint32_t)cur, 0, false, false);
                next_is_char = true;
                n_dirty = true;
            }
        }
        if (n_dirty && lane == 0) {
            next->price = n_price;
            next->back_prev = n_back;
            next->link = n_link;
        }
        if (num_avail_full < 2) continue;

        // ---- rep lengths and every continuation length, then one extension of the node range
        // lanes 0..3 own one rep each: my_len = lenTest, my_len2 = lenTest2 of that rep (:669-700)
        int need = len_end;
        int my_len2 = 0;
        {
            int lim2 = 0;
            bool more = false;
            if (my_len >= 2 && my_len < num_avail_full) {
                lim2 = num_avail_full - 1 - my_len < fb ? num_avail_full - 1 - my_len : fb;
                my_len2 = eq_fast(my_eq, my_len + 1, lim2, more);
            }
#pragma unroll 1
            for (unsigned mm = __ballot_sync(kFull, more && lane < 4); mm; mm &= mm - 1) {
                const int i = __ffs(mm) - 1;
                const int v = eq_more(__shfl_sync(kFull, my_len2, i), __shfl_sync(kFull, my_len, i) + 1, c, __shfl_sync(kFull, my_rep, i),
                                      __shfl_sync(kFull, lim2, i));
                if (lane == i) my_len2 = v;
            }
            int my_need = 0;
            if (lane < 4 && my_len >= 2) my_need = cur + my_len + (my_len2 >= 2 ? 1 + my_len2 : 0);
            const int rep_need = (int)__reduce_max_sync(kFull, (unsigned)my_need);
            if (rep_need > need) need = rep_need;
        }
        const int len_test0 = __shfl_sync(kFull, my_len, 0);
        int lit_rep0_len = 0;
        if (!next_is_char && match_byte != current_byte) {  // :637-641
            const int t = num_avail_full - 1 < fb ? num_avail_full - 1 : fb;
            bool more0;
            lit_rep0_len = eq_fast(eq[0], 1, t, more0);
            if (more0) lit_rep0_len = eq_more(lit_rep0_len, 1, c, reps[0], t);
            if (lit_rep0_len >= 2 && cur + 1 + lit_rep0_len > need) need = cur + 1 + lit_rep0_len;
        }
        const int start_len = len_test0 >= 2 ? len_test0 + 1 : 2;  // :667, :691-693
        if (new_len > num_avail) {  // :737-743
            new_len = num_avail;
            #pragma unroll 1
            for (num_distance_pairs = 0; new_len > md_len(num_distance_pairs); num_distance_pairs++) {}
            __syncwarp();
            if (lane == 0) md2[num_distance_pairs] = (uint32_t)new_len | (md2[num_distance_pairs] & 0xFFFF0000u);  // distance stays
            num_distance_pairs++;
            __syncwarp();
        }
        const bool do_matches = new_len >= start_len;
        if (do_matches) {
            if (cur + new_len > need) need = cur + new_len;
            // "match + literal + rep0" (:766-770): the continuation length was precomputed by the
            // match finder up to fb; it is only asked for when the pair is not the truncated one
            #pragma unroll 1
            for (int j = 0; j < num_distance_pairs; j++) {
                const int lj = md_len(j);
                if (lj < start_len || lj >= num_avail_full) continue;
                const int t = num_avail_full - 1 - lj;
                int l2 = md_cont(j);
                if (l2 > t) l2 = t;
                if (l2 >= 2 && cur + lj + 1 + l2 > need) need = cur + lj + 1 + l2;
            }
        }
        extend(len_end, wb, need);
        __syncwarp();

        // ---- literal + rep0 (:642-664)
        if (lit_rep0_len >= 2) {
            const int state2 = st_lit(st);
            const uint32_t ps_next = (position + 1) & pos_mask;
            const uint32_t next_rep_match_price = cur_and1_price + price1(*p_is_match(state2, ps_next)) + price1(*p_is_rep(state2));
            if (lane == 0)
                relax(cur + 1 + lit_rep0_len, next_rep_match_price + rep_price(0, lit_rep0_len, state2, ps_next), (uint32_t)cur + 1, 0,
                      true, false, 0, 0);
            __syncwarp();
        }

        // ---- reps (:669-735); most positions have no rep of length >= 2 at all
#pragma unroll 1
        for (unsigned rm = __ballot_sync(kFull, lane < 4 && my_len >= 2); rm; rm &= rm - 1) {
            const int rep_index = __ffs(rm) - 1;
            const int lt = __shfl_sync(kFull, my_len, rep_index);
            const uint32_t rp = rep_match_price + pure_rep_price(rep_index, st, pos_state);
            #pragma unroll 1
            for (int len = 2 + lane; len <= lt; len += 32)
                relax(cur + len, rp + len_price(1, len - 2, pos_state), (uint32_t)cur, (uint32_t)rep_index, false, false, 0, 0);
            __syncwarp();
            const int lt2 = __shfl_sync(kFull, my_len2, rep_index);
            if (lt2 >= 2) {  // rep + literal + rep0 (:696-734)
                int state2 = st_longrep(st);
                uint32_t ps_next = (position + lt) & pos_mask;
                uint32_t sym, prv, mb;
                if (lt < 32) {
                    sym = __shfl_sync(kFull, a_byte, lt);
                    prv = __shfl_sync(kFull, a_byte, lt - 1);
                    mb = __shfl_sync(kFull, sel4(b_byte, rep_index), lt);
                } else {
                    sym = data[c + lt];
                    prv = data[c + lt - 1];
                    mb = data[c + lt - __shfl_sync(kFull, my_rep, rep_index) - 1];
                }
                const uint32_t cur_and_len_char_price = rp + len_price(1, lt - 2, pos_state) + price0(*p_is_match(state2, ps_next)) +
                                                        lit_price(lit_coder(position + lt, prv), true, mb, sym);
                state2 = st_lit(state2);
                ps_next = (position + lt + 1) & pos_mask;
                const uint32_t next_match_price = cur_and_len_char_price + price1(*p_is_match(state2, ps_next));
                const uint32_t next_rep_match_price = next_match_price + price1(*p_is_rep(state2));
                if (lane == 0)
                    relax(cur + lt + 1 + lt2, next_rep_match_price + rep_price(0, lt2, state2, ps_next), (uint32_t)(cur + lt + 1), 0,
                          true, true, (uint32_t)cur, (uint32_t)rep_index);
                __syncwarp();
            }
        }

        // ---- matches (:744-809).  A node X = cur + L can be reached by the plain match of length L
        // and by "match l + literal + rep0" with l < L; the reference applies the latter first
        // (they come up at the pair boundary l), so all continuations go first, in pair order.
        if (do_matches) {
            normal_match_price = match_price + price0(*p_is_rep(st));
            #pragma unroll 1
            for (int j = 0; j < num_distance_pairs; j++) {
                const int lj = md_len(j);
                if (lj < start_len || lj >= num_avail_full) continue;
                const int t = num_avail_full - 1 - lj;
                int l2 = md_cont(j);
                if (l2 > t) l2 = t;
                if (l2 < 2) continue;
                const uint32_t cur_back = md_dist(j);
                const uint32_t cur_and_len_price = normal_match_price + pos_len_price(cur_back, lj, pos_state);
                int state2 = st_match(st);
                uint32_t ps_next = (position + lj) & pos_mask;
                const uint32_t sym = data[c + lj], prv = data[c + lj - 1], mb = data[c + lj - cur_back - 1];
                const uint32_t cur_and_len_char_price = cur_and_len_price + price0(*p_is_match(state2, ps_next)) +
                                                        lit_price(lit_coder(position + lj, prv), true, mb, sym);
                state2 = st_lit(state2);
                ps_next = (position + lj + 1) & pos_mask;
                const uint32_t next_match_price = cur_and_len_char_price + price1(*p_is_match(state2, ps_next));
                const uint32_t next_rep_match_price = next_match_price + price1(*p_is_rep(state2));
                if (lane == 0)
                    relax(cur + lj + 1 + l2, next_rep_match_price + rep_price(0, l2, state2, ps_next), (uint32_t)(cur + lj + 1), 0, true,
                          true, (uint32_t)cur, cur_back + kNumRepDistances);
                __syncwarp();
            }
            #pragma unroll 1
            for (int len = start_len + lane; len <= new_len; len += 32) {
                int offs = 0;
                #pragma unroll 1
                while (len > md_len(offs)) offs++;
                const uint32_t cur_back = md_dist(offs);
                relax(cur + len, normal_match_price + pos_len_price(cur_back, len, pos_state), (uint32_t)cur, cur_back + kNumRepDistances,
                      false, false, 0, 0);
            }
            __syncwarp();
        }
    }
    return backward(cur, wb, back_out);
}

// one match-type symbol: isMatch=1, isRep=0, length, posSlot, footer (encodeAMatch :976-1005 and
// WriteEndMarker :818-835 which is the same symbol with len 2 and an all-ones 32-bit "distance")
__device__ __forceinline__ void Enc::emit_match(uint32_t ps, int len, uint32_t pos, int slot) {
    uint16_t* ptr = model;
    uint32_t bit = 0;
    // batch 1: isMatch, isRep, length coder, posSlot tree
    int cnt = 2;
    if (lane == 0) { ptr = p_is_match(state, ps); bit = 1; }
    if (lane == 1) { ptr = p_is_rep(state); bit = 0; }
    const int nlen = len_decisions(lane - 2, 0, (uint32_t)(len - kMatchMinLen), ps, ptr, bit);
    cnt += nlen;
    if (lane >= cnt && lane < cnt + kNumPosSlotBits)
        tree_decision(lane - cnt, model + L.pos_slot + (len_to_pos_state(len) << kNumPosSlotBits), kNumPosSlotBits, (uint32_t)slot, ptr, bit);
    cnt += kNumPosSlotBits;
    rc_batch(&ctx->rc, cnt, ptr, bit, lane);
    // batch 2: footer
    if (slot >= kStartPosModelIndex) {
        const int footer_bits = (slot >> 1) - 1;
        const uint32_t base = (2u | (slot & 1)) << footer_bits;
        const uint32_t pos_reduced = pos - base;
        if (slot < kEndPosModelIndex) {
            if (lane < footer_bits) reverse_decision(lane, model + L.pos_dec + base - slot - 1, pos_reduced, ptr, bit);
            rc_batch(&ctx->rc, footer_bits, ptr, bit, lane);
        } else {
            rc_direct(&ctx->rc, pos_reduced >> kNumAlignBits, footer_bits - kNumAlignBits, lane);  // :998-999
            if (lane < kNumAlignBits) reverse_decision(lane, model + L.pos_align, pos_reduced & kAlignMask, ptr, bit);
            rc_batch(&ctx->rc, kNumAlignBits, ptr, bit, lane);
        }
    }
}

// encodeOne (:890-936) with its emitters (:938-1024); false once the input is used up.  With
// `finish` it emits the end marker instead (WriteEndMarker :818-835: the match symbol with len 2,
// posSlot 63 and an all-ones 32-bit "distance"), through the same emission code as every match.
template <bool PROGRESS>
__device__ __forceinline__ bool Enc::encode_one(bool finish) {
    uint32_t back = kNumRepDistances;
    int len = kMatchMinLen;
    if (!finish) len = get_optimum(now_pos, &back);
    const uint32_t ps = now_pos & pos_mask;
    __syncwarp();
    uint16_t* ptr = model;
    uint32_t bit = 0;
    if (len == 1 && back == kLit) {  // encodeSingleByteLiteral :1007-1024
        const uint32_t cur_byte = byte_at(0 - additional_offset);
        const bool matched = !st_is_char(state);
        uint32_t mb = 0;
        if (matched) mb = byte_at(0 - (int)rep_dist[0] - 1 - additional_offset);
        if (lane == 0) { ptr = p_is_match(state, ps); bit = 0; }
        else if (lane <= 8) literal_decision(lane - 1, lit_coder(now_pos, prev_byte), matched, mb, cur_byte, ptr, bit);
        rc_batch(&ctx->rc, 9, ptr, bit, lane);
        prev_byte = cur_byte;
        state = st_lit(state);
    } else {
        if (back < kNumRepDistances) {  // encodeARepetition :938-974
            int cnt;
            if (lane == 0) { ptr = p_is_match(state, ps); bit = 1; }
            if (lane == 1) { ptr = p_is_rep(state); bit = 1; }
            if (lane == 2) { ptr = p_is_rep_g0(state); bit = back != 0; }
            if (back == 0) {
                if (lane == 3) { ptr = p_is_rep0_long(state, ps); bit = len != 1; }
                cnt = 4;
            } else {
                if (lane == 3) { ptr = p_is_rep_g1(state); bit = back != 1; }
                cnt = 4;
                if (back != 1) {
                    if (lane == 4) { ptr = p_is_rep_g2(state); bit = back - 2; }
                    cnt = 5;
                }
            }
            if (len != 1) cnt += len_decisions(lane - cnt, 1, (uint32_t)(len - kMatchMinLen), ps, ptr, bit);
            rc_batch(&ctx->rc, cnt, ptr, bit, lane);
            if (len == 1) {
                state = st_shortrep(state);
            } else {
                len_count(1, ps);
                state = st_longrep(state);
            }
            const uint32_t distance = sel4(rep_dist, (int)back);
            if (back != 0) {
                if (back == 3) rep_dist[3] = rep_dist[2];
                if (back >= 2) rep_dist[2] = rep_dist[1];
                rep_dist[1] = rep_dist[0];
                rep_dist[0] = distance;
            }
        } else {  // encodeAMatch :976-1005
            uint32_t pos = back - kNumRepDistances;
            int slot = (1 << kNumPosSlotBits) - 1;
            if (finish) pos = 0xFFFFFFFFu;  // posReduced = 2^30 - 1 (26 direct one-bits, align 15): pos - base with base = 3 << 30
            else slot = pos_slot(pos);
            emit_match(ps, len, pos, slot);
            state = st_match(state);
            if (finish) return false;
            len_count(0, ps);
            if (slot >= kEndPosModelIndex) align_price_count++;
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
    if (PROGRESS && now_pos >= progress_next) progress_next = ctx_report_progress(ctx, now_pos, lane);
    if (additional_offset == 0) {
        if (match_price_count >= (1 << 7)) fill_distances_prices();
        if (align_price_count >= kAlignTableSize) fill_align_prices();
        if (avail() == 0) return false;
    }
    return true;
}

// SetStreams + CodeOneBlock loop (Encoder.java:1046-1077, 843-888); probabilities already initialised
template <bool PROGRESS>
__device__ __forceinline__ void Enc::run() {
    state = 0;
    prev_byte = 0;
    rep_dist[0] = rep_dist[1] = rep_dist[2] = rep_dist[3] = 0;
    longest_found = false;
    longest_len = 0;
    opt_end = opt_cur = 0;
    additional_offset = 0;
    m = 0;
    pf_pos = 0xFFFFFFFFu;
    pf_off_pos = 0xFFFFFFFFu;
    pf_off = kMfEmpty;
    pf_cnt = pf_pair = pf_l2 = pf_from = 0;
    now_pos = 0;
    progress_next = ctx->progress ? kProgressStep : 0xFFFFFFFFu;
    num_pairs = 0;
    match_price_count = 0;
    align_price_count = 0;
    qbase = ring;
    fill_distances_prices();
    fill_align_prices();
    #pragma unroll 1
    for (int which = 0; which < 2; which++)
        #pragma unroll 1
        for (uint32_t ps = 0; ps < (1u << pb); ps++) len_update_table(which, ps);

    bool more = avail() != 0;
    if (more) {
        // The first byte is always a plain literal (:860-878).  ReadMatchDistances at position 0 finds
        // nothing (the match finder is empty), so it only advances the window.
        m = 1;
        additional_offset = 1;
        const uint32_t cur_byte = byte_at(0 - additional_offset);
        uint16_t* ptr = model;
        uint32_t bit = 0;
        if (lane == 0) { ptr = p_is_match(state, 0); bit = 0; }
        else if (lane <= 8) literal_decision(lane - 1, lit_coder(0, prev_byte), false, 0, cur_byte, ptr, bit);
        rc_batch(&ctx->rc, 9, ptr, bit, lane);
        state = st_lit(state);
        prev_byte = cur_byte;
        additional_offset--;
        now_pos++;
        more = avail() != 0;
    }
    #pragma unroll 1
    while (more) more = encode_one<PROGRESS>(false);
    if (eos) encode_one<PROGRESS>(true);  // Flush (:837-841): the end marker if asked for ...
    rc_flush(&ctx->rc, lane);   // ... and five shiftLow
}

// MAXW = launch bound in warps: the register budget is 64 K / (32 MAXW).  Two instances: up to
// kEncWarpsLitSmem streams per SM with the literal coders in shared memory (168 registers), and up
// to kEncMaxWarps with the literal coders in global memory (L2), which is what lets a wave with more
// blocks than the first variant's slots keep 12-14 serial chains per SM in flight.
// FIXED: lc3 lp0 pb2, the properties of nearly every .lzma stream (and of the reference's defaults, Encoder.java:151-153),
// as compile-time constants: the probability-model offsets, pos_state masks and literal-coder strides fold into the
// addressing, which is worth ~200 instructions of a hot path that is bound by instruction fetch.
template <int MAXW, bool PROGRESS, bool FIXED>
__global__ void __launch_bounds__(MAXW * 32, 1) lzb_parse_kernel(ParseArgs a) {
    const int a_lc = FIXED ? 3 : a.lc, a_lp = FIXED ? 0 : a.lp, a_pb = FIXED ? 2 : a.pb;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    CtaTables* tables = reinterpret_cast<CtaTables*>(smem_raw);
    init_cta_tables(tables);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warps = blockDim.x >> 5;
    uint8_t* slice = smem_raw + sizeof(CtaTables) + (size_t)warp * a.slice_bytes;
    const size_t slot = (size_t)blockIdx.x * warps + warp;
    const ModelLayout L = make_layout(a_lc, a_lp, a_pb);
    const SliceLayout S = make_slice(a_lc, a_lp, a_pb, a.fb, a.lit_in_smem);
    uint16_t* model = reinterpret_cast<uint16_t*>(slice);
    uint16_t* lit = S.lit_in_smem ? model + L.literal : a.lit_scratch + slot * (size_t)L.n_literal;
    WarpCtx* ctx = reinterpret_cast<WarpCtx*>(slice + S.ctx);
    if (lane == 0) {
        ctx->prob_prices = tables->prob_prices;
        ctx->fast_pos = tables->fast_pos;
        ctx->model = model;
        ctx->dist_prices = reinterpret_cast<uint16_t*>(slice + S.dist_prices);
        ctx->slot_prices = reinterpret_cast<uint16_t*>(slice + S.slot_prices);
        ctx->align_prices = reinterpret_cast<uint16_t*>(slice + S.align_prices);
        ctx->len_prices = reinterpret_cast<uint16_t*>(slice + S.len_prices);
        ctx->len_counters = reinterpret_cast<int32_t*>(slice + S.len_counters);
        ctx->pb = a_pb;
        ctx->table_size = a.fb + 1 - kMatchMinLen;
        ctx->dist_table_size = a.dist_table_size;
        ctx->off_len = L.len;
        ctx->off_rep_len = L.rep_len;
        ctx->off_pos_slot = L.pos_slot;
        ctx->off_pos_dec = L.pos_dec;
        ctx->off_pos_align = L.pos_align;
    }
    __syncwarp();

    // The first block of a warp is fixed by its place (the warps of a CTA take consecutive entries of the
    // cost-sorted order: an SM's streams then belong to the same kind of data and run the same parts of this
    // kernel, which is what its instruction cache can hold); later blocks are drawn from the ticket counter.
    const uint32_t slots = gridDim.x * (uint32_t)warps;
    uint32_t b = (uint32_t)slot;
    #pragma unroll 1
    for (;; ) {
        if (b >= a.n_blocks) break;
        const uint32_t ticket_b = b;
        if (a.order) b = a.order[b];
        const uint32_t n = (uint32_t)a.in_len[b];
        uint8_t* out = a.out + a.out_off[b];
        uint64_t cap = a.out_cap[b];
        uint64_t header = 0;
        if (a.with_header) {  // LzmaAlone.java:208-217
            if (cap >= LZB_KERNEL_HEADER) {
                if (lane < LZB_KERNEL_HEADER) {
                    uint32_t v;
                    if (lane == 0) v = (uint32_t)((a_pb * 5 + a_lp) * 9 + a_lc);
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
        __syncwarp();
        #pragma unroll 1
        for (int i = lane; i < L.n_fixed; i += 32) model[i] = kProbInit;
        #pragma unroll 1
        for (int i = lane; i < L.n_literal; i += 32) lit[i] = kProbInit;
        {
            uint32_t* z = reinterpret_cast<uint32_t*>(slice + S.dist_prices);
            const uint32_t words = (S.ring - S.dist_prices) / 4;
            #pragma unroll 1
            for (uint32_t i = lane; i < words; i += 32) z[i] = 0;
        }
        __syncwarp();
        rc_init(&ctx->rc, out, cap, lane);
        if (lane == 0) {
            ctx->progress = a.progress;
            ctx->progress_in = ctx->progress_out = 0;
        }
        __syncwarp();
        Enc e;
        e.T = tables;
        e.model = model;
        e.lit = lit;
        e.dist_prices = ctx->dist_prices;
        e.slot_prices = ctx->slot_prices;
        e.align_prices = ctx->align_prices;
        e.len_prices = ctx->len_prices;
        e.len_counters = ctx->len_counters;
        e.md = reinterpret_cast<uint32_t*>(slice + S.md);
        e.md2 = reinterpret_cast<uint32_t*>(slice + S.md2);
        e.ring = reinterpret_cast<OptNode*>(slice + S.ring);
        e.rsize = S.ring_nodes;
        e.rinv = ((1u << 24) + S.ring_nodes - 1) / S.ring_nodes;
        e.gopt = reinterpret_cast<OptNode*>(a.opt_scratch) + slot * (size_t)kNumOpts;
        e.ctx = ctx;
        e.L = L;
        e.data = a.in + a.in_off[b];
        e.n = n;
        const BlockLists bl = a.lists[b];
        e.idx = reinterpret_cast<const uint32_t*>(a.pool + bl.idx_off);
        e.pairs = reinterpret_cast<const uint32_t*>(a.pool + bl.pairs_off);
        e.pairs2 = reinterpret_cast<const uint16_t*>(a.pool + bl.pairs2_off);
        e.lane = lane;
        e.lc = a_lc;
        e.lp = a_lp;
        e.pb = a_pb;
        e.fb = a.fb;
        e.table_size = a.fb + 1 - kMatchMinLen;
        e.dist_table_size = a.dist_table_size;
        e.pos_mask = (1u << a_pb) - 1;
        e.lp_mask = (1u << a_lp) - 1;
        e.eos = a.eos;
        e.run<PROGRESS>();
        if (lane == 0) a.out_len[b] = (uint64_t)ctx->rc.pos > cap ? ~0ull : (uint64_t)ctx->rc.pos + header;
        __syncwarp();
        (void)ticket_b;
        b = 0;
        if (lane == 0) b = slots + atomicAdd(a.ticket, 1u);
        b = __shfl_sync(kFull, b, 0);
    }
}

// ---- host side ---------------------------------------------------------------
// Streams per SM for a wave of `blocks_per_sm` blocks per SM.  With the literal coders in shared
// memory a stream's slice is 28.5 KB at lc3 lp0 fb 64 (8 per SM); with them in global memory 16.5 KB
// (13 per SM, the second kernel instance).  A lone stream is 2-3 % faster with shared literals, so the
// global layout is used only when the wave has more blocks than the shared layout could hold at once
// (or when the literal coders do not fit shared memory at all: lc + lp > 4).
ParseGeometry parse_geometry(int lc, int lp, int pb, int fb, uint32_t blocks_per_sm, int force_lit) {
    const uint32_t usable = 232448 - (uint32_t)sizeof(CtaTables);
    auto geo = [&](bool with_lit) {
        ParseGeometry g;
        const SliceLayout s = make_slice(lc, lp, pb, fb, with_lit);
        g.slice_bytes = (s.total + 127) & ~127u;
        int warps = (int)(usable / g.slice_bytes);
        const int cap = with_lit ? kEncWarpsLitSmem : kEncMaxWarps;
        if (warps > cap) warps = cap;
        g.max_warps = warps;  // 0: does not fit
        g.lit_in_smem = with_lit;
        g.cta_table_bytes = (uint32_t)sizeof(CtaTables);
        return g;
    };
    const ParseGeometry gs = geo(true), gg = geo(false);
    bool use_smem = gs.max_warps >= 1 && (gs.max_warps >= gg.max_warps || blocks_per_sm <= (uint32_t)gs.max_warps);
    if (force_lit == 0 && gs.max_warps >= 1) use_smem = true;   // LZB_ENC_LIT=smem  (test / tuning hook)
    if (force_lit == 1) use_smem = false;                       // LZB_ENC_LIT=global
    ParseGeometry g = use_smem ? gs : gg;
    if (g.max_warps < 1) g.max_warps = 1;
    return g;
}

size_t parse_opt_bytes_per_slot() { return sizeof(OptNode) * (size_t)kNumOpts; }

cudaError_t launch_parse(const ParseArgs& a, int grid, int warps, cudaStream_t st) {
    const size_t smem = sizeof(CtaTables) + (size_t)warps * a.slice_bytes;
    // eight instances: where the literal coders live x whether anybody listens to ICodeProgress (the test for the next
    // report sits in the symbol loop: two instructions and a register that cost 3 % on C3 when compiled in) x FIXED
    const bool fixed = a.lc == 3 && a.lp == 0 && a.pb == 2;
    decltype(&lzb_parse_kernel<kEncMaxWarps, false, false>) kern;
    if (a.lit_in_smem) {
        kern = a.progress ? (fixed ? lzb_parse_kernel<kEncWarpsLitSmem, true, true> : lzb_parse_kernel<kEncWarpsLitSmem, true, false>)
                          : (fixed ? lzb_parse_kernel<kEncWarpsLitSmem, false, true> : lzb_parse_kernel<kEncWarpsLitSmem, false, false>);
    } else {
        kern = a.progress ? (fixed ? lzb_parse_kernel<kEncMaxWarps, true, true> : lzb_parse_kernel<kEncMaxWarps, true, false>)
                          : (fixed ? lzb_parse_kernel<kEncMaxWarps, false, true> : lzb_parse_kernel<kEncMaxWarps, false, false>);
    }
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return e;
    kern<<<grid, warps * 32, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace lzb
