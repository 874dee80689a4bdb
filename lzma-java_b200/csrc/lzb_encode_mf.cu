// lzb_encode_mf.cu -- the bt4 / bt2 binary-tree match finder as a parallel
// pre-pass (LZ/BinTree.java:152-356 of rfalke/lzma-java).
//
// fillMatches0 and Skip update the hash heads and the tree identically, and
// every position goes through exactly one of them in order, so the list of
// (length, distance) pairs of a position is a pure function of the data
// (SURVEY.md section 3.1).  Two kernels reproduce it bit for bit:
//
//   link:  the three head tables _hash[h2], _hash[1024+h3], _hash[66560+h4]
//          are "latest earlier position with the same key".  One warp per
//          block replays them 32 positions per step: __match_any_sync finds
//          equal keys inside the step (the nearest lower lane is the
//          predecessor), the table supplies the predecessor from earlier
//          steps, and the highest lane of each key group writes the table.
//          Output: per position its hash-2 / hash-3 candidates and a forward
//          link to the next position of the same hash-4 bucket.
//   tree:  distinct hash-4 buckets own disjoint trees (App. C): the descent
//          for position p starts at the bucket's previous position and only
//          touches nodes linked by that bucket; p's own two slots are its
//          own.  One thread per bucket walks the forward links and performs
//          the reference's insertion for each position, with _son indexed by
//          absolute position (no cyclic reuse, so buckets never alias;
//          candidates at or below matchMinPos are cut exactly as in
//          BinTree.java:164,231).
#include "lzb_encode.cuh"

namespace lzb {

__constant__ uint32_t c_crc[256];  // CRC.java:6-20, the hash mixer of BinTree.java:171-175

static uint32_t h_crc[256];
static bool h_crc_ready = false;

cudaError_t upload_mf_tables() {
    if (!h_crc_ready) {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t r = i;
            for (int j = 0; j < 8; j++) r = (r & 1) ? (r >> 1) ^ 0xEDB88320u : r >> 1;
            h_crc[i] = r;
        }
        h_crc_ready = true;
    }
    return cudaMemcpyToSymbol(c_crc, h_crc, sizeof h_crc);
}

constexpr unsigned kFull = 0xFFFFFFFFu;

// One table replay for the lanes of this step.  Predecessor of `key` = nearest lower lane with
// the same key in this step, else the table entry; the highest lane of each key group is the one
// that will overwrite the table.  The table reads of all three hashes are issued first, the
// writes follow after one __syncwarp, so a step costs one memory round trip, not three.
struct LinkStep {
    unsigned peers, lower;
    uint32_t from_table;
    __device__ __forceinline__ void read(const uint32_t* table, uint32_t key, unsigned vm, int lane) {
        peers = __match_any_sync(vm, key);
        lower = peers & ((1u << lane) - 1u);
        from_table = 0;
        if (!lower) from_table = table[key];
    }
    __device__ __forceinline__ uint32_t prev(uint32_t base) const {
        return lower ? base + 32u - (uint32_t)__clz(lower) : from_table;
    }
    __device__ __forceinline__ void write(uint32_t* table, uint32_t key, int lane, uint32_t pos1) const {
        if ((peers >> lane) == 1u) table[key] = pos1;
    }
};

// one thread per block: a block longer than the declared max_in_len would overflow its scratch slices
__global__ void __launch_bounds__(256) lzb_mf_check_lengths(MfWave w) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < w.n_blocks && w.in_len[b] > (uint64_t)w.np - 1) atomicMax(w.overflow, 2u);
}

__global__ void __launch_bounds__(128) lzb_mf_link_kernel(MfWave w) {
    __shared__ uint32_t s_crc[256];  // constant memory would serialise the 32 different indices of a warp
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_crc[i] = c_crc[i];
    __syncthreads();
    const uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= w.n_blocks) return;
    const uint8_t* data = w.in + w.in_off[b];
    // A block longer than the caller's max_in_len is flagged by lzb_mf_check_lengths and the host gives up on the
    // batch; here it is merely cut to what its scratch slices hold.  (The check used to sit in this kernel as an
    // early return: 2048 x 1 MiB blocks then took 0.72 s instead of 0.24 s, measured by removing just that.)
    const uint64_t len64 = w.in_len[b];
    const uint32_t n = len64 > (uint64_t)w.np - 1 ? w.np - 1 : (uint32_t)len64;
    uint32_t* heads = w.heads + (size_t)b * w.hash_stride;
    uint32_t* next = w.next + (size_t)b * w.np;
    uint32_t* prev2 = w.prev2 + (size_t)b * w.np;
    uint32_t* prev3 = w.prev3 + (size_t)b * w.np;
    uint32_t* idx = w.idx + (size_t)b * w.np;
    uint16_t* cnt_out = w.cnt + (size_t)b * w.np;
    uint32_t* head3 = heads + kHash2Size;
    uint32_t* head4 = w.bt4 ? heads + kHash2Size + kHash3Size : heads;  // kFixHashSize, BinTree.java:57-69
    const uint32_t min_check = w.bt4 ? 4 : 3;                           // kMinMatchCheck

    // the four bytes at a position, packed; requested one step ahead so that a step waits for one memory round trip (the
    // table gather) instead of two (window bytes, then the gather)
    auto load4 = [&](uint32_t q) -> uint32_t {
        uint32_t v = 0;
        if (q + 3 < n) v = (uint32_t)data[q] | ((uint32_t)data[q + 1] << 8) | ((uint32_t)data[q + 2] << 16) | ((uint32_t)data[q + 3] << 24);
        else if (q + 1 < n) v = (uint32_t)data[q] | ((uint32_t)data[q + 1] << 8);  // bt2 hashes two bytes (positions with >= 3 left)
        return v;
    };
    uint32_t cur4 = load4((uint32_t)lane);
    for (uint32_t base = 0; base < n; base += 32) {
        const uint32_t p = base + lane, pos1 = p + 1;
        const uint32_t nxt4 = load4(p + 32);
        const bool in_range = p < n;
        // positions with lenLimit < kMinMatchCheck are not inserted at all (BinTree.java:158-161)
        const bool valid = in_range && n - p >= min_check;
        const unsigned vm = __ballot_sync(kFull, valid);
        uint32_t h2 = 0, h3 = 0, h4 = 0;
        LinkStep s2, s3, s4;
        if (valid) {
            const uint32_t d0 = cur4 & 0xFF, d1 = (cur4 >> 8) & 0xFF, d2 = (cur4 >> 16) & 0xFF, d3 = cur4 >> 24;
            if (w.bt4) {  // BinTree.java:171-175
                uint32_t t = s_crc[d0] ^ d1;
                h2 = t & (kHash2Size - 1);
                t ^= d2 << 8;
                h3 = t & (kHash3Size - 1);
                h4 = (t ^ (s_crc[d3] << 5)) & w.hash_mask;
                s2.read(heads, h2, vm, lane);
                s3.read(head3, h3, vm, lane);
            } else {
                h4 = d0 ^ (d1 << 8);
            }
            s4.read(head4, h4, vm, lane);
        }
        __syncwarp();  // every lane has read the tables before the group leaders overwrite them
        if (valid) {
            uint32_t c2 = 0, c3 = 0;
            if (w.bt4) {
                c2 = s2.prev(base);
                c3 = s3.prev(base);
                s2.write(heads, h2, lane, pos1);
                s3.write(head3, h3, lane, pos1);
            }
            const uint32_t prev4 = s4.prev(base);
            s4.write(head4, h4, lane, pos1);
            if (prev4) next[prev4] = pos1;
            prev2[pos1] = c2 | (prev4 ? 0u : kHeadFlag);
            prev3[pos1] = c3;
        } else if (in_range) {
            idx[pos1] = kMfEmpty;
            cnt_out[pos1] = 0;
            prev2[pos1] = 0;
        }
        __syncwarp();
        cur4 = nxt4;
    }
}

// ---- byte-run comparison, eight bytes per step ---------------------------------
// The descent compares the window at the new position with the window at a candidate, starting
// at the length both are already known to share (BinTree.java:244-248: a byte loop).  A byte loop
// costs one dependent load round trip per byte; this reads both windows as unaligned 64-bit words
// (two aligned loads + a funnel shift) and finds the first differing byte with one ffs.
// Only aligned words that contain a byte the byte loop could have read are touched.
__device__ __forceinline__ uint64_t load_u64_at(const uint8_t* p) {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(reinterpret_cast<uintptr_t>(p) & ~uintptr_t(7));
    const unsigned s = (unsigned)(reinterpret_cast<uintptr_t>(p) & 7u) * 8u;
    const uint64_t lo = w[0];
    if (s == 0) return lo;
    return (lo >> s) | (w[1] << (64u - s));
}
// first i in [len, limit] with a[i] != c[i] (or limit); `room` = bytes that exist from a[0] on (room >= limit), c < a
__device__ __forceinline__ uint32_t extend_run(const uint8_t* a, const uint8_t* c, uint32_t len, uint32_t limit, uint32_t room) {
    while (len < limit && len + 8 <= room) {
        const uint64_t x = load_u64_at(a + len) ^ load_u64_at(c + len);
        if (x) {
            len += (uint32_t)(__ffsll((long long)x) - 1) >> 3;
            return len < limit ? len : limit;
        }
        len += 8;
    }
    if (len > limit) len = limit;
    while (len < limit && a[len] == c[len]) len++;
    return len;
}

// ---- pieces shared by the two tree kernels ------------------------------------
struct TreeBlock {
    const uint8_t* buf;  // buf[pos1] is the byte at 1-based position pos1
    const uint32_t* prev2;
    const uint32_t* prev3;
    const uint32_t* next;
    uint32_t* son;
    uint32_t* idx;
    uint16_t* cnt;
    uint32_t* pairs_out;
    uint16_t* pairs2_out;
    uint32_t n;
    __device__ __forceinline__ TreeBlock(const MfWave& w, uint32_t b) {
        buf = w.in + w.in_off[b] - 1;
        prev2 = w.prev2 + (size_t)b * w.np;
        prev3 = w.prev3 + (size_t)b * w.np;
        next = w.next + (size_t)b * w.np;
        son = w.son + (size_t)b * 2 * w.np;
        idx = w.idx + (size_t)b * w.np;
        cnt = w.cnt + (size_t)b * w.np;
        pairs_out = w.pairs + (size_t)b * w.pair_cap;
        pairs2_out = w.pairs2 + (size_t)b * w.pair_cap;
        n = (uint32_t)w.in_len[b];
    }
};

// hash-2 / hash-3 candidates (BinTree.java:183-208) or the bt2 direct-byte check (:218-226)
// A thread's list under construction: distances and lengths side by side (local memory).
struct PairBuf {
    uint32_t dist[kMatchMaxLen];
    uint16_t len[kMatchMaxLen];
    __device__ __forceinline__ void put(uint32_t i, uint32_t l, uint32_t d) {
        dist[i] = d;
        len[i] = (uint16_t)l;
    }
};

__device__ __forceinline__ void tree_prepairs(const MfWave& w, const TreeBlock& t, uint32_t pos1, uint32_t cur_match,
                                              uint32_t match_min_pos, PairBuf& pairs, uint32_t& cnt, uint32_t& max_len) {
    const uint8_t* cur = t.buf + pos1;
    max_len = 1;  // kStartMaxLen
    cnt = 0;
    if (w.bt4) {
        uint32_t c2 = t.prev2[pos1] & ~kHeadFlag;
        const uint32_t c3 = t.prev3[pos1];
        if (c2 > match_min_pos && t.buf[c2] == cur[0]) {
            max_len = 2;
            pairs.put(cnt++, 2, pos1 - c2 - 1);
        }
        if (c3 > match_min_pos && t.buf[c3] == cur[0]) {
            if (c3 == c2) cnt--;
            max_len = 3;
            pairs.put(cnt++, 3, pos1 - c3 - 1);
            c2 = c3;
        }
        if (cnt != 0 && c2 == cur_match) {
            cnt--;
            max_len = 1;
        }
    } else if (cur_match > match_min_pos && t.buf[cur_match + 2] != cur[2]) {
        max_len = 2;
        pairs.put(cnt++, 2, pos1 - cur_match - 1);
    }
}

// write the finished list of a position into the block's temporary area (bump-allocated).  The second half
// of a pair word, the "match + literal + rep0" continuation (Encoder.java:766-770), is filled in by the
// compaction pass (lzb_list_gather): it is a function of the data only, costs a memory round trip per pair,
// and a bucket's insertions are a serial chain that must not wait for it.
__device__ __forceinline__ void tree_store_list(const MfWave& w, const TreeBlock& t, uint32_t b, uint32_t pos1, const PairBuf& pairs,
                                                uint32_t cnt) {
    uint32_t where = kMfEmpty;
    if (cnt) {
        const uint32_t off = atomicAdd(&w.pair_used[b], cnt + 1);
        if (off + cnt + 1 <= w.pair_cap) {
            where = off;
            t.pairs_out[off] = cnt;
            for (uint32_t i = 0; i < cnt; i++) {
                const uint32_t len = pairs.len[i];
                t.pairs_out[off + 1 + i] = pair_word(len, pairs.dist[i]);
                t.pairs2_out[off + 1 + i] = pair2_word(len, 0);
            }
        } else {
            atomicMax(w.overflow, 1u);
        }
    }
    t.idx[pos1] = where;
    t.cnt[pos1] = (uint16_t)cnt;
}

// the continuation of one pair of the list of 1-based position pos1: GetMatchLen(len, dist, fb) one byte after the match
__device__ __forceinline__ uint32_t pair_continuation(const uint8_t* buf1, uint32_t n, uint32_t fb, uint32_t pos1, uint32_t len, uint32_t dist) {
    const uint32_t s1 = pos1 + len + 1;  // 1-based start of the continuation
    if (s1 > n) return 0;
    uint32_t lim = n + 1 - s1;
    const uint32_t room = lim;
    if (lim > fb) lim = fb;
    const uint8_t* a = buf1 + s1;
    return extend_run(a, a - dist - 1, 0, lim, room);
}

// One thread per hash-4 bucket.  A bucket with more than kLongChain positions hands the rest of its
// chain to lzb_mf_long_kernel, which pipelines the insertions of one chain across a warp.
__global__ void __launch_bounds__(256) lzb_mf_tree_kernel(MfWave w) {
    const uint32_t b = blockIdx.y;
    if (w.in_len[b] > (uint64_t)w.np - 1) return;  // flagged by the link kernel
    const uint32_t n = (uint32_t)w.in_len[b];
    const uint32_t p0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (p0 >= n) return;
    uint32_t pos1 = p0 + 1;
    if (!((w.prev2 + (size_t)b * w.np)[pos1] & kHeadFlag)) return;  // not the first position of a bucket

    const TreeBlock t(w, b);
    uint32_t* son = t.son;
    const uint32_t direct = w.bt4 ? 0 : 2;  // kNumHashDirectBytes

    PairBuf pairs;
    uint32_t cur_match = 0;  // the hash-4 head: previous position of this bucket (0 = kEmptyHashValue)
    uint32_t done = 0;
    while (pos1 != 0) {
        if (done == kLongChain) {
            const uint32_t k = atomicAdd(w.long_count, 1u);
            w.long_items[k] = make_uint4(b, cur_match, pos1, 0);
            return;
        }
        done++;
        const uint32_t nxt = t.next[pos1];
        const uint32_t remaining = n - (pos1 - 1);
        const uint32_t len_limit = remaining < (uint32_t)w.fb ? remaining : (uint32_t)w.fb;  // BinTree.java:153-162
        const uint32_t match_min_pos = pos1 > w.cyclic_size ? pos1 - w.cyclic_size : 0;     // :164
        const uint8_t* cur = t.buf + pos1;
        uint32_t max_len, cnt;
        tree_prepairs(w, t, pos1, cur_match, match_min_pos, pairs, cnt, max_len);

        uint32_t ptr0 = 2 * pos1 + 1, ptr1 = 2 * pos1;
        uint32_t len0 = direct, len1 = direct;
        int32_t count = w.cut;
        uint32_t cm = cur_match;
        for (;;) {  // :230-270
            if (cm <= match_min_pos || count-- == 0) {
                son[ptr0] = 0;
                son[ptr1] = 0;
                break;
            }
            const uint8_t* pby1 = t.buf + cm;
            // both children of the candidate in one 8-byte load, issued together with the first byte
            // compare: one dependent memory round trip per tree level instead of two.  (The slots
            // written below belong to other nodes, never to `cm`, so reading early is safe.)
            const uint2 kids = *reinterpret_cast<const uint2*>(son + 2 * cm);
            uint32_t len = len0 < len1 ? len0 : len1;
            if (pby1[len] == cur[len]) {
                len = extend_run(cur, pby1, len + 1, len_limit, remaining);
                if (max_len < len) {
                    max_len = len;
                    pairs.put(cnt++, len, pos1 - cm - 1);
                    if (len == len_limit) {
                        son[ptr1] = kids.x;
                        son[ptr0] = kids.y;
                        break;
                    }
                }
            }
            if (pby1[len] < cur[len]) {
                son[ptr1] = cm;
                ptr1 = 2 * cm + 1;
                cm = kids.y;
                len1 = len;
            } else {
                son[ptr0] = cm;
                ptr0 = 2 * cm;
                cm = kids.x;
                len0 = len;
            }
        }
        tree_store_list(w, t, b, pos1, pairs, cnt);
        cur_match = pos1;
        pos1 = nxt;
    }
}

// Long buckets: one warp per chain, a rolling window of 32 consecutive insertions in flight.
// The insertion of a position rewrites links strictly top-down: at any time it owns exactly two
// "pending" slots (the reference's ptr0 / ptr1), every link above them is final, everything below
// is untouched.  Pending slots hold kPending; a later insertion that needs such a link waits (it
// retries in the next round), so it can never overtake an earlier one, and each insertion sees
// exactly the tree the sequential order would have shown it.
// A lane that finishes its insertion takes the next position of the chain at once (a drained batch
// would leave the pipeline half empty: a frequent 4-gram in sorted-ish data degenerates the tree, every
// insertion walks `cut` levels, and the chain's time is rounds per insertion x chain length).  The chain
// itself is a linked list (`next`): one pointer-chasing load per round runs ahead of the insertions and
// feeds a 32-entry FIFO held one entry per lane, and a round's loads -- the candidate's two links, the
// window bytes, the chase -- are issued together, so a round costs one memory round trip.
constexpr uint32_t kPending = 0xFFFFFFFFu;

__global__ void __launch_bounds__(64) lzb_mf_long_kernel(MfWave w) {
    const int lane = threadIdx.x & 31;
    const uint32_t direct = w.bt4 ? 0 : 2;
    PairBuf pairs;
    for (;;) {
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(w.long_ticket, 1u);
        item = __shfl_sync(kFull, item, 0);
        if (item >= *w.long_count) break;
        const uint4 it = w.long_items[item];
        const uint32_t b = it.x;
        const TreeBlock t(w, b);
        volatile uint32_t* son = t.son;
        auto put = [&](uint32_t slot, uint32_t v) { son[slot] = v; };
        // the chain ahead: FIFO entry i lives in lane (i & 31)
        uint32_t q = lane == 0 ? it.z : 0u;
        uint32_t qhead = 0, qcount = 1, tailpos = it.z, last_assigned = it.y;
        bool chain_end = false;
        // the insertion this lane is working on
        bool done = true;
        uint32_t pos1 = 0, len_limit = 0, match_min_pos = 0, max_len = 1, cnt = 0, remaining = 0;
        uint32_t ptr0 = 0, ptr1 = 0, len0 = direct, len1 = direct, cm = 0, len = 0;
        int32_t count = 0;
        bool have_cmp = false;
        for (;;) {
            // the chase: one link ahead per round (every lane loads the same word)
            const bool do_chase = !chain_end && qcount < 32;
            uint32_t chased = 0;
            if (do_chase) chased = t.next[tailpos];
            // free lanes take the next positions of the chain, in order
            const unsigned free_m = __ballot_sync(kFull, done);
            uint32_t take = (uint32_t)__popc(free_m);
            if (take > qcount) take = qcount;
            if (take) {
                const uint32_t r = (uint32_t)__popc(free_m & ((1u << lane) - 1u));  // rank among the free lanes
                const uint32_t p = __shfl_sync(kFull, q, (qhead + r) & 31u);
                const uint32_t before = __shfl_sync(kFull, q, (qhead + r + 31u) & 31u);
                const uint32_t newest = __shfl_sync(kFull, q, (qhead + take - 1u) & 31u);
                if (done && r < take) {
                    pos1 = p;
                    cm = r ? before : last_assigned;  // the bucket's previous position is where the descent starts
                    remaining = t.n - (pos1 - 1);
                    len_limit = remaining < (uint32_t)w.fb ? remaining : (uint32_t)w.fb;
                    match_min_pos = pos1 > w.cyclic_size ? pos1 - w.cyclic_size : 0;
                    tree_prepairs(w, t, pos1, cm, match_min_pos, pairs, cnt, max_len);
                    ptr0 = 2 * pos1 + 1;
                    ptr1 = 2 * pos1;
                    son[ptr0] = kPending;
                    son[ptr1] = kPending;
                    len0 = len1 = direct;
                    count = w.cut;
                    have_cmp = false;
                    done = false;
                }
                last_assigned = newest;
                qhead += take;
                qcount -= take;
            }
            __syncwarp();
            if (!done) {
                if (cm <= match_min_pos || count == 0) {  // BinTree.java:231-235
                    put(ptr0, 0);
                    put(ptr1, 0);
                    done = true;
                } else {
                    const uint32_t kx = son[2 * cm], ky = son[2 * cm + 1];  // issued with the window bytes below
                    const uint8_t* cur = t.buf + pos1;
                    const uint8_t* pby1 = t.buf + cm;
                    if (!have_cmp) {  // kept across retries
                        len = len0 < len1 ? len0 : len1;
                        if (pby1[len] == cur[len]) len = extend_run(cur, pby1, len + 1, len_limit, remaining);
                        have_cmp = true;
                    }
                    const bool full = len == len_limit && max_len < len;  // :249-256 ends the insertion
                    const bool right = !full && pby1[len] < cur[len];
                    const bool ready = full ? (kx != kPending && ky != kPending) : (right ? ky != kPending : kx != kPending);
                    if (ready) {
                        count--;
                        have_cmp = false;
                        if (max_len < len) {
                            max_len = len;
                            pairs.put(cnt++, len, pos1 - cm - 1);
                        }
                        if (full) {
                            put(ptr1, kx);
                            put(ptr0, ky);
                            done = true;
                        } else if (right) {
                            put(ptr1, cm);
                            ptr1 = 2 * cm + 1;
                            put(ptr1, kPending);
                            cm = ky;
                            len1 = len;
                        } else {
                            put(ptr0, cm);
                            ptr0 = 2 * cm;
                            put(ptr0, kPending);
                            cm = kx;
                            len0 = len;
                        }
                    }
                }
                if (done) tree_store_list(w, t, b, pos1, pairs, cnt);
            }
            __syncwarp();
            if (do_chase) {
                if (chased) {
                    if (lane == (int)((qhead + qcount) & 31u)) q = chased;
                    qcount++;
                    tailpos = chased;
                } else {
                    chain_end = true;
                }
            }
            if (chain_end && qcount == 0 && __all_sync(kFull, done)) break;
        }
    }
}

// ---- list compaction: temporary lists -> position-ordered lists in the wave's pool -----------------
// The tree threads allocate their lists in the order they finish, so the lists of neighbouring
// positions are scattered over the block's temporary area.  The parser walks positions in order;
// compacting the lists in position order turns its three reads per position (idx, pairs, pair2) into
// forward walks (measured on the parser: DRAM traffic per input byte, profiles/r02_*), sizes the
// resident lists exactly (6 bytes per pair word instead of a worst-case 36 bytes per input byte), and
// frees the match finder's scratch for the next group of blocks.
__device__ __forceinline__ uint32_t cta_exclusive_scan_256(uint32_t v, uint32_t* s_warp, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(kFull, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < 8 ? s_warp[lane] : 0, z = w;
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            const uint32_t y = __shfl_up_sync(kFull, z, d);
            if (lane >= d) z += y;
        }
        if (lane < 8) s_warp[lane] = z - w;
        if (lane == 7) s_warp[8] = z;
    }
    __syncthreads();
    const uint32_t r = s_warp[warp] + x - v;
    if (total) *total = s_warp[8];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(256) lzb_list_tile_sums(MfWave w, uint32_t tiles_max, uint32_t* tile_sum) {
    __shared__ uint32_t s_warp[9];
    const uint32_t b = blockIdx.y, t = blockIdx.x;
    if (w.in_len[b] > (uint64_t)w.np - 1) return;  // flagged by the link kernel
    const uint32_t n = (uint32_t)w.in_len[b];
    if ((uint64_t)t * kListTile >= n) return;
    const uint16_t* cnt = w.cnt + (size_t)b * w.np;
    const uint32_t p0 = t * kListTile + threadIdx.x * 8;
    uint32_t s = 0;
#pragma unroll
    for (uint32_t k = 0; k < 8; k++) {
        const uint32_t p = p0 + k;
        const uint32_t c = p < n ? cnt[p + 1] : 0;
        s += c ? c + 1 : 0;
    }
    uint32_t total;
    cta_exclusive_scan_256(s, s_warp, &total);
    if (threadIdx.x == 0) tile_sum[(size_t)b * tiles_max + t] = total;
}

__global__ void __launch_bounds__(256) lzb_list_scan_tiles(MfWave w, uint32_t tiles_max, uint32_t* tile_sum, uint32_t* w_total) {
    __shared__ uint32_t s_warp[9];
    const uint32_t b = blockIdx.x;
    const uint32_t n = w.in_len[b] > (uint64_t)w.np - 1 ? 0u : (uint32_t)w.in_len[b];
    const uint32_t tiles = (n + kListTile - 1) / kListTile;
    uint32_t* ts = tile_sum + (size_t)b * tiles_max;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < tiles; base += 256) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < tiles ? ts[i] : 0;
        uint32_t total;
        const uint32_t ex = cta_exclusive_scan_256(v, s_warp, &total);
        if (i < tiles) ts[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) w_total[b] = carry;
}

__global__ void __launch_bounds__(256) lzb_list_gather(MfWave w, uint32_t tiles_max, const uint32_t* tile_off, const BlockLists* lists,
                                                       uint8_t* pool) {
    __shared__ uint32_t s_warp[9];
    const uint32_t b = blockIdx.y, t = blockIdx.x;
    const uint32_t n = (uint32_t)w.in_len[b];
    if ((uint64_t)t * kListTile >= n) return;
    const uint16_t* cnt = w.cnt + (size_t)b * w.np;
    const uint32_t* idx_tmp = w.idx + (size_t)b * w.np;
    const uint32_t* pairs_tmp = w.pairs + (size_t)b * w.pair_cap;
    const uint16_t* pairs2_tmp = w.pairs2 + (size_t)b * w.pair_cap;
    const BlockLists L = lists[b];
    const uint8_t* buf1 = w.in + w.in_off[b] - 1;  // buf1[pos1] is the byte at 1-based position pos1
    uint32_t* idx = reinterpret_cast<uint32_t*>(pool + L.idx_off);
    uint32_t* pairs = reinterpret_cast<uint32_t*>(pool + L.pairs_off);
    uint16_t* pairs2 = reinterpret_cast<uint16_t*>(pool + L.pairs2_off);
    const uint32_t p0 = t * kListTile + threadIdx.x * 8;
    uint32_t c[8], s = 0;
#pragma unroll
    for (uint32_t k = 0; k < 8; k++) {
        const uint32_t p = p0 + k;
        c[k] = p < n ? cnt[p + 1] : 0;
        s += c[k] ? c[k] + 1 : 0;
    }
    uint32_t off = tile_off[(size_t)b * tiles_max + t] + cta_exclusive_scan_256(s, s_warp, nullptr);
#pragma unroll 1
    for (uint32_t k = 0; k < 8; k++) {
        const uint32_t p = p0 + k;
        if (p >= n) break;
        if (c[k] == 0) {
            idx[p + 1] = kMfEmpty;
            continue;
        }
        idx[p + 1] = off;
        const uint32_t from = idx_tmp[p + 1];
        pairs[off] = c[k];
        pairs2[off] = 0;
        for (uint32_t i = 1; i <= c[k]; i++) {
            const uint32_t wd = pairs_tmp[from + i], w2 = pairs2_tmp[from + i];
            const uint32_t len = pair_len(wd, w2);
            pairs[off + i] = wd;
            pairs2[off + i] = pair2_word(len, pair_continuation(buf1, n, (uint32_t)w.fb, p + 1, len, pair_dist(wd)));
        }
        off += c[k] + 1;
    }
}

cudaError_t launch_list_scan(const MfWave& w, uint32_t max_len, uint32_t* tile_sum, uint32_t* w_total, cudaStream_t st) {
    if (w.n_blocks == 0) return cudaSuccess;
    const uint32_t tiles_max = (max_len + kListTile - 1) / kListTile;
    if (tiles_max) {
        lzb_list_tile_sums<<<dim3(tiles_max, w.n_blocks), 256, 0, st>>>(w, tiles_max, tile_sum);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    lzb_list_scan_tiles<<<w.n_blocks, 256, 0, st>>>(w, tiles_max ? tiles_max : 1, tile_sum, w_total);
    return cudaGetLastError();
}

cudaError_t launch_list_gather(const MfWave& w, uint32_t max_len, const uint32_t* tile_off, const BlockLists* lists, uint8_t* pool,
                               cudaStream_t st) {
    const uint32_t tiles_max = (max_len + kListTile - 1) / kListTile;
    if (w.n_blocks == 0 || tiles_max == 0) return cudaSuccess;
    lzb_list_gather<<<dim3(tiles_max, w.n_blocks), 256, 0, st>>>(w, tiles_max, tile_off, lists, pool);
    return cudaGetLastError();
}

// `ev` (optional, LZB_ENC_TIMING): three events recorded after the link, tree and long kernels
cudaError_t launch_mf(const MfWave& w, uint32_t max_len, int num_sms, cudaStream_t st, cudaEvent_t* ev) {
    if (w.n_blocks == 0) return cudaSuccess;
    const uint32_t warps_per_cta = 4;
    lzb_mf_check_lengths<<<(w.n_blocks + 255) / 256, 256, 0, st>>>(w);
    lzb_mf_link_kernel<<<(w.n_blocks + warps_per_cta - 1) / warps_per_cta, warps_per_cta * 32, 0, st>>>(w);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (ev) cudaEventRecord(ev[0], st);
    if (max_len == 0) return cudaSuccess;
    dim3 grid((max_len + 255) / 256, w.n_blocks);
    lzb_mf_tree_kernel<<<grid, 256, 0, st>>>(w);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (ev) cudaEventRecord(ev[1], st);
    // small CTAs: a warp that runs out of chains gives its slot back at once (the last long chains keep a
    // few warps busy for a long time, and the next group's kernels are waiting on the other stream)
    lzb_mf_long_kernel<<<num_sms * 32, 64, 0, st>>>(w);
    if (ev) cudaEventRecord(ev[2], st);
    return cudaGetLastError();
}

}  // namespace lzb
