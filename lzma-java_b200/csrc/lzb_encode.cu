// lzb_encode.cu -- encoder pipeline orchestration: scratch, groups, waves, launches.
//
// Replaces Encoder.Code (LZMA/Encoder.java:1064-1077) for a batch of independent blocks.
//
//   match finder, GROUP by group:   memset(heads, links) -> lzb_mf_link_kernel -> lzb_mf_tree_kernel
//                                   -> lzb_mf_long_kernel -> lzb_list_tile_sums / lzb_list_scan_tiles
//                                   (host: place the group's blocks in the list pool) -> lzb_list_gather
//   parse, WAVE by wave:            lzb_parse_kernel over every block whose lists are in the pool
//
// A group is as many blocks as the match finder's scratch holds (64 bytes per input byte: hash heads,
// links, tree, temporary lists); its output is compacted into the wave's list pool, position-ordered and
// exactly sized (4 bytes per input byte + 6 bytes per pair word: ~27 B/B for text, ~4 B/B for random data),
// and the scratch is reused.  A wave is every group that fits the pool: the parser then sees as many blocks
// at once as possible, handed out longest-expected-first, so that its slots stay busy to the end.
#include <algorithm>
#include <chrono>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "lzb_encode.cuh"

namespace lzb {

namespace {

struct Carver {
    uint8_t* base;
    size_t off = 0;
    template <typename T>
    T* take(size_t count) {
        off = (off + 255) & ~size_t(255);
        T* r = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += count * sizeof(T);
        return r;
    }
};

uint32_t hash_stride_for(int32_t dict, bool bt4, uint32_t* mask_out) {  // BinTree.Create, BinTree.java:113-129
    if (!bt4) {
        *mask_out = 0;
        return kBT2HashSize;
    }
    uint32_t hs = (uint32_t)dict - 1;
    hs |= hs >> 1;
    hs |= hs >> 2;
    hs |= hs >> 4;
    hs |= hs >> 8;
    hs >>= 1;
    hs |= 0xFFFF;
    if (hs > (1u << 24)) hs >>= 1;
    *mask_out = hs;
    return hs + 1 + kHash2Size + kHash3Size;
}

// Group scratch.  With base == nullptr only the size is computed.  *zero_len = bytes at the start that
// must be zero before the match finder runs.
struct GroupScratch {
    uint32_t* tile_sum;
    uint32_t* w_total;
    BlockLists* lists;  // placement of the group's blocks in the pool (host-written)
};
size_t carve_group(uint8_t* base, uint32_t gb, uint32_t np, uint32_t hash_stride, uint32_t pair_cap, MfWave* w, GroupScratch* g,
                   size_t* zero_len) {
    Carver c{base};
    uint32_t* ctl = c.take<uint32_t>(64);  // [0] long-bucket count, [1] long-bucket ticket
    uint32_t* pair_used = c.take<uint32_t>(gb);
    uint32_t* heads = c.take<uint32_t>((size_t)gb * hash_stride);
    uint32_t* next = c.take<uint32_t>((size_t)gb * np);
    if (zero_len) *zero_len = c.off;
    uint32_t* prev2 = c.take<uint32_t>((size_t)gb * np);
    uint32_t* prev3 = c.take<uint32_t>((size_t)gb * np);
    uint32_t* son = c.take<uint32_t>((size_t)gb * 2 * np);
    uint4* long_items = c.take<uint4>((size_t)gb * (np / kLongChain + 1));  // a block has at most n / kLongChain long buckets
    uint32_t* idx = c.take<uint32_t>((size_t)gb * np);
    uint16_t* cnt = c.take<uint16_t>((size_t)gb * np);
    uint32_t* pairs = c.take<uint32_t>((size_t)gb * pair_cap);
    uint16_t* pairs2 = c.take<uint16_t>((size_t)gb * pair_cap);
    const uint32_t tiles_max = (np - 1 + kListTile - 1) / kListTile;
    uint32_t* tile_sum = c.take<uint32_t>((size_t)gb * std::max<uint32_t>(tiles_max, 1));
    uint32_t* w_total = c.take<uint32_t>(gb + 1);  // + the overflow flag, so that one copy brings both back
    BlockLists* lists = c.take<BlockLists>(gb);
    if (w) {
        w->heads = heads;
        w->next = next;
        w->prev2 = prev2;
        w->prev3 = prev3;
        w->son = son;
        w->idx = idx;
        w->cnt = cnt;
        w->pairs = pairs;
        w->pairs2 = pairs2;
        w->pair_used = pair_used;
        w->overflow = w_total + gb;
        w->long_count = ctl;
        w->long_ticket = ctl + 1;
        w->long_items = long_items;
    }
    if (g) {
        g->tile_sum = tile_sum;
        g->w_total = w_total;
        g->lists = lists;
    }
    return (c.off + 255) & ~size_t(255);
}

// Per-handle areas that live as long as the parser runs: its ticket, the block order, the placement of
// every block of the wave, and per resident slot the _optimum spill area and (maybe) the literal coders.
struct Fixed {
    uint32_t* ticket;
    uint32_t* order;
    BlockLists* lists;
    void* opt;
    uint16_t* lit;
};
size_t carve_fixed(uint8_t* base, uint32_t max_wave, size_t slots, size_t lit_slots, Fixed* f) {
    Carver c{base};
    uint32_t* ticket = c.take<uint32_t>(64);
    uint32_t* order = c.take<uint32_t>(max_wave);
    BlockLists* lists = c.take<BlockLists>(max_wave);
    void* opt = c.take<uint8_t>(slots * parse_opt_bytes_per_slot());
    uint16_t* lit = c.take<uint16_t>(lit_slots);
    if (f) {
        f->ticket = ticket;
        f->order = order;
        f->lists = lists;
        f->opt = opt;
        f->lit = lit;
    }
    return (c.off + 255) & ~size_t(255);
}

inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }
// bytes of the pool a block of n input bytes with w pair words takes
inline size_t pool_need(uint64_t n, uint64_t w) {
    return align256(4 * (n + 1)) + align256(4 * (w + kListSlack)) + align256(2 * (w + kListSlack));
}

cudaError_t grow(void** p, size_t* cap, size_t need, cudaStream_t st) {
    if (need <= *cap) return cudaSuccess;
    if (*p) {
        cudaStreamSynchronize(st);
        cudaFree(*p);
        *p = nullptr;
        *cap = 0;
    }
    cudaError_t e = cudaMalloc(p, need);
    if (e == cudaSuccess) *cap = need;
    return e;
}

uint32_t pair_cap_for(uint32_t pair_mul, uint64_t max_in_len) {
    return (uint32_t)std::min<uint64_t>((uint64_t)pair_mul * max_in_len + 4096, 0xFFFFFFF0ull);
}

struct Plan {
    uint32_t hash_stride, hash_mask, np;
    int dic_log;
    ParseGeometry geo;
    size_t lit_per_slot;  // literal coder in global memory (0 when it lives in shared memory)
};

// the wave under construction (host side)
struct Wave {
    uint32_t first = 0;               // index of its first block in the batch
    std::vector<BlockLists> lists;    // placement of its blocks in the pool
    std::vector<uint64_t> cost;       // parser cost estimate per block: bytes + pair words
    size_t pool_used = 0;
};

cudaError_t parse_wave(const EncodeArgs& a, const Plan& P, const Fixed& F, const uint8_t* pool, Wave& wv, int num_sms, cudaStream_t st,
                       int* nl, unsigned long long* h_progress) {
    const uint32_t wb = (uint32_t)wv.lists.size();
    if (wb == 0) return cudaSuccess;
    cudaError_t e;
    ParseArgs pa;
    pa.in = a.in;
    pa.in_off = a.in_off + wv.first;
    pa.in_len = a.in_len + wv.first;
    pa.n_blocks = wb;
    pa.pool = pool;
    pa.lists = F.lists;
    pa.out = a.out;
    pa.out_off = a.out_off + wv.first;
    pa.out_cap = a.out_cap + wv.first;
    pa.out_len = a.out_len + wv.first;
    pa.ticket = F.ticket;
    pa.opt_scratch = F.opt;
    pa.lit_scratch = F.lit;
    pa.dict_size = a.dict_size;
    pa.dist_table_size = P.dic_log * 2;
    pa.lc = a.lc;
    pa.lp = a.lp;
    pa.pb = a.pb;
    pa.fb = a.fb;
    pa.eos = a.eos;
    pa.with_header = a.with_header;
    pa.slice_bytes = P.geo.slice_bytes;
    pa.lit_in_smem = P.geo.lit_in_smem;
    pa.progress = a.progress_fn ? h_progress : nullptr;
    e = cudaMemsetAsync(F.ticket, 0, 64 * sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyAsync(F.lists, wv.lists.data(), (size_t)wb * sizeof(BlockLists), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    // one CTA per SM (every SM gets work, blocks are drawn by ticket), as many warps as the wave can fill
    const int grid = (int)std::min<uint32_t>(wb, (uint32_t)num_sms);
    int warps = (int)((wb + (uint32_t)grid - 1) / (uint32_t)grid);
    warps = std::min(std::max(warps, 1), P.geo.max_warps);
    if (a.tune_warps > 0) warps = std::min((int)a.tune_warps, P.geo.max_warps);  // tuning knob
    // More blocks than parser slots: a block that starts late must not be a slow one, or the wave ends on it
    // with the GPU idle.  The parser's cost grows with the bytes to code and with the match pairs it has to
    // price, so blocks are handed out by decreasing (length + pair words).  With fewer blocks the same order
    // still pays: the warps of a CTA draw neighbouring tickets, so an SM's streams are alike and run the
    // same parts of the kernel, which its instruction cache likes (profiles/r01_parse_kernel_mixed_w8_w2_ncu.txt).
    pa.order = nullptr;
    std::vector<uint32_t> order;
    if (wb > (uint32_t)num_sms && !a.tune_fifo) {
        order.resize(wb);
        for (uint32_t i = 0; i < wb; i++) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return wv.cost[x] > wv.cost[y]; });
        // One stream fewer per SM when the wave can afford it.  The streams of an SM share its instruction fetch, so
        // every stream runs faster with one neighbour less (2048 mixed blocks: parse 4.33 s at 14 per SM, 4.10 s at 13);
        // the L blocks that then have no slot of their own are the cheapest ones and wait for the first slots to free up.
        // That pays only if they are done before the expensive blocks are: the L cheapest blocks, each queued behind one
        // of the L next-cheapest, must fit well inside the most expensive block's time (the cost is a rough estimate,
        // hence the margin); a wave of equal blocks keeps every slot.
        if (a.tune_warps <= 0 && warps >= 4 && warps == P.geo.max_warps && wb <= (uint32_t)grid * (uint32_t)warps &&
            wb > (uint32_t)grid * (uint32_t)(warps - 1)) {
            const uint32_t L = wb - (uint32_t)grid * (uint32_t)(warps - 1);
            if (2 * L <= wb) {
                const uint64_t c_max = wv.cost[order[0]];
                const uint64_t c_left = wv.cost[order[wb - L]], c_early = wv.cost[order[wb - 2 * L]];  // dearest of each set
                if ((c_left + c_early) * 10 <= c_max * 7) warps--;
            }
        }
        if (!a.tune_blocked) {
            // Warp w of CTA c starts on order[c * warps + w].  At 12-14 streams per SM the parser is bound by the SM's
            // instruction issue, so what counts is that every SM gets the same amount of work: the blocks that start at
            // once are dealt to the CTAs like cards, heaviest first, back and forth (CTA 0..G-1, G-1..0, ...), each CTA
            // taking as many as it has warps below n_blocks.  Blocks drawn later by ticket keep the decreasing order.
            const uint32_t first = std::min<uint32_t>(wb, (uint32_t)grid * (uint32_t)warps);
            std::vector<uint32_t> dealt(first), fill((size_t)grid, 0);
            uint32_t k = 0;
            for (uint32_t pass = 0; k < first; pass++) {
                for (int i = 0; i < grid && k < first; i++) {
                    const uint32_t c = (pass & 1) ? (uint32_t)(grid - 1 - i) : (uint32_t)i;
                    const uint32_t p = c * (uint32_t)warps + fill[c];
                    if (fill[c] >= (uint32_t)warps || p >= first) continue;
                    dealt[p] = order[k++];
                    fill[c]++;
                }
            }
            std::copy(dealt.begin(), dealt.end(), order.begin());
        }
        e = cudaMemcpyAsync(F.order, order.data(), (size_t)wb * sizeof(uint32_t), cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return e;
        pa.order = F.order;
    }
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    if (a.tune_timing) {
        cudaEventCreate(&t0);
        cudaEventCreate(&t1);
        cudaEventRecord(t0, st);
    }
    e = launch_parse(pa, grid, warps, st);
    if (e != cudaSuccess) return e;
    *nl += 1;
    if (a.tune_timing) cudaEventRecord(t1, st);
    if (a.progress_fn && h_progress) {
        // ICodeProgress: the streams add to two counters in pinned memory; report them while the kernel runs
        volatile unsigned long long* hp = h_progress;
        unsigned long long seen = hp[0];
        while (cudaStreamQuery(st) == cudaErrorNotReady) {
            const unsigned long long in_now = hp[0], out_now = hp[1];
            if (in_now != seen) {
                seen = in_now;
                a.progress_fn(a.progress_user, in_now, out_now);
            }
            std::this_thread::sleep_for(std::chrono::microseconds(200));
        }
    }
    // the host vectors above are pageable and die with this scope; the pool is reused by the next wave
    e = cudaStreamSynchronize(st);
    if (a.tune_timing) {
        float ms = 0;
        cudaEventElapsedTime(&ms, t0, t1);
        fprintf(stderr, "lzb_enc wave: %u blocks, %d warps x %d CTAs, list pool %.2f GB, parse %.1f ms\n", wb, warps, grid,
                wv.pool_used / 1e9, ms);
        cudaEventDestroy(t0);
        cudaEventDestroy(t1);
    }
    wv.first += wb;
    wv.lists.clear();
    wv.cost.clear();
    wv.pool_used = 0;
    return e;
}

}  // namespace

void EncScratch::release() {
    for (int k = 0; k < kEncGroupsInFlight; k++) {
        if (gp[k]) cudaFree(gp[k]);
        gp[k] = nullptr;
        gcap[k] = 0;
        if (gs[k]) cudaStreamDestroy(gs[k]);
        gs[k] = nullptr;
        if (gev[k]) cudaEventDestroy(gev[k]);
        gev[k] = nullptr;
    }
    streams_ready = false;
    if (pool) cudaFree(pool);
    if (fixed) cudaFree(fixed);
    if (h_totals) cudaFreeHost(h_totals);
    if (h_progress) cudaFreeHost(h_progress);
    h_progress = nullptr;
    pool = fixed = nullptr;
    h_totals = nullptr;
    pool_cap = fixed_cap = h_totals_cap = 0;
}

cudaError_t run_encode(const EncodeArgs& a, EncScratch& scratch, int num_sms, cudaStream_t st, int* launches, MfTrace* mf_only) {
    int nl = 0;
    if (launches) *launches = 0;
    if (a.n == 0) return cudaSuccess;
    if (a.max_in_len > kEncMaxBlock) return cudaErrorNotSupported;
    cudaError_t e = upload_mf_tables();
    if (e != cudaSuccess) return e;

    if (a.progress_fn) {
        if (!scratch.h_progress) {
            e = cudaHostAlloc((void**)&scratch.h_progress, 2 * sizeof(unsigned long long), cudaHostAllocMapped);
            if (e != cudaSuccess) return e;
        }
        scratch.h_progress[0] = scratch.h_progress[1] = 0;
    }
    Plan P;
    P.hash_stride = hash_stride_for(a.dict_size, a.bt4, &P.hash_mask);
    P.np = (uint32_t)a.max_in_len + 1;
    P.dic_log = 0;
    while ((uint32_t)a.dict_size > (1u << P.dic_log)) P.dic_log++;  // Encoder.java:1141-1144
    // streams per SM and where the literal coders live follow from how many blocks want to run at once
    uint32_t per_sm = (a.n + (uint32_t)num_sms - 1) / (uint32_t)num_sms;
    if (a.tune_warps > 0) per_sm = std::min<uint32_t>(per_sm, (uint32_t)a.tune_warps);
    P.geo = parse_geometry(a.lc, a.lp, a.pb, a.fb, per_sm, a.tune_lit);
    P.lit_per_slot = P.geo.lit_in_smem ? 0 : ((size_t)0x300 << (a.lc + a.lp));
    const size_t slots = (size_t)num_sms * P.geo.max_warps;  // one CTA per SM, up to max_warps streams each

    size_t free_b = 0, total_b = 0;
    e = cudaMemGetInfo(&free_b, &total_b);
    if (e != cudaSuccess) return e;
    // What this call may hold at most: 85 % of what is free (plus what the handle already holds).  It takes
    // what the batch needs, not the budget: the pool is sized from the batch (below) and grows only on demand.
    size_t held = scratch.pool_cap + scratch.fixed_cap;
    for (int k = 0; k < kEncGroupsInFlight; k++) held += scratch.gcap[k];
    const size_t budget = std::max<size_t>((free_b + held) / 20 * 17, size_t(512) << 20);

    uint32_t pair_mul = 6;  // temporary pair slots per input byte (text needs ~4.4); doubled when a group overflows
    if (a.tune_pair_mul > 0) pair_mul = (uint32_t)a.tune_pair_mul;  // test knob for the retry path

    // ---- fixed areas
    Fixed F;
    const size_t fixed_bytes = carve_fixed(nullptr, a.n, slots, slots * P.lit_per_slot, nullptr);
    e = grow(&scratch.fixed, &scratch.fixed_cap, fixed_bytes, st);
    if (e != cudaSuccess) return e;
    carve_fixed((uint8_t*)scratch.fixed, a.n, slots, slots * P.lit_per_slot, &F);

    // ---- groups: K of them are in flight, each with its own scratch and stream; together they may take 45 % of
    // the budget.  Many small groups beat few large ones: lzb_mf_long_kernel ends on its longest hash bucket (a
    // serial chain: ~0.25 s per MiB of ordered records, whatever the group's size), and while a few warps finish
    // those chains the other groups' kernels fill the GPU.
    int K = mf_only ? 1 : kEncGroupsDefault;
    if (!mf_only && a.tune_inflight > 0) K = std::min<int>(a.tune_inflight, kEncGroupsInFlight);
    auto group_bytes = [&](uint32_t gb, uint32_t pm) {
        return carve_group(nullptr, gb, P.np, P.hash_stride, pair_cap_for(pm, a.max_in_len), nullptr, nullptr, nullptr);
    };
    uint32_t gmax = std::min<uint32_t>(a.n, 32768);
    // the groups get 45 % of the budget, or whatever the list pool is not expected to need if that is more (the match
    // finder's time goes with the number of rounds: 2048 x 1 MiB in 2 rounds of 6 x 171 instead of 3 of 6 x 114)
    const uint64_t pool_per_block = pool_need(a.max_in_len, (uint64_t)(4.5 * (double)a.max_in_len));
    const uint64_t pool_expect = std::min<uint64_t>((uint64_t)a.n * pool_per_block, (uint64_t)1 << 46);
    const size_t avail = budget - std::min(budget, fixed_bytes);
    size_t groups_total = avail / 20 * 9;
    if (!mf_only && avail > pool_expect) groups_total = std::min<size_t>(std::max<size_t>(groups_total, avail - (size_t)pool_expect), avail / 10 * 7);
    const size_t group_budget = std::max<size_t>(groups_total / (size_t)K, group_bytes(1, pair_mul));
    if (group_bytes(gmax, pair_mul) > group_budget) {
        uint32_t lo = 1, hi = gmax;  // group_bytes is monotonic: binary search the largest count that fits
        while (lo < hi) {
            const uint32_t mid = lo + (hi - lo + 1) / 2;
            if (group_bytes(mid, pair_mul) <= group_budget) lo = mid;
            else hi = mid - 1;
        }
        gmax = lo;
    }
    if (!mf_only) {
        // Whole rounds of K equal groups: the last groups of a ragged batch would otherwise run alone, on a GPU that the
        // match finder's serial chains cannot fill (2048 blocks as 13 x 154 + 46: the last two groups took 0.5 s by themselves).
        const uint32_t rounds = (a.n + gmax * (uint32_t)K - 1) / (gmax * (uint32_t)K);
        gmax = std::max<uint32_t>((a.n + rounds * (uint32_t)K - 1) / (rounds * (uint32_t)K), 1);
    }
    if (a.tune_group > 0) gmax = std::min<uint32_t>(gmax, (uint32_t)a.tune_group);  // test knob: small groups

    // ---- list pool: what the batch is expected to need (4 B per input byte + 6 B per pair word at 4.5 pair words
    // per byte), bounded by what the budget leaves; a wave ends when the next group does not fit
    size_t pool_cap = scratch.pool_cap;
    if (!mf_only) {
        const uint64_t expect = pool_expect;
        const size_t gbytes = (size_t)K * group_bytes(gmax, pair_mul);
        const size_t room = budget > fixed_bytes + gbytes ? budget - fixed_bytes - gbytes : 0;
        size_t want = (size_t)std::min<uint64_t>(expect, room);
        want = std::max<size_t>(want, pool_need(a.max_in_len, (uint64_t)pair_cap_for(pair_mul, a.max_in_len)));  // any one block fits
        if (a.tune_pool > 0) want = std::min<size_t>(want, (size_t)a.tune_pool);  // test knob: tiny pool, many waves
        e = grow(&scratch.pool, &scratch.pool_cap, want, st);
        if (e != cudaSuccess) return e;
        pool_cap = a.tune_pool > 0 ? std::min<size_t>(scratch.pool_cap, (size_t)a.tune_pool) : scratch.pool_cap;
    }
    uint8_t* pool = (uint8_t*)scratch.pool;

    Wave wv;
    std::vector<uint64_t> lens(a.n);
    e = cudaMemcpyAsync(lens.data(), a.in_len, (size_t)a.n * sizeof(uint64_t), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return e;

    // K groups are in flight on K streams, each with its own scratch (see "groups" above)
    struct InFlight {
        uint32_t first = 0, gb = 0, pair_cap = 0;
        int k = 0;
        MfWave w;
        GroupScratch G;
        cudaEvent_t mev[5];
        std::vector<BlockLists> place;  // host copy of the group's placement (must outlive the copy below)
    };
    if (!scratch.streams_ready) {
        for (int k = 0; k < kEncGroupsInFlight; k++) {
            if (k) {
                e = cudaStreamCreateWithFlags(&scratch.gs[k], cudaStreamNonBlocking);
                if (e != cudaSuccess) return e;
            }
            e = cudaEventCreateWithFlags(&scratch.gev[k], cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
        }
        scratch.streams_ready = true;
    }
    if (scratch.h_totals_cap < (size_t)gmax + 1) {
        if (scratch.h_totals) cudaFreeHost(scratch.h_totals);
        scratch.h_totals = nullptr;
        scratch.h_totals_cap = 0;
        e = cudaHostAlloc((void**)&scratch.h_totals, (size_t)kEncGroupsInFlight * ((size_t)gmax + 1) * sizeof(uint32_t), cudaHostAllocDefault);
        if (e != cudaSuccess) return e;
        scratch.h_totals_cap = (size_t)gmax + 1;
    }
    cudaStream_t S[kEncGroupsInFlight];
    S[0] = st;
    for (int k = 1; k < kEncGroupsInFlight; k++) S[k] = scratch.gs[k];
    {   // the side streams start after whatever produced the inputs on the caller's stream
        e = cudaEventRecord(scratch.gev[0], st);
        if (e != cudaSuccess) return e;
        for (int k = 1; k < K; k++) {
            e = cudaStreamWaitEvent(S[k], scratch.gev[0], 0);
            if (e != cudaSuccess) return e;
        }
    }
    auto sync_all = [&]() {
        cudaError_t first = cudaSuccess;
        for (int k = 0; k < K; k++) {
            const cudaError_t ek = cudaStreamSynchronize(S[k]);
            if (first == cudaSuccess) first = ek;
        }
        return first;
    };

    auto enqueue_group = [&](InFlight& f, uint32_t first, int k) -> cudaError_t {
        f.first = first;
        f.gb = std::min<uint32_t>(gmax, a.n - first);
        f.k = k;
        f.pair_cap = pair_cap_for(pair_mul, a.max_in_len);
        cudaError_t err = grow(&scratch.gp[k], &scratch.gcap[k], group_bytes(f.gb, pair_mul), S[k]);
        if (err != cudaSuccess) return err;
        size_t zero_len = 0;
        carve_group((uint8_t*)scratch.gp[k], f.gb, P.np, P.hash_stride, f.pair_cap, &f.w, &f.G, &zero_len);
        MfWave& w = f.w;
        w.in = a.in;
        w.in_off = a.in_off + first;
        w.in_len = a.in_len + first;
        w.n_blocks = f.gb;
        w.np = P.np;
        w.hash_stride = P.hash_stride;
        w.pair_cap = f.pair_cap;
        w.hash_mask = P.hash_mask;
        w.cyclic_size = (uint32_t)a.dict_size + 1;
        w.fb = a.fb;
        w.cut = 16 + (a.fb >> 1);  // BinTree.java:98
        w.bt4 = a.bt4;
        if (a.tune_timing) {
            for (auto& x : f.mev) cudaEventCreate(&x);
            cudaEventRecord(f.mev[0], S[k]);
        }
        err = cudaMemsetAsync(scratch.gp[k], 0, zero_len, S[k]);
        if (err != cudaSuccess) return err;
        err = cudaMemsetAsync(w.overflow, 0, sizeof(uint32_t), S[k]);
        if (err != cudaSuccess) return err;
        err = launch_mf(w, (uint32_t)a.max_in_len, num_sms, S[k], a.tune_timing ? f.mev + 1 : nullptr);
        if (err != cudaSuccess) return err;
        nl += a.max_in_len ? 4 : 2;
        err = launch_list_scan(w, (uint32_t)a.max_in_len, f.G.tile_sum, f.G.w_total, S[k]);
        if (err != cudaSuccess) return err;
        nl += a.max_in_len ? 2 : 1;
        if (a.tune_timing) cudaEventRecord(f.mev[4], S[k]);
        // pair words per block + the overflow flag
        err = cudaMemcpyAsync(scratch.h_totals + (size_t)k * scratch.h_totals_cap, f.G.w_total, (size_t)(f.gb + 1) * sizeof(uint32_t),
                              cudaMemcpyDeviceToHost, S[k]);
        if (err != cudaSuccess) return err;
        return cudaEventRecord(scratch.gev[k], S[k]);
    };

    InFlight ring[kEncGroupsInFlight];
    int head = 0, count = 0;  // ring[head .. head + count) are in flight, oldest first; group i of the ring uses buffer / stream i
    uint32_t cursor = 0;
    while (cursor < a.n || count) {
        while (cursor < a.n && count < K) {
            const int k = (head + count) % K;
            // the buffer is free: the gather of the group that used it was synchronised below
            e = enqueue_group(ring[k], cursor, k);
            if (e != cudaSuccess) return e;
            cursor += ring[k].gb;
            count++;
        }
        InFlight& cur = ring[head];
        {
            e = cudaEventSynchronize(scratch.gev[cur.k]);
            if (e != cudaSuccess) return e;
            const uint32_t* totals = scratch.h_totals + (size_t)cur.k * scratch.h_totals_cap;
            const uint32_t gb = cur.gb;
            if (a.tune_timing) {
                float t_all = 0, t_link = 0, t_tree = 0, t_long = 0, t_scan = 0;
                cudaEventElapsedTime(&t_all, cur.mev[0], cur.mev[4]);
                if (a.max_in_len) {
                    cudaEventElapsedTime(&t_link, cur.mev[0], cur.mev[1]);
                    cudaEventElapsedTime(&t_tree, cur.mev[1], cur.mev[2]);
                    cudaEventElapsedTime(&t_long, cur.mev[2], cur.mev[3]);
                    cudaEventElapsedTime(&t_scan, cur.mev[3], cur.mev[4]);
                }
                fprintf(stderr, "lzb_enc group: %u blocks (max %llu B), match finder %.1f ms (zero+link %.1f tree %.1f long %.1f scan %.1f), "
                                "%d groups in flight\n", gb, (unsigned long long)a.max_in_len, t_all, t_link, t_tree, t_long, t_scan, count);
                for (auto& x : cur.mev) cudaEventDestroy(x);
            }
            const uint32_t overflow = totals[gb];
            if (overflow >= 2) {  // a block longer than the declared max_in_len
                sync_all();
                return cudaErrorInvalidValue;
            }
            if (overflow) {  // rare: this group ran out of temporary pair slots; run it (and what followed) again with twice the room
                e = sync_all();
                if (e != cudaSuccess) return e;
                if (a.tune_timing)
                    for (int i = 1; i < count; i++)
                        for (auto& x : ring[(head + i) % K].mev) cudaEventDestroy(x);
                if (pair_mul >= 512) return cudaErrorMemoryAllocation;
                pair_mul *= 2;
                cursor = cur.first;
                head = 0;
                count = 0;
                continue;
            }
            if (mf_only) {  // trace tap: hand back block 0's temporary lists instead of parsing
                uint32_t used = 0;
                e = cudaMemcpyAsync(&used, cur.w.pair_used, sizeof used, cudaMemcpyDeviceToHost, st);
                if (e != cudaSuccess) return e;
                e = cudaStreamSynchronize(st);
                if (e != cudaSuccess) return e;
                mf_only->idx = cur.w.idx;
                mf_only->pairs = cur.w.pairs;
                mf_only->pairs2 = cur.w.pairs2;
                mf_only->pair_words = used;
                if (launches) *launches = nl;
                return cudaSuccess;
            }
            // place the group's blocks in the pool; when they do not fit, the wave so far is parsed first
            size_t need = 0;
            for (uint32_t i = 0; i < gb; i++) need += pool_need(lens[cur.first + i], totals[i]);
            if (wv.pool_used + need > pool_cap && !wv.lists.empty()) {
                e = sync_all();  // every gather into the pool has finished (and the groups in flight, early)
                if (e != cudaSuccess) return e;
                e = parse_wave(a, P, F, pool, wv, num_sms, st, &nl, scratch.h_progress);
                if (e != cudaSuccess) return e;
            }
            if (need > scratch.pool_cap) {  // one group alone is larger than the pool: enlarge it (nothing is parked in it now)
                e = sync_all();
                if (e != cudaSuccess) return e;
                e = grow(&scratch.pool, &scratch.pool_cap, need + need / 8, st);
                if (e != cudaSuccess) return e;
                pool = (uint8_t*)scratch.pool;
            }
            if (need > pool_cap) pool_cap = scratch.pool_cap;
            std::vector<BlockLists>& place = cur.place;
            place.resize(gb);
            for (uint32_t i = 0; i < gb; i++) {
                const uint64_t n = lens[cur.first + i], wds = totals[i];
                BlockLists L;
                L.idx_off = wv.pool_used;
                L.pairs_off = L.idx_off + align256(4 * (n + 1));
                L.pairs2_off = L.pairs_off + align256(4 * (wds + kListSlack));
                wv.pool_used = L.pairs2_off + align256(2 * (wds + kListSlack));
                place[i] = L;
                wv.lists.push_back(L);
                wv.cost.push_back(n + wds);
            }
            e = cudaMemcpyAsync(cur.G.lists, place.data(), (size_t)gb * sizeof(BlockLists), cudaMemcpyHostToDevice, S[cur.k]);
            if (e != cudaSuccess) return e;
            e = launch_list_gather(cur.w, (uint32_t)a.max_in_len, cur.G.tile_sum, cur.G.lists, pool, S[cur.k]);
            if (e != cudaSuccess) return e;
            nl += a.max_in_len ? 1 : 0;
            // No wait here: the gather is ordered before the next group on this stream (which reuses this scratch),
            // and everything is synchronised before a wave is parsed.
        }
        head = (head + 1) % K;
        count--;
    }
    e = sync_all();
    if (e != cudaSuccess) return e;
    e = parse_wave(a, P, F, pool, wv, num_sms, st, &nl, scratch.h_progress);
    if (launches) *launches = nl;
    return e;
}

}  // namespace lzb
