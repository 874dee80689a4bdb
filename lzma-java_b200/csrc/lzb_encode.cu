// lzb_encode.cu -- encoder pipeline orchestration: scratch carving, waves, launches.
//
// Replaces Encoder.Code (LZMA/Encoder.java:1064-1077) for a batch of
// independent blocks.  Per wave of blocks:  memset(heads, next, counters) ->
// lzb_mf_link_kernel -> lzb_mf_tree_kernel -> lzb_mf_long_kernel -> lzb_parse_kernel.
// A wave holds as many blocks as the memory budget allows.  The parser takes a wave's blocks
// longest-expected-first (most match pairs first): a wave with more blocks than resident parser
// slots then does not end on a late-started slow block, and the streams that share an SM are alike.
#include <algorithm>
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

// Scratch comes in two parts with different lifetimes.  The match-finder part (hash heads, bucket
// links, candidate arrays, tree) is dead once the lists are written; the list part (per-position
// lists, parser spill space) lives until the parse of its blocks has finished.
// With base == nullptr only sizes are computed.  *zero_len = bytes at the start of the part that
// must be zero before the match finder runs.
size_t carve_mf(uint8_t* base, uint32_t wb, uint32_t np, uint32_t hash_stride, MfWave* w, size_t* zero_len) {
    Carver c{base};
    uint32_t* ctl = c.take<uint32_t>(64);                  // [0] long-bucket count, [1] long-bucket ticket
    uint32_t* heads = c.take<uint32_t>((size_t)wb * hash_stride);
    uint32_t* next = c.take<uint32_t>((size_t)wb * np);
    if (zero_len) *zero_len = c.off;
    uint32_t* prev2 = c.take<uint32_t>((size_t)wb * np);
    uint32_t* prev3 = c.take<uint32_t>((size_t)wb * np);
    uint32_t* son = c.take<uint32_t>((size_t)wb * 2 * np);
    uint4* long_items = c.take<uint4>((size_t)wb * (np / kLongChain + 1));  // a block has at most n / kLongChain long buckets
    if (w) {
        w->heads = heads;
        w->next = next;
        w->prev2 = prev2;
        w->prev3 = prev3;
        w->son = son;
        w->long_count = ctl;
        w->long_ticket = ctl + 1;
        w->long_items = long_items;
    }
    return (c.off + 255) & ~size_t(255);
}

size_t carve_lists(uint8_t* base, uint32_t wb, uint32_t np, uint32_t pair_cap, size_t slots, size_t lit_slots, MfWave* w,
                   ParseArgs* pa, size_t* zero_len, uint32_t** order_out = nullptr) {
    Carver c{base};
    uint32_t* ctl = c.take<uint32_t>(64);                  // [0] parser ticket, [1] pair overflow
    uint32_t* pair_used = c.take<uint32_t>(wb);
    if (zero_len) *zero_len = c.off;
    uint32_t* order = c.take<uint32_t>(wb);                // the parser's block order (run_waves)
    if (order_out) *order_out = order;
    uint32_t* idx = c.take<uint32_t>((size_t)wb * np);
    uint32_t* pairs = c.take<uint32_t>((size_t)wb * pair_cap + 64);   // + slack: the parser prefetches 32 slots blindly
    uint16_t* pairs2 = c.take<uint16_t>((size_t)wb * pair_cap + 64);
    void* opt = c.take<uint8_t>(slots * parse_opt_bytes_per_slot());
    uint16_t* lit = c.take<uint16_t>(lit_slots);
    if (w) {
        w->idx = idx;
        w->pairs = pairs;
        w->pairs2 = pairs2;
        w->pair_used = pair_used;
        w->overflow = ctl + 1;
    }
    if (pa) {
        pa->ticket = ctl;
        pa->opt_scratch = opt;
        pa->lit_scratch = lit;
    }
    return (c.off + 255) & ~size_t(255);
}

// ---- what a batch has in common ----------------------------------------------------------
struct Plan {
    uint32_t hash_stride, hash_mask, np;
    int dic_log;
    ParseGeometry geo;
    size_t lit_per_slot;  // literal coder in global memory (0 when it lives in shared memory)
    bool timing;
};

// MfWave + ParseArgs for blocks [first, first + wb): match-finder scratch at `mf_base`, lists at `list_base`
void bind_wave(const EncodeArgs& a, const Plan& P, uint8_t* mf_base, uint8_t* list_base, uint32_t first, uint32_t wb,
               uint32_t pair_cap, size_t slots, MfWave* w, ParseArgs* pa, size_t* mf_zero, size_t* list_zero, uint32_t** order) {
    carve_mf(mf_base, wb, P.np, P.hash_stride, w, mf_zero);
    carve_lists(list_base, wb, P.np, pair_cap, slots, slots * P.lit_per_slot, w, pa, list_zero, order);
    w->in = a.in;
    w->in_off = a.in_off + first;
    w->in_len = a.in_len + first;
    w->n_blocks = wb;
    w->np = P.np;
    w->hash_stride = P.hash_stride;
    w->pair_cap = pair_cap;
    w->hash_mask = P.hash_mask;
    w->cyclic_size = (uint32_t)a.dict_size + 1;
    w->fb = a.fb;
    w->cut = 16 + (a.fb >> 1);  // BinTree.java:98
    w->bt4 = a.bt4;
    pa->out = a.out;
    pa->out_off = a.out_off + first;
    pa->out_cap = a.out_cap + first;
    pa->out_len = a.out_len + first;
    pa->dict_size = a.dict_size;
    pa->dist_table_size = P.dic_log * 2;
    pa->lc = a.lc;
    pa->lp = a.lp;
    pa->pb = a.pb;
    pa->fb = a.fb;
    pa->eos = a.eos;
    pa->with_header = a.with_header;
    pa->slice_bytes = P.geo.slice_bytes;
    pa->lit_in_smem = P.geo.lit_in_smem;
}

cudaError_t grow(EncScratch& scratch, size_t need, cudaStream_t st) {
    if (need <= scratch.cap) return cudaSuccess;
    if (scratch.p) {
        cudaStreamSynchronize(st);
        cudaFree(scratch.p);
        scratch.p = nullptr;
        scratch.cap = 0;
    }
    cudaError_t e = cudaMalloc(&scratch.p, need);
    if (e == cudaSuccess) scratch.cap = need;
    return e;
}

uint32_t pair_cap_for(uint32_t pair_mul, uint64_t max_in_len) {
    return (uint32_t)std::min<uint64_t>((uint64_t)pair_mul * max_in_len + 4096, 0xFFFFFFF0ull);
}

// Blocks [first, first + count) wave by wave on `st`: every wave is match finder -> parse, the
// host waits for the match finder (pair-slot overflow check) and between waves (scratch reuse).
cudaError_t run_waves(const EncodeArgs& a, const Plan& P, EncScratch& scratch, int num_sms, cudaStream_t st, size_t budget,
                      uint32_t first, uint32_t count, uint32_t pair_mul, int* nl, MfTrace* mf_only) {
    const size_t slots = (size_t)num_sms * P.geo.max_warps;  // one CTA per SM, up to max_warps streams each
    cudaError_t e;
    uint32_t done = 0;
    while (done < count) {
        const uint32_t pair_cap = pair_cap_for(pair_mul, a.max_in_len);
        // largest wave that fits the budget, the remaining blocks spread evenly over the waves still to come
        uint32_t wb = std::min<uint32_t>(count - done, 32768);
        auto wave_bytes = [&](uint32_t blocks) {
            return carve_mf(nullptr, blocks, P.np, P.hash_stride, nullptr, nullptr) +
                   carve_lists(nullptr, blocks, P.np, pair_cap, slots, slots * P.lit_per_slot, nullptr, nullptr, nullptr);
        };
        if (wave_bytes(wb) > budget) {
            uint32_t lo = 1, hi = wb;  // wave_bytes is monotonic: binary search the largest count that fits
            while (lo < hi) {
                const uint32_t mid = lo + (hi - lo + 1) / 2;
                if (wave_bytes(mid) <= budget) lo = mid;
                else hi = mid - 1;
            }
            const uint32_t left = count - done, waves = (left + lo - 1) / lo;
            wb = (left + waves - 1) / waves;
        }
        e = grow(scratch, wave_bytes(wb), st);
        if (e != cudaSuccess) return e;
        MfWave w;
        ParseArgs pa;
        size_t mf_zero = 0, list_zero = 0;
        uint8_t* list_base = (uint8_t*)scratch.p + carve_mf(nullptr, wb, P.np, P.hash_stride, nullptr, nullptr);
        uint32_t* order_dev = nullptr;
        bind_wave(a, P, (uint8_t*)scratch.p, list_base, first + done, wb, pair_cap, slots, &w, &pa, &mf_zero, &list_zero, &order_dev);
        // LZB_ENC_TIMING=1 (developer hook): phase times of every wave on stderr
        cudaEvent_t tev[3] = {nullptr, nullptr, nullptr}, mev[4] = {nullptr, nullptr, nullptr, nullptr};
        if (P.timing) {
            for (auto& x : tev) cudaEventCreate(&x);
            for (auto& x : mev) cudaEventCreate(&x);
            cudaEventRecord(tev[0], st);
        }
        e = cudaMemsetAsync(scratch.p, 0, mf_zero, st);
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(list_base, 0, list_zero, st);
        if (e != cudaSuccess) return e;
        if (P.timing) cudaEventRecord(mev[0], st);
        e = launch_mf(w, (uint32_t)a.max_in_len, num_sms, st, P.timing ? mev + 1 : nullptr);
        if (e != cudaSuccess) return e;
        *nl += a.max_in_len ? 3 : 1;

        // did any block run out of pair slots?  (rare: retry the wave with twice the room)
        uint32_t overflow = 0;
        e = cudaMemcpyAsync(&overflow, w.overflow, sizeof overflow, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return e;
        if (P.timing) cudaEventRecord(tev[1], st);
        if (overflow >= 2) return cudaErrorInvalidValue;  // a block longer than the declared max_in_len
        if (overflow) {
            if (P.timing) {
                for (auto& x : tev) cudaEventDestroy(x);
                for (auto& x : mev) cudaEventDestroy(x);
            }
            if (pair_mul >= 512) return cudaErrorMemoryAllocation;
            pair_mul *= 2;
            continue;
        }

        if (mf_only) {  // trace tap: hand back block 0's lists instead of parsing
            uint32_t used = 0;
            e = cudaMemcpyAsync(&used, w.pair_used, sizeof used, cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) return e;
            e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) return e;
            mf_only->idx = w.idx;
            mf_only->pairs = w.pairs;
            mf_only->pairs2 = w.pairs2;
            mf_only->pair_words = used;
            return cudaSuccess;
        }
        pa.mf = w;
        // one CTA per SM (every SM gets work, blocks are drawn by ticket), as many warps as the wave can fill
        const int grid = (int)std::min<uint32_t>(wb, (uint32_t)num_sms);
        int warps = (int)((wb + (uint32_t)grid - 1) / (uint32_t)grid);
        warps = std::min(std::max(warps, 1), P.geo.max_warps);
        if (a.tune_warps > 0) warps = std::min((int)a.tune_warps, P.geo.max_warps);  // tuning knob
        // More blocks than parser slots: a block that starts late must not be a slow one, or the wave
        // ends on it with the GPU idle.  The parser's cost grows with the bytes to code and with the
        // match pairs it has to price, so blocks are handed out by decreasing (length + pair words).
        // With fewer blocks the same order still pays: the warps of a CTA draw neighbouring tickets, so an
        // SM's streams are alike and run the same parts of the kernel, which its instruction cache likes
        // (profiles/r01_parse_kernel_mixed_w8_w2_ncu.txt).
        pa.order = nullptr;
        if (wb > (uint32_t)num_sms && !a.tune_fifo) {
            std::vector<uint32_t> used(wb), order(wb);
            std::vector<uint64_t> len(wb);
            e = cudaMemcpyAsync(used.data(), w.pair_used, (size_t)wb * sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) return e;
            e = cudaMemcpyAsync(len.data(), w.in_len, (size_t)wb * sizeof(uint64_t), cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) return e;
            e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) return e;
            for (uint32_t i = 0; i < wb; i++) order[i] = i;
            std::stable_sort(order.begin(), order.end(),
                             [&](uint32_t x, uint32_t y) { return len[x] + used[x] > len[y] + used[y]; });
            e = cudaMemcpyAsync(order_dev, order.data(), (size_t)wb * sizeof(uint32_t), cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) return e;
            e = cudaStreamSynchronize(st);  // `order` is pageable host memory that dies with this scope
            if (e != cudaSuccess) return e;
            pa.order = order_dev;
        }
        e = launch_parse(pa, grid, warps, st);
        if (e != cudaSuccess) return e;
        *nl += 1;
        if (P.timing) {
            cudaEventRecord(tev[2], st);
            cudaEventSynchronize(tev[2]);
            float t_mf = 0, t_parse = 0, t_link = 0, t_tree = 0, t_long = 0;
            cudaEventElapsedTime(&t_mf, tev[0], tev[1]);
            cudaEventElapsedTime(&t_parse, tev[1], tev[2]);
            if (a.max_in_len) {
                cudaEventElapsedTime(&t_link, mev[0], mev[1]);
                cudaEventElapsedTime(&t_tree, mev[1], mev[2]);
                cudaEventElapsedTime(&t_long, mev[2], mev[3]);
            }
            fprintf(stderr, "lzb_enc wave: %u blocks (max %llu B), %d warps x %d CTAs, match finder %.1f ms (link %.1f tree %.1f long %.1f), parse %.1f ms\n",
                    wb, (unsigned long long)a.max_in_len, warps, grid, t_mf, t_link, t_tree, t_long, t_parse);
            for (auto& x : tev) cudaEventDestroy(x);
            for (auto& x : mev) cudaEventDestroy(x);
        }
        done += wb;
        if (done < count) {  // the next wave reuses the scratch
            e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) return e;
        }
    }
    return cudaSuccess;
}

}  // namespace

void EncScratch::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

cudaError_t run_encode(const EncodeArgs& a, EncScratch& scratch, int num_sms, cudaStream_t st, int* launches, MfTrace* mf_only) {
    int nl = 0;
    if (launches) *launches = 0;
    if (a.n == 0) return cudaSuccess;
    if (a.max_in_len > kEncMaxBlock) return cudaErrorNotSupported;
    cudaError_t e = upload_mf_tables();
    if (e != cudaSuccess) return e;

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
    P.timing = a.tune_timing;

    size_t free_b = 0, total_b = 0;
    e = cudaMemGetInfo(&free_b, &total_b);
    if (e != cudaSuccess) return e;
    // the scratch may take up to 7/8 of what is free (plus what this handle already holds)
    const size_t budget = std::max<size_t>((free_b + scratch.cap) / 8 * 7, size_t(256) << 20);

    uint32_t pair_mul = 6;  // pair slots per input byte (text needs ~4.4); doubled when a wave overflows
    if (a.tune_pair_mul > 0) pair_mul = (uint32_t)a.tune_pair_mul;  // test knob for the retry path

    e = run_waves(a, P, scratch, num_sms, st, budget, 0, a.n, pair_mul, &nl, mf_only);
    if (launches) *launches = nl;
    return e;
}

}  // namespace lzb
