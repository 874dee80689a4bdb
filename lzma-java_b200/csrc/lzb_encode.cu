// lzb_encode.cu -- encoder pipeline orchestration: scratch carving, waves, launches.
//
// Replaces Encoder.Code (LZMA/Encoder.java:1064-1077) for a batch of
// independent blocks.  Per wave of blocks:  memset(heads, next, counters) ->
// lzb_mf_link_kernel -> lzb_mf_tree_kernel -> lzb_mf_long_kernel -> lzb_parse_kernel.
// A batch that fits the resident parser slots is one wave (run_waves); a larger one is cut into
// groups that flow through several lanes (run_pipelined) so that parser slots never sit idle
// behind a wave's slowest block and the match finder overlaps the parsers.
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

// carve all scratch for a wave of `wb` blocks; with base == nullptr only sizes are computed
size_t carve(uint8_t* base, uint32_t wb, uint32_t np, uint32_t hash_stride, uint32_t pair_cap, size_t slots, size_t lit_slots,
             MfWave* w, ParseArgs* pa, uint32_t** ctrl) {
    Carver c{base};
    uint32_t* ctl = c.take<uint32_t>(64);                  // [0] parser ticket, [1] pair overflow
    uint32_t* pair_used = c.take<uint32_t>(wb);
    uint32_t* heads = c.take<uint32_t>((size_t)wb * hash_stride);
    uint32_t* next = c.take<uint32_t>((size_t)wb * np);
    const size_t zero_end = c.off;                          // everything above is zeroed per wave
    uint32_t* prev2 = c.take<uint32_t>((size_t)wb * np);
    uint32_t* prev3 = c.take<uint32_t>((size_t)wb * np);
    uint32_t* son = c.take<uint32_t>((size_t)wb * 2 * np);
    uint32_t* idx = c.take<uint32_t>((size_t)wb * np);
    uint32_t* pairs = c.take<uint32_t>((size_t)wb * pair_cap + 64);   // + slack: the parser prefetches 32 slots blindly
    uint16_t* pairs2 = c.take<uint16_t>((size_t)wb * pair_cap + 64);
    uint4* long_items = c.take<uint4>((size_t)wb * (np / kLongChain + 1));  // a block has at most n / kLongChain long buckets
    void* opt = c.take<uint8_t>(slots * parse_opt_bytes_per_slot());
    uint16_t* lit = c.take<uint16_t>(lit_slots);
    if (w) {
        w->heads = heads;
        w->next = next;
        w->prev2 = prev2;
        w->prev3 = prev3;
        w->son = son;
        w->idx = idx;
        w->pairs = pairs;
        w->pairs2 = pairs2;
        w->pair_used = pair_used;
        w->overflow = ctl + 1;
        w->long_count = ctl + 2;
        w->long_ticket = ctl + 3;
        w->long_items = long_items;
    }
    if (pa) {
        pa->ticket = ctl;
        pa->opt_scratch = opt;
        pa->lit_scratch = lit;
    }
    if (ctrl) *ctrl = reinterpret_cast<uint32_t*>(zero_end);  // smuggles the zeroed prefix length
    return c.off;
}

// ---- what a batch has in common ----------------------------------------------------------
struct Plan {
    uint32_t hash_stride, hash_mask, np;
    int dic_log;
    ParseGeometry geo;
    size_t lit_per_slot;  // literal coder in global memory (0 when it lives in shared memory)
    bool timing;
};

// MfWave + ParseArgs for blocks [first, first + wb) over the scratch set at `base`
void bind_wave(const EncodeArgs& a, const Plan& P, uint8_t* base, uint32_t first, uint32_t wb, uint32_t pair_cap, size_t slots,
               MfWave* w, ParseArgs* pa, size_t* zero_len) {
    uint32_t* zero_len_ptr = nullptr;
    carve(base, wb, P.np, P.hash_stride, pair_cap, slots, slots * P.lit_per_slot, w, pa, &zero_len_ptr);
    *zero_len = reinterpret_cast<size_t>(zero_len_ptr);
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
    pa->slice_budget = P.geo.slice_budget;
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
        // largest wave that fits the budget
        uint32_t wb = std::min<uint32_t>(count - done, 32768);
        while (wb > 1 && carve(nullptr, wb, P.np, P.hash_stride, pair_cap, slots, slots * P.lit_per_slot, nullptr, nullptr, nullptr) > budget)
            wb = (wb + 1) / 2;
        const size_t need = carve(nullptr, wb, P.np, P.hash_stride, pair_cap, slots, slots * P.lit_per_slot, nullptr, nullptr, nullptr);
        e = grow(scratch, need, st);
        if (e != cudaSuccess) return e;
        MfWave w;
        ParseArgs pa;
        size_t zero_len = 0;
        bind_wave(a, P, (uint8_t*)scratch.p, first + done, wb, pair_cap, slots, &w, &pa, &zero_len);
        // LZB_ENC_TIMING=1 (developer hook): phase times of every wave on stderr
        cudaEvent_t tev[3] = {nullptr, nullptr, nullptr};
        if (P.timing) {
            for (auto& x : tev) cudaEventCreate(&x);
            cudaEventRecord(tev[0], st);
        }
        e = cudaMemsetAsync(scratch.p, 0, zero_len, st);
        if (e != cudaSuccess) return e;
        e = launch_mf(w, (uint32_t)a.max_in_len, num_sms, st);
        if (e != cudaSuccess) return e;
        *nl += a.max_in_len ? 3 : 1;

        // did any block run out of pair slots?  (rare: retry the wave with twice the room)
        uint32_t overflow = 0;
        e = cudaMemcpyAsync(&overflow, w.overflow, sizeof overflow, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return e;
        if (P.timing) cudaEventRecord(tev[1], st);
        if (overflow) {
            if (P.timing)
                for (auto& x : tev) cudaEventDestroy(x);
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
            mf_only->pair_words = used;
            return cudaSuccess;
        }
        pa.mf = w;
        int warps = (int)((wb + (uint32_t)num_sms - 1) / (uint32_t)num_sms);
        warps = std::min(std::max(warps, 1), P.geo.max_warps);
        if (const char* ev = getenv("LZB_ENC_WARPS")) warps = std::min(std::max(atoi(ev), 1), P.geo.max_warps);  // tuning knob
        const int grid = (int)std::min<uint32_t>((wb + (uint32_t)warps - 1) / (uint32_t)warps, (uint32_t)num_sms);
        e = launch_parse(pa, grid, warps, st);
        if (e != cudaSuccess) return e;
        *nl += 1;
        if (P.timing) {
            cudaEventRecord(tev[2], st);
            cudaEventSynchronize(tev[2]);
            float t_mf = 0, t_parse = 0;
            cudaEventElapsedTime(&t_mf, tev[0], tev[1]);
            cudaEventElapsedTime(&t_parse, tev[1], tev[2]);
            fprintf(stderr, "lzb_enc wave: %u blocks (max %llu B), %d warps x %d CTAs, match finder %.1f ms, parse %.1f ms\n", wb,
                    (unsigned long long)a.max_in_len, warps, grid, t_mf, t_parse);
            for (auto& x : tev) cudaEventDestroy(x);
        }
        done += wb;
        if (done < count) {  // the next wave reuses the scratch
            e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) return e;
        }
    }
    return cudaSuccess;
}

// Batches with more blocks than resident parser slots: groups of `group` blocks go round-robin
// over `lanes` lanes, each lane a CUDA stream with its own scratch set running
// match finder -> parse for its groups in order.  The parser runs one warp per CTA here, so the
// hardware block scheduler keeps every SM's parser slots filled from whichever lanes have
// parse CTAs pending, and the match finder of later groups runs in the issue slots the
// latency-bound parsers leave idle.  No host synchronisation until the end: a group that ran out
// of pair slots is skipped by its parse kernel (flag checked on the device) and redone afterwards.
cudaError_t run_pipelined(const EncodeArgs& a, const Plan& P, EncScratch& scratch, int num_sms, cudaStream_t st, size_t budget,
                          uint32_t group, uint32_t lanes, uint32_t pair_mul, int* nl) {
    const uint32_t pair_cap = pair_cap_for(pair_mul, a.max_in_len);
    const uint32_t n_groups = (a.n + group - 1) / group;
    const size_t set_bytes =
        (carve(nullptr, group, P.np, P.hash_stride, pair_cap, group, group * P.lit_per_slot, nullptr, nullptr, nullptr) + 255) & ~size_t(255);
    const size_t head_bytes = ((size_t)n_groups * sizeof(uint32_t) + 255) & ~size_t(255);  // one overflow flag per group
    cudaError_t e = grow(scratch, head_bytes + (size_t)lanes * set_bytes, st);
    if (e != cudaSuccess) return e;
    e = scratch.ensure_lanes(lanes);
    if (e != cudaSuccess) return e;
    uint32_t* d_ovf = reinterpret_cast<uint32_t*>(scratch.p);
    uint8_t* sets = (uint8_t*)scratch.p + head_bytes;

    cudaEvent_t tev[2] = {nullptr, nullptr};
    if (P.timing) {
        for (auto& x : tev) cudaEventCreate(&x);
        cudaEventRecord(tev[0], st);
    }
    e = cudaMemsetAsync(d_ovf, 0, head_bytes, st);
    if (e != cudaSuccess) return e;
    e = cudaEventRecord(scratch.fork, st);
    if (e != cudaSuccess) return e;
    for (uint32_t l = 0; l < lanes; l++) {
        e = cudaStreamWaitEvent(scratch.lanes[l], scratch.fork, 0);
        if (e != cudaSuccess) return e;
    }
    for (uint32_t g = 0; g < n_groups; g++) {
        const uint32_t l = g % lanes, first = g * group, gb = std::min(group, a.n - first);
        cudaStream_t ls = scratch.lanes[l];
        uint8_t* base = sets + (size_t)l * set_bytes;
        MfWave w;
        ParseArgs pa;
        size_t zero_len = 0;
        bind_wave(a, P, base, first, gb, pair_cap, group, &w, &pa, &zero_len);
        w.overflow = d_ovf + g;
        e = cudaMemsetAsync(base, 0, zero_len, ls);
        if (e != cudaSuccess) return e;
        e = launch_mf(w, (uint32_t)a.max_in_len, num_sms, ls);
        if (e != cudaSuccess) return e;
        pa.mf = w;
        e = launch_parse(pa, (int)gb, 1, ls);
        if (e != cudaSuccess) return e;
        *nl += a.max_in_len ? 4 : 2;
    }
    for (uint32_t l = 0; l < lanes; l++) {
        e = cudaEventRecord(scratch.joins[l], scratch.lanes[l]);
        if (e != cudaSuccess) return e;
        e = cudaStreamWaitEvent(st, scratch.joins[l], 0);
        if (e != cudaSuccess) return e;
    }
    std::vector<uint32_t> ovf(n_groups, 0);
    e = cudaMemcpyAsync(ovf.data(), d_ovf, (size_t)n_groups * sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return e;
    if (P.timing) cudaEventRecord(tev[1], st);
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return e;
    if (P.timing) {
        float ms = 0;
        cudaEventElapsedTime(&ms, tev[0], tev[1]);
        fprintf(stderr, "lzb_enc pipeline: %u blocks (max %llu B) in %u groups of %u over %u lanes, %.1f ms\n", a.n,
                (unsigned long long)a.max_in_len, n_groups, group, lanes, ms);
        for (auto& x : tev) cudaEventDestroy(x);
    }
    for (uint32_t g = 0; g < n_groups; g++) {  // rare: redo the groups whose match lists did not fit
        if (!ovf[g]) continue;
        if (pair_mul >= 512) return cudaErrorMemoryAllocation;
        const uint32_t first = g * group, gb = std::min(group, a.n - first);
        e = run_waves(a, P, scratch, num_sms, st, budget, first, gb, pair_mul * 2, nl, nullptr);
        if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

int env_int(const char* name, int fallback) {
    const char* ev = getenv(name);
    return ev ? atoi(ev) : fallback;
}

}  // namespace

void EncScratch::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    for (cudaStream_t s : lanes) cudaStreamDestroy(s);
    for (cudaEvent_t ev : joins) cudaEventDestroy(ev);
    if (fork) cudaEventDestroy(fork);
    lanes.clear();
    joins.clear();
    fork = nullptr;
}

cudaError_t EncScratch::ensure_lanes(uint32_t count) {
    cudaError_t e;
    if (!fork) {
        e = cudaEventCreateWithFlags(&fork, cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    while (lanes.size() < count) {
        cudaStream_t s = nullptr;
        e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
        if (e != cudaSuccess) return e;
        lanes.push_back(s);
        cudaEvent_t ev = nullptr;
        e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
        joins.push_back(ev);
    }
    return cudaSuccess;
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
    P.geo = parse_geometry(a.lc, a.lp, a.pb, a.fb);
    P.lit_per_slot = P.geo.lit_in_smem ? 0 : ((size_t)0x300 << (a.lc + a.lp));
    P.timing = getenv("LZB_ENC_TIMING") != nullptr;

    size_t free_b = 0, total_b = 0;
    e = cudaMemGetInfo(&free_b, &total_b);
    if (e != cudaSuccess) return e;
    // the scratch may take up to 3/4 of what is free (plus what this handle already holds)
    const size_t budget = std::max<size_t>((free_b + scratch.cap) / 4 * 3, size_t(256) << 20);

    uint32_t pair_mul = 6;  // pair slots per input byte (text needs ~4.4); doubled when a wave overflows
    if (const char* ev = getenv("LZB_PAIR_MUL")) pair_mul = (uint32_t)std::max(atoi(ev), 1);  // test knob for the retry path

    // More blocks than resident parser slots: pipeline groups over lanes (run_pipelined).
    // LZB_ENC_PIPE=0/1 forces the choice, LZB_ENC_GROUP / LZB_ENC_LANES set the shape (test knobs).
    const size_t resident = (size_t)num_sms * P.geo.max_warps;
    const int pipe = env_int("LZB_ENC_PIPE", -1);
    if (!mf_only && pipe != 0 && (pipe == 1 || a.n > resident)) {
        const uint32_t lanes = (uint32_t)std::min(std::max(env_int("LZB_ENC_LANES", 6), 1), 16);
        const uint32_t pair_cap = pair_cap_for(pair_mul, a.max_in_len);
        const size_t per_block =
            carve(nullptr, 64, P.np, P.hash_stride, pair_cap, 64, 64 * P.lit_per_slot, nullptr, nullptr, nullptr) / 64 + 1;
        uint32_t group = (uint32_t)std::min<size_t>(2 * (size_t)num_sms, budget / lanes / per_block);
        if (const char* ev = getenv("LZB_ENC_GROUP")) group = (uint32_t)std::max(atoi(ev), 1);
        if (group >= 8 || pipe == 1) {
            group = std::max(group, 1u);
            e = run_pipelined(a, P, scratch, num_sms, st, budget, group, std::min(lanes, (a.n + group - 1) / group), pair_mul, &nl);
            if (launches) *launches = nl;
            return e;
        }
    }
    e = run_waves(a, P, scratch, num_sms, st, budget, 0, a.n, pair_mul, &nl, mf_only);
    if (launches) *launches = nl;
    return e;
}

}  // namespace lzb
