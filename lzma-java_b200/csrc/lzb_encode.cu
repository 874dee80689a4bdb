// lzb_encode.cu -- encoder pipeline orchestration: scratch carving, waves, launches.
//
// Replaces Encoder.Code (LZMA/Encoder.java:1064-1077) for a batch of
// independent blocks.  Per wave of blocks:  memset(heads, next, counters) ->
// lzb_mf_link_kernel -> lzb_mf_tree_kernel -> lzb_parse_kernel.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "lzb_encode.cuh"

namespace lzb {

void EncScratch::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

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

}  // namespace

cudaError_t run_encode(const EncodeArgs& a, EncScratch& scratch, int num_sms, cudaStream_t st, int* launches, MfTrace* mf_only) {
    int nl = 0;
    if (launches) *launches = 0;
    if (a.n == 0) return cudaSuccess;
    if (a.max_in_len > kEncMaxBlock) return cudaErrorNotSupported;
    cudaError_t e = upload_mf_tables();
    if (e != cudaSuccess) return e;

    uint32_t hash_mask = 0;
    const uint32_t hash_stride = hash_stride_for(a.dict_size, a.bt4, &hash_mask);
    const uint32_t np = (uint32_t)a.max_in_len + 1;
    int dic_log = 0;
    while ((uint32_t)a.dict_size > (1u << dic_log)) dic_log++;  // Encoder.java:1141-1144

    // resident parser slots: one CTA per SM, up to kEncMaxWarps streams each
    const ParseGeometry geo = parse_geometry(a.lc, a.lp, a.pb, a.fb);
    const size_t slots = (size_t)num_sms * geo.max_warps;
    const size_t lit_slots = geo.lit_in_smem ? 0 : slots * ((size_t)0x300 << (a.lc + a.lp));

    size_t free_b = 0, total_b = 0;
    e = cudaMemGetInfo(&free_b, &total_b);
    if (e != cudaSuccess) return e;
    // a wave may take up to 3/4 of what is free (plus what this handle already holds)
    const size_t budget = std::max<size_t>((free_b + scratch.cap) / 4 * 3, size_t(256) << 20);

    uint32_t pair_mul = 6;  // pair slots per input byte (text needs ~4.4); doubled when a wave overflows
    if (const char* ev = getenv("LZB_PAIR_MUL")) pair_mul = (uint32_t)std::max(atoi(ev), 1);  // test knob for the retry path
    std::vector<uint64_t> h_off;  // unused; lengths stay on the device

    uint32_t done = 0;
    while (done < a.n) {
        uint32_t pair_cap = (uint32_t)std::min<uint64_t>((uint64_t)pair_mul * a.max_in_len + 4096, 0xFFFFFFF0ull);
        // largest wave that fits the budget
        uint32_t wb = std::min<uint32_t>(a.n - done, 32768);
        while (wb > 1 && carve(nullptr, wb, np, hash_stride, pair_cap, slots, lit_slots, nullptr, nullptr, nullptr) > budget)
            wb = (wb + 1) / 2;
        const size_t need = carve(nullptr, wb, np, hash_stride, pair_cap, slots, lit_slots, nullptr, nullptr, nullptr);
        if (need > scratch.cap) {
            if (scratch.p) {
                cudaStreamSynchronize(st);
                cudaFree(scratch.p);
                scratch.p = nullptr;
                scratch.cap = 0;
            }
            e = cudaMalloc(&scratch.p, need);
            if (e != cudaSuccess) return e;
            scratch.cap = need;
        }
        MfWave w;
        ParseArgs pa;
        uint32_t* zero_len_ptr = nullptr;
        carve((uint8_t*)scratch.p, wb, np, hash_stride, pair_cap, slots, lit_slots, &w, &pa, &zero_len_ptr);
        const size_t zero_len = reinterpret_cast<size_t>(zero_len_ptr);
        w.in = a.in;
        w.in_off = a.in_off + done;
        w.in_len = a.in_len + done;
        w.n_blocks = wb;
        w.np = np;
        w.hash_stride = hash_stride;
        w.pair_cap = pair_cap;
        w.hash_mask = hash_mask;
        w.cyclic_size = (uint32_t)a.dict_size + 1;
        w.fb = a.fb;
        w.cut = 16 + (a.fb >> 1);  // BinTree.java:98
        w.bt4 = a.bt4;
        e = cudaMemsetAsync(scratch.p, 0, zero_len, st);
        if (e != cudaSuccess) return e;
        e = launch_mf(w, (uint32_t)a.max_in_len, num_sms, st);
        if (e != cudaSuccess) return e;
        nl += a.max_in_len ? 3 : 1;

        // did any block run out of pair slots?  (rare: retry the wave with twice the room)
        uint32_t overflow = 0;
        e = cudaMemcpyAsync(&overflow, w.overflow, sizeof overflow, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return e;
        if (overflow) {
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
            if (launches) *launches = nl;
            return cudaSuccess;
        }
        pa.mf = w;
        pa.out = a.out;
        pa.out_off = a.out_off + done;
        pa.out_cap = a.out_cap + done;
        pa.out_len = a.out_len + done;
        pa.dict_size = a.dict_size;
        pa.dist_table_size = dic_log * 2;
        pa.lc = a.lc;
        pa.lp = a.lp;
        pa.pb = a.pb;
        pa.fb = a.fb;
        pa.eos = a.eos;
        pa.with_header = a.with_header;
        int warps = (int)((wb + (uint32_t)num_sms - 1) / (uint32_t)num_sms);
        warps = std::min(std::max(warps, 1), geo.max_warps);
        if (const char* ev = getenv("LZB_ENC_WARPS")) warps = std::min(std::max(atoi(ev), 1), geo.max_warps);  // tuning knob
        pa.slice_bytes = geo.slice_bytes;
        pa.slice_budget = geo.slice_budget;
        int grid = (int)std::min<uint32_t>((wb + (uint32_t)warps - 1) / (uint32_t)warps, (uint32_t)num_sms);
        e = launch_parse(pa, grid, warps, st);
        if (e != cudaSuccess) return e;
        nl += 1;
        done += wb;
        if (done < a.n) {  // the next wave reuses the scratch
            e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) return e;
        }
    }
    if (launches) *launches = nl;
    return cudaSuccess;
}

}  // namespace lzb
