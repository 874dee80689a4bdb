// lzb_api_enc.cu -- encoder half of the C ABI (include/lzma_b200.h).
#include "../../include/lzma_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "lzb_encode.cuh"
#include "lzb_host.h"

using namespace lzbhost;

struct lzb_enc {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    // Encoder.java:26-27,151-158,172 -- class defaults
    int32_t dict_size = 1 << 22;
    int32_t fb = 32;
    int32_t mf = 1;
    int32_t lc = 3, lp = 0, pb = 2;
    int32_t eos = 0;
    lzb::EncScratch scratch;
    DevBuf d_in, d_out, d_meta;
    PinBuf h_meta;
    // developer / test hooks (DESIGN.md "test hooks"), read ONCE when the handle is created
    int32_t tune_warps = 0, tune_pair_mul = 0, tune_lit = -1, tune_group = 0, tune_inflight = 0;
    int64_t tune_pool = 0;
    bool tune_fifo = false, tune_timing = false, tune_blocked = false;
    lzb_progress_fn progress_fn = nullptr;  // ICodeProgress of the next Code calls
    void* progress_user = nullptr;
};

extern "C" {

lzb_enc* lzb_enc_create(int device) {
    lzb_enc* e = new (std::nothrow) lzb_enc();
    if (!e) {
        fail(LZB_E_NOMEM, "out of host memory");
        return nullptr;
    }
    e->device = device;
    if (open_device(device, &e->stream, &e->num_sms) != LZB_OK) {
        delete e;
        return nullptr;
    }
    if (const char* v = getenv("LZB_ENC_WARPS")) e->tune_warps = atoi(v) > 0 ? atoi(v) : 0;
    if (const char* v = getenv("LZB_PAIR_MUL")) e->tune_pair_mul = atoi(v) > 0 ? atoi(v) : 0;
    if (const char* v = getenv("LZB_ENC_LIT")) e->tune_lit = v[0] == 's' ? 0 : v[0] == 'g' ? 1 : -1;  // smem / global
    if (const char* v = getenv("LZB_ENC_GROUP")) e->tune_group = atoi(v) > 0 ? atoi(v) : 0;
    if (const char* v = getenv("LZB_ENC_POOL_MB")) e->tune_pool = atoll(v) > 0 ? atoll(v) << 20 : 0;
    e->tune_fifo = getenv("LZB_ENC_FIFO") != nullptr;
    e->tune_blocked = getenv("LZB_ENC_BLOCKED") != nullptr;
    if (const char* v = getenv("LZB_ENC_INFLIGHT")) e->tune_inflight = atoi(v) > 0 ? atoi(v) : 0;
    e->tune_timing = getenv("LZB_ENC_TIMING") != nullptr;
    return e;
}

void lzb_enc_destroy(lzb_enc* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    e->scratch.release();
    e->d_in.release();
    e->d_out.release();
    e->d_meta.release();
    e->h_meta.release();
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

int lzb_enc_set_dictionary_size(lzb_enc* e, int32_t v) {  // Encoder.java:1135-1146
    if (!e) return fail(LZB_E_ARG, "null handle");
    if (v < 1 || v > (1 << 29)) return LZB_FALSE;
    e->dict_size = v;
    return LZB_OK;
}
int lzb_enc_set_num_fast_bytes(lzb_enc* e, int32_t v) {  // :1148-1154
    if (!e) return fail(LZB_E_ARG, "null handle");
    if (v < 5 || v > lzb::kMatchMaxLen) return LZB_FALSE;
    e->fb = v;
    return LZB_OK;
}
int lzb_enc_set_match_finder(lzb_enc* e, int32_t v) {  // :1156-1167
    if (!e) return fail(LZB_E_ARG, "null handle");
    if (v < 0 || v > 2) return LZB_FALSE;
    e->mf = v;
    return LZB_OK;
}
int lzb_enc_set_lc_lp_pb(lzb_enc* e, int32_t lc, int32_t lp, int32_t pb) {  // :1169-1180
    if (!e) return fail(LZB_E_ARG, "null handle");
    if (lp < 0 || lp > 4 || lc < 0 || lc > 8 || pb < 0 || pb > 4) return LZB_FALSE;
    e->lc = lc;
    e->lp = lp;
    e->pb = pb;
    return LZB_OK;
}
int lzb_enc_set_end_marker_mode(lzb_enc* e, int32_t v) {  // :1182-1184
    if (!e) return fail(LZB_E_ARG, "null handle");
    e->eos = v ? 1 : 0;
    return LZB_OK;
}
int lzb_enc_set_algorithm(int32_t) { return LZB_OK; }  // :1127-1133, a no-op in the reference too

int lzb_enc_write_coder_properties(const lzb_enc* e, uint8_t out[5]) {  // :1079-1085
    if (!e || !out) return fail(LZB_E_ARG, "null argument");
    out[0] = (uint8_t)((e->pb * 5 + e->lp) * 9 + e->lc);
    for (int i = 0; i < 4; i++) out[1 + i] = (uint8_t)((uint32_t)e->dict_size >> (8 * i));
    return LZB_OK;
}

uint64_t lzb_enc_bound(uint64_t in_len) { return in_len + in_len / 3 + 128; }

int lzb_enc_set_progress(lzb_enc* e, lzb_progress_fn fn, void* user) {  // the ICodeProgress argument of Encoder.Code (:1064)
    if (!e) return fail(LZB_E_ARG, "null handle");
    e->progress_fn = fn;
    e->progress_user = user;
    return LZB_OK;
}

int lzb_enc_code_batch_device(lzb_enc* e, const uint8_t* d_in, const uint64_t* d_in_off, const uint64_t* d_in_len,
                              uint32_t n, uint64_t max_in_len, uint8_t* d_out, const uint64_t* d_out_off,
                              const uint64_t* d_out_cap, uint64_t* d_out_len, int32_t with_header13,
                              void* cuda_stream) {
    if (!e) return fail(LZB_E_ARG, "null handle");
    if (n == 0) return LZB_OK;
    if (!d_in || !d_in_off || !d_in_len || !d_out || !d_out_off || !d_out_cap || !d_out_len)
        return fail(LZB_E_ARG, "null argument");
    // positions are 32-bit and BinTree.Normalize (BinTree.java:358-375, at 2^30 - 1 positions) is not built
    if (max_in_len > lzb::kEncMaxBlockBytes)
        return fail(LZB_E_UNSUPPORTED, "stream of %llu bytes: one stream may have at most %llu bytes",
                    (unsigned long long)max_in_len, (unsigned long long)lzb::kEncMaxBlockBytes);
    CUDA_TRY(cudaSetDevice(e->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : e->stream;
    lzb::EncodeArgs a;
    a.in = d_in;
    a.in_off = d_in_off;
    a.in_len = d_in_len;
    a.out = d_out;
    a.out_off = d_out_off;
    a.out_cap = d_out_cap;
    a.out_len = d_out_len;
    a.n = n;
    a.max_in_len = max_in_len;
    a.dict_size = e->dict_size;
    a.fb = e->fb;
    a.bt4 = e->mf != 0;
    a.lc = e->lc;
    a.lp = e->lp;
    a.pb = e->pb;
    a.eos = e->eos;
    a.with_header = with_header13 != 0;
    a.progress_fn = e->progress_fn;
    a.progress_user = e->progress_user;
    a.tune_warps = e->tune_warps;
    a.tune_pair_mul = e->tune_pair_mul;
    a.tune_lit = e->tune_lit;
    a.tune_group = e->tune_group;
    a.tune_pool = e->tune_pool;
    a.tune_fifo = e->tune_fifo;
    a.tune_blocked = e->tune_blocked;
    a.tune_inflight = e->tune_inflight;
    a.tune_timing = e->tune_timing;
    int launches = 0;
    cudaError_t err = lzb::run_encode(a, e->scratch, e->num_sms, st, &launches);
    add_launches(launches);
    if (err == cudaErrorInvalidValue) return fail(LZB_E_ARG, "encode: a block is longer than max_in_len = %llu", (unsigned long long)max_in_len);
    if (err != cudaSuccess)
        return fail(err == cudaErrorMemoryAllocation ? LZB_E_NOMEM : LZB_E_CUDA, "encode: %s", cudaGetErrorString(err));
    return LZB_OK;
}

int lzb_enc_code_batch(lzb_enc* e, const uint8_t* in, const uint64_t* in_off, const uint64_t* in_len, uint32_t n,
                       uint8_t* out, const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len,
                       int32_t with_header13) {
    if (!e) return fail(LZB_E_ARG, "null handle");
    if (n == 0) return LZB_OK;
    if (!in_off || !in_len || !out || !out_off || !out_cap || !out_len) return fail(LZB_E_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    const Span si = span_of(in_off, in_len, n);
    const Span so = span_of(out_off, out_cap, n);
    if (si.sum && !in) return fail(LZB_E_ARG, "null input");
    uint64_t max_in = 0;
    for (uint32_t i = 0; i < n; i++)
        if (in_len[i] > max_in) max_in = in_len[i];
    const size_t in_bytes = si.hi - si.lo, out_bytes = so.hi - so.lo;
    CUDA_TRY(e->d_in.reserve(in_bytes + 16));
    CUDA_TRY(e->d_out.reserve(out_bytes + 16));
    const size_t meta_bytes = (size_t)n * 5 * sizeof(uint64_t);
    CUDA_TRY(e->d_meta.reserve(meta_bytes));
    CUDA_TRY(e->h_meta.reserve(meta_bytes));
    uint64_t* hm = (uint64_t*)e->h_meta.p;
    for (uint32_t i = 0; i < n; i++) {
        hm[i] = in_off[i] - si.lo;
        hm[n + i] = in_len[i];
        hm[2 * (size_t)n + i] = out_off[i] - so.lo;
        hm[3 * (size_t)n + i] = out_cap[i];
    }
    uint64_t* dm = (uint64_t*)e->d_meta.p;
    CUDA_TRY(cudaMemcpyAsync(dm, hm, (size_t)n * 4 * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    if (in_bytes) CUDA_TRY(cudaMemcpyAsync(e->d_in.p, in + si.lo, in_bytes, cudaMemcpyHostToDevice, st));
    int rc = lzb_enc_code_batch_device(e, (const uint8_t*)e->d_in.p, dm, dm + n, n, max_in, (uint8_t*)e->d_out.p,
                                       dm + 2 * (size_t)n, dm + 3 * (size_t)n, dm + 4 * (size_t)n, with_header13, st);
    if (rc != LZB_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(hm + 4 * (size_t)n, dm + 4 * (size_t)n, (size_t)n * sizeof(uint64_t),
                             cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    // compressed sizes are known now: fetch exactly the bytes that were produced
    for (uint32_t i = 0; i < n; i++) {
        const uint64_t len = hm[4 * (size_t)n + i];
        out_len[i] = len;
        if (len == ~0ull) continue;
        if (len)
            CUDA_TRY(cudaMemcpyAsync(out + out_off[i], (const uint8_t*)e->d_out.p + (out_off[i] - so.lo), len,
                                     cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    for (uint32_t i = 0; i < n; i++)
        if (out_len[i] == ~0ull) {
            out_len[i] = 0;
            return fail(LZB_E_CAPACITY, "block %u: output capacity %llu too small", i, (unsigned long long)out_cap[i]);
        }
    return LZB_OK;
}

int lzb_enc_trace_matches(lzb_enc* e, const uint8_t* in, uint64_t in_len, uint32_t* counts, uint32_t* pairs,
                          uint64_t pairs_cap, uint64_t* pairs_used) {
    if (!e) return fail(LZB_E_ARG, "null handle");
    if (pairs_used) *pairs_used = 0;
    if (in_len == 0) return LZB_OK;
    if (!in || !counts || (!pairs && pairs_cap)) return fail(LZB_E_ARG, "null argument");
    if (in_len > lzb::kEncMaxBlockBytes) return fail(LZB_E_UNSUPPORTED, "block of %llu bytes is too large", (unsigned long long)in_len);
    CUDA_TRY(cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    CUDA_TRY(e->d_in.reserve(in_len + 16));
    CUDA_TRY(e->d_meta.reserve(2 * sizeof(uint64_t)));
    const uint64_t meta[2] = {0, in_len};
    CUDA_TRY(cudaMemcpyAsync(e->d_meta.p, meta, sizeof meta, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(e->d_in.p, in, in_len, cudaMemcpyHostToDevice, st));
    lzb::EncodeArgs a = {};
    a.in = (const uint8_t*)e->d_in.p;
    a.in_off = (const uint64_t*)e->d_meta.p;
    a.in_len = (const uint64_t*)e->d_meta.p + 1;
    a.n = 1;
    a.max_in_len = in_len;
    a.dict_size = e->dict_size;
    a.fb = e->fb;
    a.bt4 = e->mf != 0;
    a.lc = e->lc;
    a.lp = e->lp;
    a.pb = e->pb;
    int launches = 0;
    lzb::MfTrace tr;
    cudaError_t err = lzb::run_encode(a, e->scratch, e->num_sms, st, &launches, &tr);
    add_launches(launches);
    if (err != cudaSuccess) return fail(LZB_E_CUDA, "match finder: %s", cudaGetErrorString(err));
    std::vector<uint32_t> idx, words;
    std::vector<uint16_t> words2;
    try {
        idx.resize((size_t)in_len + 1);
        words.resize(tr.pair_words ? tr.pair_words : 1);
        words2.resize(tr.pair_words ? tr.pair_words : 1);
    } catch (...) {
        return fail(LZB_E_NOMEM, "out of host memory");
    }
    CUDA_TRY(cudaMemcpy(idx.data(), tr.idx, idx.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (tr.pair_words) {
        CUDA_TRY(cudaMemcpy(words.data(), tr.pairs, (size_t)tr.pair_words * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(words2.data(), tr.pairs2, (size_t)tr.pair_words * sizeof(uint16_t), cudaMemcpyDeviceToHost));
    }
    uint64_t used = 0;
    for (uint64_t p = 0; p < in_len; p++) {
        const uint32_t off = idx[p + 1];
        const uint32_t cnt = off == 0xFFFFFFFFu ? 0 : words[off];
        counts[p] = cnt;
        for (uint32_t k = 0; k < cnt; k++, used++) {
            if (used < pairs_cap) {
                const uint32_t w = words[off + 1 + k];
                pairs[2 * used] = lzb::pair_len(w, words2[off + 1 + k]);
                pairs[2 * used + 1] = lzb::pair_dist(w);
            }
        }
    }
    if (pairs_used) *pairs_used = used;
    return used <= pairs_cap ? LZB_OK : fail(LZB_E_CAPACITY, "pairs_cap %llu < %llu pairs", (unsigned long long)pairs_cap, (unsigned long long)used);
}

int lzb_enc_code(lzb_enc* e, const uint8_t* in, uint64_t in_len, uint8_t* out, uint64_t out_cap, uint64_t* out_len) {
    if (!e) return fail(LZB_E_ARG, "null handle");
    if ((!in && in_len) || !out) return fail(LZB_E_ARG, "null buffer");
    uint64_t off = 0, ooff = 0, olen = 0;
    int rc = lzb_enc_code_batch(e, in, &off, &in_len, 1, out, &ooff, &out_cap, &olen, 0);
    if (out_len) *out_len = olen;
    return rc;
}

}  // extern "C"
