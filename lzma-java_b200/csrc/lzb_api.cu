// lzb_api.cu -- the C ABI of include/lzma_b200.h: handles, property setters
// with the reference's range checks, host<->device staging and kernel
// launches.  No CPU fallback: every code path ends in a kernel launch or an
// error.
#include "../../include/lzma_b200.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <thread>
#include <vector>

#include "lzb_common.cuh"
#include "lzb_kernels.h"
#include "lzb_host.h"

namespace {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
}  // namespace

namespace lzbhost {

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

void add_launches(int n) { g_launches += (uint64_t)n; }

int open_device(int device, cudaStream_t* stream, int* num_sms) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(LZB_E_CUDA, "no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(LZB_E_ARG, "device %d out of range [0,%d)", device, count);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(LZB_E_CUDA, "device %d is sm_%d%d; kernels are built for sm_100a only", device, prop.major, prop.minor);
    *num_sms = prop.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(stream, cudaStreamNonBlocking));
    return LZB_OK;
}

}  // namespace lzbhost

using namespace lzbhost;

// ===========================================================================
// decoder
// ===========================================================================
struct lzb_dec {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    uint8_t props[5] = {0x5D, 0, 0, 0x40, 0};
    bool props_set = false;
    DevBuf ctrl;     // [0] ticket, [1] max spill
    DevBuf lit;      // spilled literal models
    DevBuf d_in, d_out, d_meta;
    PinBuf h_meta, h_progress;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;  // host-buffer batches overlap transfers with the kernels
    cudaStream_t kstream[4] = {nullptr, nullptr, nullptr, nullptr};  // ... whose chunks share the SMs
    // developer / test hooks, read ONCE when the handle is created (DESIGN.md "test hooks")
    int env_mode = -1;    // LZB_DEC_MODE: 0 = all shared, 1 = hybrid
    int env_chunks = 0;   // LZB_DEC_CHUNKS: 1..16
    int env_marks = 0;    // LZB_DEC_MARKS: 1..32
    int env_zero_copy = 1;  // LZB_DEC_ZEROCOPY=0: always stage host input through a device copy
};

namespace {

// the host-side twin of lzb_decode_scan_headers: running max of lc + lp + 1 and pb + 1 over well-formed headers
void header_scan(const uint8_t* stream, uint64_t len, uint32_t* max_lclp1, uint32_t* max_pb1) {
    if (len < LZB_HEADER_SIZE) return;
    const uint32_t v = stream[0];
    const uint32_t lc = v % 9, rem = v / 9, lp = rem % 5, pb = rem / 5;
    if (pb > 4) return;
    if (lc + lp + 1 > *max_lclp1) *max_lclp1 = lc + lp + 1;
    if (pb + 1 > *max_pb1) *max_pb1 = pb + 1;
}

// Where the models of a launch live (lzb_kernels.h, DecMode).  `resident` = streams that will be
// in flight together: the hybrid mode trades a slower literal-after-match for twice the
// streams per SM, which only pays when the all-shared mode could not hold them at once.
int pick_dec_mode(uint32_t max_lclp1, uint32_t max_pb1, uint64_t resident, int num_sms, int forced) {
    if (max_lclp1 > 4) return lzb::kDecGlobal;
    if (max_lclp1 == 0) return lzb::kDecSmem;  // no well-formed header: every stream returns 0 at once
    const int lclp = (int)max_lclp1 - 1;
    const lzb::ModelLayout L = lzb::make_layout(lclp, 0, (int)max_pb1 - 1);
    const bool fits_smem = (size_t)(L.n_fixed + L.n_literal) * 2 <= lzb::dec_mode_model(lzb::kDecSmem);
    const bool fits_hybrid = (size_t)(L.n_fixed + (0x110 << lclp)) * 2 <= lzb::dec_mode_model(lzb::kDecHybrid);
    if (!fits_smem && !fits_hybrid) return lzb::kDecGlobal;  // lc + lp = 3 with pb = 4
    if (!fits_smem) return lzb::kDecHybrid;
    if (!fits_hybrid) return lzb::kDecSmem;
    if (forced == 0) return lzb::kDecSmem;  // test hook
    if (forced == 1) return lzb::kDecHybrid;
    return resident > (uint64_t)num_sms * lzb::kDecMaxWarps ? lzb::kDecHybrid : lzb::kDecSmem;
}

// enqueue the decode kernel for n streams (no header scan, no synchronisation)
int dec_enqueue(lzb_dec* d, const uint8_t* d_in, const uint64_t* d_in_off, const uint64_t* d_in_len, uint32_t n,
                uint8_t* d_out, const uint64_t* d_out_off, const uint64_t* d_out_cap, uint64_t* d_out_len,
                int32_t* d_status, uint32_t max_lclp1, uint32_t max_pb1, int mode, uint32_t* ticket, cudaStream_t st, uint32_t region = 0,
                uint32_t n_regions = 1, uint32_t* progress = nullptr, uint32_t* progress_dev = nullptr, uint32_t marks = 0,
                const uint32_t* mark_at = nullptr) {
    lzb::DecodeArgs a;
    a.in = d_in;
    a.in_off = d_in_off;
    a.in_len = d_in_len;
    a.out = d_out;
    a.out_off = d_out_off;
    a.out_cap = d_out_cap;
    a.out_len = d_out_len;
    a.status = d_status;
    a.n = n;
    a.ticket = ticket;
    a.lit_scratch = nullptr;
    a.lit_stride = 0;
    a.progress = progress;
    a.progress_dev = progress_dev;
    a.marks = marks;
    for (uint32_t m = 0; m < marks && m < lzb::kDecMaxMarks; m++) a.mark_at[m] = mark_at[m];
    if (mode != lzb::kDecSmem && max_lclp1) {
        a.lit_stride = (size_t)(mode == lzb::kDecHybrid ? 0x200 : 0x300) << (max_lclp1 - 1);
        const size_t slots = (size_t)d->num_sms * lzb::dec_mode_warps(mode);
        // launches that may run side by side (one per `kstream`) get a region each
        CUDA_TRY(d->lit.reserve(n_regions * slots * a.lit_stride * sizeof(uint16_t)));
        a.lit_scratch = (uint16_t*)d->lit.p + region * slots * a.lit_stride;
    }
    CUDA_TRY(lzb::launch_decode(a, mode, d->num_sms, st, nullptr, nullptr));
    add_launches(1);
    return LZB_OK;
}

}  // namespace

extern "C" {

const char* lzb_version(void) { return "lzma_b200 0.1 sm_100a"; }

int lzb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char* lzb_last_error(void) { return g_err; }
uint64_t lzb_kernel_launches(void) { return g_launches.load(); }

void* lzb_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        fail(LZB_E_NOMEM, "cudaHostAlloc(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}
void lzb_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

lzb_dec* lzb_dec_create(int device) {
    lzb_dec* d = new (std::nothrow) lzb_dec();
    if (!d) {
        fail(LZB_E_NOMEM, "out of host memory");
        return nullptr;
    }
    d->device = device;
    if (open_device(device, &d->stream, &d->num_sms) != LZB_OK) {
        delete d;
        return nullptr;
    }
    if (const char* e = getenv("LZB_DEC_MODE")) d->env_mode = e[0] == '0' ? 0 : e[0] == '1' ? 1 : -1;
    if (const char* e = getenv("LZB_DEC_CHUNKS")) d->env_chunks = atoi(e);
    if (const char* e = getenv("LZB_DEC_MARKS")) d->env_marks = atoi(e);
    if (const char* e = getenv("LZB_DEC_ZEROCOPY")) d->env_zero_copy = atoi(e) != 0;
    return d;
}

void lzb_dec_destroy(lzb_dec* d) {
    if (!d) return;
    cudaSetDevice(d->device);
    d->ctrl.release();
    d->lit.release();
    d->d_in.release();
    d->d_out.release();
    d->d_meta.release();
    d->h_meta.release();
    d->h_progress.release();
    if (d->stream) cudaStreamDestroy(d->stream);
    if (d->copy_in) cudaStreamDestroy(d->copy_in);
    if (d->copy_out) cudaStreamDestroy(d->copy_out);
    for (cudaStream_t k : d->kstream)
        if (k) cudaStreamDestroy(k);
    delete d;
}

int lzb_dec_set_decoder_properties(lzb_dec* d, const uint8_t* props, uint32_t n_props) {
    if (!d || !props) return fail(LZB_E_ARG, "null argument");
    if (n_props < 5) return LZB_FALSE;  // Decoder.java:304-306
    const int val = props[0];
    const int pb = (val / 9) / 5;
    int32_t dict = 0;
    for (int i = 0; i < 4; i++) dict += (int32_t)((uint32_t)props[1 + i] << (8 * i));
    if (pb > 4) return LZB_FALSE;    // SetLcLpPb, Decoder.java:172-175
    if (dict < 0) return LZB_FALSE;  // SetDictionarySize, Decoder.java:160-163
    memcpy(d->props, props, 5);
    d->props_set = true;
    return LZB_OK;
}

int lzb_dec_code_batch_device(lzb_dec* d, const uint8_t* d_in, const uint64_t* d_in_off, const uint64_t* d_in_len,
                              uint32_t n, uint8_t* d_out, const uint64_t* d_out_off, const uint64_t* d_out_cap,
                              uint64_t* d_out_len, int32_t* d_status, void* cuda_stream) {
    if (!d) return fail(LZB_E_ARG, "null handle");
    if (n == 0) return LZB_OK;
    if (!d_in || !d_in_off || !d_in_len || !d_out || !d_out_off || !d_out_cap || !d_out_len || !d_status)
        return fail(LZB_E_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(d->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : d->stream;
    CUDA_TRY(d->ctrl.reserve(64));
    CUDA_TRY(cudaMemsetAsync(d->ctrl.p, 0, 64, st));
    uint32_t* ctrl = (uint32_t*)d->ctrl.p;

    // the largest lc + lp among the headers decides where the literal tables can live
    CUDA_TRY(lzb::launch_decode_scan(d_in, d_in_off, d_in_len, n, ctrl + 1, st));
    add_launches(1);
    uint32_t scan[2] = {0, 0};
    CUDA_TRY(cudaMemcpyAsync(scan, ctrl + 1, sizeof scan, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));

    return dec_enqueue(d, d_in, d_in_off, d_in_len, n, d_out, d_out_off, d_out_cap, d_out_len, d_status, scan[0], scan[1],
                       pick_dec_mode(scan[0], scan[1], n, d->num_sms, d->env_mode), ctrl, st);
}

int lzb_dec_code_batch(lzb_dec* d, const uint8_t* in, const uint64_t* in_off, const uint64_t* in_len, uint32_t n,
                       uint8_t* out, const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len,
                       int32_t* status) {
    if (!d) return fail(LZB_E_ARG, "null handle");
    if (n == 0) return LZB_OK;
    if (!in || !in_off || !in_len || !out || !out_off || !out_cap || !out_len || !status)
        return fail(LZB_E_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(d->device));
    cudaStream_t st = d->stream;
    if (!d->copy_in) CUDA_TRY(cudaStreamCreateWithFlags(&d->copy_in, cudaStreamNonBlocking));
    if (!d->copy_out) CUDA_TRY(cudaStreamCreateWithFlags(&d->copy_out, cudaStreamNonBlocking));

    // Device images keep the caller's layout (offsets relative to the smallest covering span).
    const Span si = span_of(in_off, in_len, n);
    const Span so = span_of(out_off, out_cap, n);
    const size_t in_bytes = si.hi - si.lo, out_bytes = so.hi - so.lo;
    // Input in pinned (or registered) host memory is not copied at all: the decoder's input ring is filled by
    // TMA bulk copies that read the mapped host pages directly, half a ring ahead of the range decoder, so the
    // kernels start at once instead of after a 300 MB host-to-device copy (12 ms of a 100 ms step on C2).
    const uint8_t* in_mapped = nullptr;
    if (d->env_zero_copy && in_bytes) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, in + si.lo) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
            in_mapped = (const uint8_t*)attr.devicePointer;
        else
            cudaGetLastError();  // pageable memory: not an error, stage it
    }
    if (!in_mapped) CUDA_TRY(d->d_in.reserve(in_bytes + 16));
    CUDA_TRY(d->d_out.reserve(out_bytes + 16));
    // meta: in_off in_len out_off out_cap out_len (u64 x n each) + status (i32 x n)
    const size_t meta_bytes = (size_t)n * (5 * sizeof(uint64_t) + sizeof(int32_t));
    CUDA_TRY(d->d_meta.reserve(meta_bytes));
    CUDA_TRY(d->h_meta.reserve(meta_bytes));
    uint64_t* hm = (uint64_t*)d->h_meta.p;
    uint32_t max_lclp1 = 0, max_pb1 = 0;
    for (uint32_t i = 0; i < n; i++) {
        hm[i] = in_off[i] - si.lo;
        hm[n + i] = in_len[i];
        hm[2 * (size_t)n + i] = out_off[i] - so.lo;
        hm[3 * (size_t)n + i] = out_cap[i];
        header_scan(in + in_off[i], in_len[i], &max_lclp1, &max_pb1);
    }
    uint64_t* dm = (uint64_t*)d->d_meta.p;
    const uint8_t* d_in = in_mapped ? in_mapped : (const uint8_t*)d->d_in.p;
    uint8_t* d_out = (uint8_t*)d->d_out.p;

    // Two ways to keep the PCIe bus busy while the kernels run.
    // (a) Row-shaped output (every stream the same capacity, >= 64 KiB, at one pitch: a block
    //     codec's layout): ONE launch for the whole batch, and the output is read back
    //     progressively behind the decoders (see below).  Cutting such a batch into chunks was
    //     measured to lose: co-resident chunks share the issue slots, so a chunk that starts late
    //     finishes late and runs its tail on a mostly idle GPU (4096 x 256 KiB: 1 chunk 10.8 GB/s,
    //     2 chunks 10.5, 4 chunks 10.0).
    // (b) Any other layout: chunks of about 7 streams per SM, each with its own input copy, kernel
    //     launch and output copy; the kernels go round-robin over up to four streams so that four
    //     chunks share the SMs (4 x 7 warps = the residency of one big launch).
    const int mode = pick_dec_mode(max_lclp1, max_pb1, n, d->num_sms, d->env_mode);
    const uint32_t n_k = mode == lzb::kDecGlobal ? 1u : 4u;  // kDecGlobal: scratch can be GBs, one launch at a time
    for (uint32_t k = 0; k < n_k; k++)
        if (!d->kstream[k]) CUDA_TRY(cudaStreamCreateWithFlags(&d->kstream[k], cudaStreamNonBlocking));
    const uint32_t per_chunk = (uint32_t)d->num_sms * 7;
    uint32_t n_chunks = (n + per_chunk - 1) / per_chunk;
    if (n_chunks > 16) n_chunks = 16;
    {
        bool whole = out_cap[0] >= (64u << 10);
        const uint64_t pitch = n > 1 ? out_off[1] - out_off[0] : out_cap[0];
        for (uint32_t i = 1; i < n && whole; i++) whole = out_cap[i] == out_cap[0] && out_off[i] == out_off[0] + (uint64_t)i * pitch;
        if (whole) n_chunks = 1;
    }
    if (d->env_chunks >= 1 && d->env_chunks <= 16 && (uint32_t)d->env_chunks <= n) n_chunks = (uint32_t)d->env_chunks;  // test hook
    std::vector<uint32_t> first(n_chunks + 1);
    std::vector<Span> cin(n_chunks), cout(n_chunks);
    for (uint32_t c = 0; c <= n_chunks; c++) first[c] = (uint32_t)((uint64_t)n * c / n_chunks);
    bool ordered = true;
    for (uint32_t c = 0; c < n_chunks; c++) {
        cin[c] = span_of(in_off + first[c], in_len + first[c], first[c + 1] - first[c]);
        cout[c] = span_of(out_off + first[c], out_cap + first[c], first[c + 1] - first[c]);
        if (c && (cin[c].lo < cin[c - 1].hi || cout[c].lo < cout[c - 1].hi)) ordered = false;
    }
    if (!ordered) {  // interleaved layout: one chunk, copies around the kernel
        n_chunks = 1;
        first[1] = n;
        cin[0] = si;
        cout[0] = so;
    }
    // Progressive read-back.  Streams of a chunk advance together, so their output can leave
    // the device while they are still being decoded: when a chunk's outputs are rows of one
    // pitch and one capacity (the layout of a block codec), the kernel counts, per mark, the
    // streams whose output below that mark is complete (DecodeArgs::progress, in pinned host
    // memory), and this thread issues one strided copy per (chunk, mark) as the counts fill up.
    // Other layouts are copied chunk by chunk after their kernel.
    // Every copy but the last one hides behind the decoders, so the marks are not evenly spaced: 1/4, 1/2, 3/4,
    // 7/8, 15/16, 31/32 of a row and its end -- few wide copies first (a strided copy of thousands of rows has a
    // cost of its own, and so has a mark inside the kernel), and only 1/32 of the output left to fetch once the
    // kernel has ended.  LZB_DEC_MARKS = k asks for k even marks instead (test hook).
    constexpr uint32_t kMaxMarks = lzb::kDecMaxMarks;
    const uint32_t kMarks = kMaxMarks;  // stride of the progress counters per chunk
    const uint32_t even_marks = d->env_marks >= 1 && d->env_marks <= (int)kMaxMarks ? (uint32_t)d->env_marks : 0;
    struct Rows {
        bool on = false;
        uint64_t pitch = 0, cap = 0;
        uint32_t marks = 0, issued = 0;
        uint32_t at[lzb::kDecMaxMarks] = {};  // at[m] = end column of mark m (a multiple of 512 but for the last = cap)
    };
    std::vector<Rows> rows(n_chunks);
    std::vector<char> span_copy(n_chunks, 1);
    uint32_t pending_marks = 0;
    for (uint32_t c = 0; c < n_chunks && ordered; c++) {
        const uint32_t s0 = first[c], cnt = first[c + 1] - s0;
        if (cnt == 0) continue;
        Rows r;
        r.cap = out_cap[s0];
        r.pitch = cnt > 1 ? out_off[s0 + 1] - out_off[s0] : r.cap;
        r.on = r.cap >= (64u << 10) && r.cap < (1ull << 31) && r.pitch >= r.cap;
        for (uint32_t i = 1; i < cnt && r.on; i++)
            r.on = out_cap[s0 + i] == r.cap && out_off[s0 + i] == out_off[s0] + (uint64_t)i * r.pitch;
        if (!r.on) continue;
        if (even_marks) {
            const uint32_t step = (uint32_t)(((r.cap + even_marks - 1) / even_marks + 511) & ~(uint64_t)511);
            for (uint64_t e = step; e < r.cap; e += step) r.at[r.marks++] = (uint32_t)e;
        } else {
            static const uint32_t num[6] = {8, 16, 24, 28, 30, 31};  // 32nds of a row
            for (uint32_t k = 0; k < 6; k++) {
                const uint32_t e = (uint32_t)((r.cap * num[k] / 32) & ~(uint64_t)511);
                if (e > (r.marks ? r.at[r.marks - 1] : 0u) && e < r.cap) r.at[r.marks++] = e;
            }
        }
        r.at[r.marks++] = (uint32_t)r.cap;
        pending_marks += r.marks;
        rows[c] = r;
    }
    CUDA_TRY(d->h_progress.reserve((size_t)n_chunks * kMaxMarks * sizeof(uint32_t)));
    volatile uint32_t* prog = (volatile uint32_t*)d->h_progress.p;
    memset(d->h_progress.p, 0, (size_t)n_chunks * kMaxMarks * sizeof(uint32_t));
    auto copy_mark = [&](uint32_t c, uint32_t m) {  // output columns [m * step, (m + 1) * step) of every row of chunk c
        const Rows& r = rows[c];
        const uint32_t s0 = first[c], cnt = first[c + 1] - s0;
        const uint64_t col = m ? r.at[m - 1] : 0;
        const uint64_t width = r.at[m] - col;
        return cudaMemcpy2DAsync(out + out_off[s0] + col, r.pitch, d_out + (out_off[s0] - so.lo) + col, r.pitch, width, cnt,
                                 cudaMemcpyDeviceToHost, d->copy_out);
    };

    // ctrl: [0, 16) the chunks' tickets, [64, 64 + 16 * kMaxMarks) their device-side progress counters
    const size_t ctrl_words = 64 + 16 * (size_t)kMaxMarks;
    CUDA_TRY(d->ctrl.reserve(ctrl_words * sizeof(uint32_t)));
    CUDA_TRY(cudaMemsetAsync(d->ctrl.p, 0, ctrl_words * sizeof(uint32_t), st));
    CUDA_TRY(cudaMemcpyAsync(dm, hm, (size_t)n * 4 * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    cudaEvent_t ready, ev_in[16], ev_k[16];
    uint32_t n_events = 0;
    CUDA_TRY(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    cudaEventRecord(ready, st);
    cudaStreamWaitEvent(d->copy_in, ready, 0);  // the device images are free once earlier work on `st` is done
    cudaStreamWaitEvent(d->copy_out, ready, 0);
    int rc = LZB_OK;
    cudaError_t err = cudaSuccess;
    for (uint32_t c = 0; c < n_chunks && rc == LZB_OK && err == cudaSuccess; c++) {
        const uint32_t s0 = first[c], cnt = first[c + 1] - s0;
        cudaEventCreateWithFlags(&ev_in[c], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ev_k[c], cudaEventDisableTiming);
        n_events = c + 1;
        if (!in_mapped && cin[c].hi > cin[c].lo)
            err = cudaMemcpyAsync((uint8_t*)d->d_in.p + (cin[c].lo - si.lo), in + cin[c].lo, cin[c].hi - cin[c].lo, cudaMemcpyHostToDevice,
                                  d->copy_in);
        if (err != cudaSuccess) break;
        cudaEventRecord(ev_in[c], d->copy_in);
        cudaStream_t ks = d->kstream[c % n_k];
        cudaStreamWaitEvent(ks, ready, 0);
        cudaStreamWaitEvent(ks, ev_in[c], 0);
        rc = dec_enqueue(d, d_in, dm + s0, dm + n + s0, cnt, d_out, dm + 2 * (size_t)n + s0, dm + 3 * (size_t)n + s0,
                         dm + 4 * (size_t)n + s0, (int32_t*)(dm + 5 * (size_t)n) + s0, max_lclp1, max_pb1, mode,
                         (uint32_t*)d->ctrl.p + c, ks, c % n_k, n_k, rows[c].on ? (uint32_t*)d->h_progress.p + c * kMarks : nullptr,
                         (uint32_t*)d->ctrl.p + 64 + c * kMaxMarks, rows[c].marks, rows[c].at);
        if (rc != LZB_OK) break;
        cudaEventRecord(ev_k[c], ks);
        cudaStreamWaitEvent(st, ev_k[c], 0);  // the out_len / status read-back below follows every kernel
        // A chunk whose regions tile its span may leave in one copy (bytes past out_len[i] inside a
        // stream's own [out_off, out_off + out_cap) are the stream's to clobber).  With gaps between
        // the regions the caller may keep other data there: such chunks are fetched stream by stream
        // once the lengths are known (below).
        span_copy[c] = cout[c].sum == cout[c].hi - cout[c].lo;
        if (!rows[c].on && span_copy[c] && cout[c].hi > cout[c].lo) {
            cudaStreamWaitEvent(d->copy_out, ev_k[c], 0);
            err = cudaMemcpyAsync(out + cout[c].lo, d_out + (cout[c].lo - so.lo), cout[c].hi - cout[c].lo, cudaMemcpyDeviceToHost,
                                  d->copy_out);
        }
    }
    // follow the progress counters of the row-shaped chunks
    for (uint32_t idle = 0; pending_marks && rc == LZB_OK && err == cudaSuccess;) {
        bool moved = false;
        for (uint32_t c = 0; c < n_events && err == cudaSuccess; c++) {
            Rows& r = rows[c];
            while (r.on && r.issued < r.marks && prog[c * kMarks + r.issued] >= first[c + 1] - first[c] && err == cudaSuccess) {
                err = copy_mark(c, r.issued++);
                pending_marks--;
                moved = true;
            }
        }
        if (moved) {
            idle = 0;
        } else if (++idle % 64 == 0) {
            // nothing new: if every kernel is gone (finished, or failed) the counters are final
            bool running = false;
            for (uint32_t k = 0; k < n_k; k++) running = running || cudaStreamQuery(d->kstream[k]) == cudaErrorNotReady;
            if (!running) break;
        } else {
            std::this_thread::sleep_for(std::chrono::microseconds(50));  // marks are milliseconds apart
        }
    }
    // whatever the loop did not see (a kernel that failed, or counters read just before the kernels ended)
    for (uint32_t c = 0; c < n_events && rc == LZB_OK && err == cudaSuccess; c++) {
        Rows& r = rows[c];
        if (!r.on || r.issued == r.marks) continue;
        cudaStreamWaitEvent(d->copy_out, ev_k[c], 0);
        while (r.issued < r.marks && err == cudaSuccess) err = copy_mark(c, r.issued++);
    }
    if (rc == LZB_OK && err == cudaSuccess)
        err = cudaMemcpyAsync(hm + 4 * (size_t)n, dm + 4 * (size_t)n, (size_t)n * (sizeof(uint64_t) + sizeof(int32_t)),
                              cudaMemcpyDeviceToHost, st);
    cudaError_t e1 = cudaStreamSynchronize(st);
    if (rc == LZB_OK && err == cudaSuccess && e1 == cudaSuccess) {
        const uint64_t* lens = hm + 4 * (size_t)n;
        for (uint32_t c = 0; c < n_events && err == cudaSuccess; c++) {
            if (rows[c].on || span_copy[c]) continue;
            for (uint32_t i = first[c]; i < first[c + 1] && err == cudaSuccess; i++)
                if (lens[i])
                    err = cudaMemcpyAsync(out + out_off[i], d_out + (out_off[i] - so.lo), lens[i], cudaMemcpyDeviceToHost, d->copy_out);
        }
    }
    const cudaError_t e2 = cudaStreamSynchronize(d->copy_out), e3 = cudaStreamSynchronize(d->copy_in);
    for (uint32_t k = 0; k < n_k; k++) {
        const cudaError_t ek = cudaStreamSynchronize(d->kstream[k]);
        if (e1 == cudaSuccess) e1 = ek;
    }
    cudaEventDestroy(ready);
    for (uint32_t c = 0; c < n_events; c++) {
        cudaEventDestroy(ev_in[c]);
        cudaEventDestroy(ev_k[c]);
    }
    if (rc != LZB_OK) return rc;
    CUDA_TRY(err);
    CUDA_TRY(e1);
    CUDA_TRY(e2);
    CUDA_TRY(e3);
    memcpy(out_len, hm + 4 * (size_t)n, (size_t)n * sizeof(uint64_t));
    memcpy(status, hm + 5 * (size_t)n, (size_t)n * sizeof(int32_t));
    return LZB_OK;
}

int lzb_dec_code(lzb_dec* d, const uint8_t* in, uint64_t in_len, uint8_t* out, uint64_t out_cap, int64_t out_size,
                 uint64_t* written) {
    if (!d) return fail(LZB_E_ARG, "null handle");
    if ((!in && in_len) || (!out && out_cap)) return fail(LZB_E_ARG, "null buffer");
    if (written) *written = 0;
    // Frame the payload the way LzmaAlone does (LzmaAlone.java:208-217) so the
    // single-stream call shares the batch kernel.
    std::vector<uint8_t> framed;
    try {
        framed.resize((size_t)in_len + LZB_HEADER_SIZE);
    } catch (...) {
        return fail(LZB_E_NOMEM, "out of host memory");
    }
    memcpy(framed.data(), d->props, 5);
    for (int i = 0; i < 8; i++) framed[5 + i] = (uint8_t)((uint64_t)out_size >> (8 * i));
    if (in_len) memcpy(framed.data() + LZB_HEADER_SIZE, in, (size_t)in_len);
    uint64_t off = 0, len = framed.size(), ooff = 0, olen = 0;
    int32_t status = 0;
    uint8_t dummy = 0;
    int rc = lzb_dec_code_batch(d, framed.data(), &off, &len, 1, out ? out : &dummy, &ooff, &out_cap, &olen, &status);
    if (rc != LZB_OK) return rc;
    if (written) *written = olen;
    if (status == LZB_E_CAPACITY) return fail(LZB_E_CAPACITY, "output capacity %llu too small", (unsigned long long)out_cap);
    return status;
}

}  // extern "C"
