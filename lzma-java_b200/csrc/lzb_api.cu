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
    PinBuf h_meta;
};

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
    if (d->stream) cudaStreamDestroy(d->stream);
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

    // which streams need their literal model outside shared memory?
    CUDA_TRY(lzb::launch_decode_scan(d_in, d_in_off, d_in_len, n, ctrl + 1, st));
    add_launches(1);
    uint32_t max_spill = 0;
    CUDA_TRY(cudaMemcpyAsync(&max_spill, ctrl + 1, sizeof max_spill, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));

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
    a.ticket = ctrl;
    a.lit_scratch = nullptr;
    a.lit_stride = 0;
    if (max_spill) {
        a.lit_stride = max_spill;
        const size_t slots = (size_t)d->num_sms * lzb::kDecMaxWarps;
        CUDA_TRY(d->lit.reserve(slots * a.lit_stride * sizeof(uint16_t)));
        a.lit_scratch = (uint16_t*)d->lit.p;
    }
    CUDA_TRY(lzb::launch_decode(a, max_spill == 0, d->num_sms, st, nullptr, nullptr));
    add_launches(1);
    return LZB_OK;
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

    // Device images keep the caller's layout: one copy per direction over the
    // smallest covering span.
    const Span si = span_of(in_off, in_len, n);
    const Span so = span_of(out_off, out_cap, n);
    const size_t in_bytes = si.hi - si.lo, out_bytes = so.hi - so.lo;
    CUDA_TRY(d->d_in.reserve(in_bytes + 16));
    CUDA_TRY(d->d_out.reserve(out_bytes + 16));
    // meta: in_off in_len out_off out_cap out_len (u64 x n each) + status (i32 x n)
    const size_t meta_bytes = (size_t)n * (5 * sizeof(uint64_t) + sizeof(int32_t));
    CUDA_TRY(d->d_meta.reserve(meta_bytes));
    CUDA_TRY(d->h_meta.reserve(meta_bytes));
    uint64_t* hm = (uint64_t*)d->h_meta.p;
    for (uint32_t i = 0; i < n; i++) {
        hm[i] = in_off[i] - si.lo;
        hm[n + i] = in_len[i];
        hm[2 * (size_t)n + i] = out_off[i] - so.lo;
        hm[3 * (size_t)n + i] = out_cap[i];
    }
    uint64_t* dm = (uint64_t*)d->d_meta.p;
    CUDA_TRY(cudaMemcpyAsync(dm, hm, (size_t)n * 4 * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d->d_in.p, in + si.lo, in_bytes, cudaMemcpyHostToDevice, st));

    int rc = lzb_dec_code_batch_device(d, (const uint8_t*)d->d_in.p, dm, dm + n, n, (uint8_t*)d->d_out.p,
                                       dm + 2 * (size_t)n, dm + 3 * (size_t)n, dm + 4 * (size_t)n,
                                       (int32_t*)(dm + 5 * (size_t)n), st);
    if (rc != LZB_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(hm + 4 * (size_t)n, dm + 4 * (size_t)n, (size_t)n * (sizeof(uint64_t) + sizeof(int32_t)),
                             cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out + so.lo, d->d_out.p, out_bytes, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
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
