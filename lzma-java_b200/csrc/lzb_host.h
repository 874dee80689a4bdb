// lzb_host.h -- host-side helpers shared by the two halves of the C-ABI layer.
#pragma once
#include <cstdarg>
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/lzma_b200.h"

namespace lzbhost {

int fail(int code, const char* fmt, ...);  // records lzb_last_error(), returns code
void add_launches(int n);

#define CUDA_TRY(expr)                                                                                      \
    do {                                                                                                    \
        cudaError_t _e = (expr);                                                                            \
        if (_e != cudaSuccess)                                                                              \
            return lzbhost::fail(_e == cudaErrorMemoryAllocation ? LZB_E_NOMEM : LZB_E_CUDA, "%s: %s", #expr, \
                                 cudaGetErrorString(_e));                                                   \
    } while (0)

// grow-only device buffer
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// grow-only pinned host buffer
struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaHostAlloc(&p, bytes + 256, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = bytes + 256;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

int open_device(int device, cudaStream_t* stream, int* num_sms);

// smallest span of a base pointer covering every [off, off+len), and the sum of lengths
struct Span {
    uint64_t lo = 0, hi = 0, sum = 0;
};
inline Span span_of(const uint64_t* off, const uint64_t* len, uint32_t n) {
    Span s;
    if (n == 0) return s;
    s.lo = ~0ull;
    for (uint32_t i = 0; i < n; i++) {
        if (off[i] < s.lo) s.lo = off[i];
        if (off[i] + len[i] > s.hi) s.hi = off[i] + len[i];
        s.sum += len[i];
    }
    return s;
}

}  // namespace lzbhost
