// lzb_encode.cu -- encoder pipeline orchestration (placeholder until the kernels land).
#include "lzb_common.cuh"
#include "lzb_kernels.h"

namespace lzb {

void EncScratch::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

cudaError_t run_encode(const EncodeArgs&, EncScratch&, int, cudaStream_t, int* launches) {
    if (launches) *launches = 0;
    return cudaErrorNotSupported;
}

}  // namespace lzb
