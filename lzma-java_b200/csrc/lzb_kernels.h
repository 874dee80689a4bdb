// lzb_kernels.h -- internal interface between the C-ABI layer (lzb_api.cu)
// and the sm_100a kernels.  Not installed; the public boundary is
// include/lzma_b200.h.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#define LZB_KERNEL_HEADER 13
#define LZB_KERNEL_E_CAPACITY (-4)
#define LZB_KERNEL_E_UNSUPPORTED (-5)

namespace lzb {

// ---- decoder ---------------------------------------------------------------
// Where a stream's probability model lives.  Streams per SM are bounded by shared memory, and
// two thirds of an lc=3 model are the matched-literal tables (Decoder.java:79-95) that only a
// literal right after a match touches: the hybrid mode keeps them in global memory (L2) and
// almost doubles the resident streams.  It pays when there are more streams than kDecSmem slots.
enum DecMode : int {
    kDecSmem = 0,    // whole model in shared memory (lc + lp <= 3)
    kDecHybrid = 1,  // matched-literal tables in global memory, except their top three levels (lc + lp <= 3)
    kDecGlobal = 2,  // all literal tables in global memory (any lc, lp)
};
constexpr int kDecMaxWarps = 15;          // streams resident per SM, kDecSmem / kDecGlobal
constexpr int kDecHybridWarps = 28;       // 28 warps * 72 registers fill the register file
// Shared memory of one stream = its model slice + the input ring (a power of two, aligned to its size) + the ring's
// 64-byte header.  Model slices: the whole lc + lp = 3 model up to pb = 3 (15 088 B); in the hybrid mode the fixed part
// (pb = 3: 2800 B) + 8 normal literal trees (4096 B) + the top of the matched trees (256 B); fixed part only (pb = 4: 3696 B).
constexpr size_t kDecRingHeader = 64;
__host__ __device__ constexpr size_t dec_mode_model(int mode) { return mode == kDecSmem ? 15088 : mode == kDecHybrid ? 7152 : 3696; }
__host__ __device__ constexpr uint32_t dec_mode_ring(int mode) { return mode == kDecSmem ? 256u : mode == kDecHybrid ? 512u : 1024u; }
__host__ __device__ constexpr int dec_mode_warps(int mode) { return mode == kDecHybrid ? kDecHybridWarps : kDecMaxWarps; }
__host__ __device__ constexpr size_t dec_mode_slice(int mode) { return dec_mode_model(mode) + kDecRingHeader + dec_mode_ring(mode); }
static_assert(kDecMaxWarps * dec_mode_slice(kDecSmem) <= 232448 && kDecHybridWarps * dec_mode_slice(kDecHybrid) <= 232448, "227 KB per CTA");

constexpr uint32_t kDecMaxMarks = 32;
struct DecodeArgs {
    const uint8_t* in;
    const uint64_t* in_off;
    const uint64_t* in_len;
    uint8_t* out;
    const uint64_t* out_off;
    const uint64_t* out_cap;
    uint64_t* out_len;
    int32_t* status;
    uint32_t n;
    uint32_t* ticket;        // zeroed before launch
    uint16_t* lit_scratch;   // literal tables kept in global memory (kDecHybrid, kDecGlobal)
    size_t lit_stride;       // in 16-bit slots, per resident stream
    // Optional progress report for a host that copies output back while the kernel still runs:
    // progress[m] becomes n once the output below mark_at[m] bytes of ALL n streams is complete and
    // visible system-wide (the last mark, m = marks - 1: all streams finished).  The streams count themselves
    // in device memory (progress_dev); only the one that completes a mark touches host memory -- system-scope
    // fences and atomics from thousands of warps serialise on the PCIe link (0.7 ms per mark for 4096 streams).
    uint32_t* progress = nullptr;  // pinned host memory (device-accessible), zeroed before launch
    uint32_t* progress_dev = nullptr;  // [marks] device counters, zeroed before launch: streams that passed each mark
    uint32_t marks = 0;
    uint32_t mark_at[kDecMaxMarks] = {};  // increasing; entries [0, marks - 1) are used
};

// header pre-pass over well-formed streams: d_max_lclp1[0] = max(lc + lp + 1) (0 if there is none), [1] = max(pb + 1)
cudaError_t launch_decode_scan(const uint8_t* in, const uint64_t* in_off, const uint64_t* in_len, uint32_t n,
                               uint32_t* d_max_lclp1, cudaStream_t st);
cudaError_t launch_decode(const DecodeArgs& a, int mode, int num_sms, cudaStream_t st, int* grid_out, int* warps_out);

// ---- encoder ---------------------------------------------------------------
struct EncodeArgs {
    const uint8_t* in;
    const uint64_t* in_off;
    const uint64_t* in_len;
    uint8_t* out;
    const uint64_t* out_off;
    const uint64_t* out_cap;
    uint64_t* out_len;       // UINT64_MAX when out_cap was too small
    uint32_t n;
    uint64_t max_in_len;
    int32_t dict_size, fb;
    bool bt4;
    int32_t lc, lp, pb;
    bool eos, with_header;
    // ICodeProgress: called from the calling thread while the parser runs, with batch totals (in, out); may be null
    void (*progress_fn)(void* user, uint64_t in_size, uint64_t out_size) = nullptr;
    void* progress_user = nullptr;
    // developer / test hooks, read once when the handle is created (lzb_enc_create)
    int32_t tune_warps = 0;      // LZB_ENC_WARPS: parser streams per SM (0 = automatic)
    int32_t tune_pair_mul = 0;   // LZB_PAIR_MUL: initial match-pair budget in slots per input byte (0 = default)
    int32_t tune_lit = -1;       // LZB_ENC_LIT: 0 = literal coders in shared memory, 1 = in global memory (-1 = automatic)
    int32_t tune_group = 0;      // LZB_ENC_GROUP: at most this many blocks per match-finder group (0 = what the scratch budget allows)
    int64_t tune_pool = 0;       // LZB_ENC_POOL_MB: cap of the list pool in bytes (0 = sized from the batch): small values force many waves
    bool tune_fifo = false;      // LZB_ENC_FIFO: plain block order inside a wave
    int32_t tune_inflight = 0;   // LZB_ENC_INFLIGHT: match-finder groups in flight (0 = kEncGroupsDefault)
    bool tune_blocked = false;   // LZB_ENC_BLOCKED: cost-sorted order, neighbours on one SM (no dealing across the SMs)
    bool tune_timing = false;    // LZB_ENC_TIMING: phase times of every wave on stderr
};

// device scratch owned by an encoder handle (grow-only): the match finder's group scratch (`p`), the wave's
// list pool and the parser's per-slot areas
constexpr int kEncGroupsInFlight = 8;   // buffers / streams a handle can hold; kEncGroupsDefault of them are used unless tuned
constexpr int kEncGroupsDefault = 6;
struct EncScratch {
    void* gp[kEncGroupsInFlight] = {};        // group scratch, one per group in flight
    size_t gcap[kEncGroupsInFlight] = {};
    cudaStream_t gs[kEncGroupsInFlight] = {};  // [0] unused (the caller's stream)
    cudaEvent_t gev[kEncGroupsInFlight] = {};
    bool streams_ready = false;
    uint32_t* h_totals = nullptr;  // pinned: per-group pair totals + overflow flag, one row per group in flight
    size_t h_totals_cap = 0;
    void* pool = nullptr;
    size_t pool_cap = 0;
    void* fixed = nullptr;
    size_t fixed_cap = 0;
    unsigned long long* h_progress = nullptr;  // pinned: (in, out) totals of the call in flight
    void release();
};

constexpr uint64_t kEncMaxBlockBytes = (1ull << 30) - 1;  // one stream; BinTree.Normalize (BinTree.java:358-375) is not built
constexpr int kEncWarpsLitSmem = 9;        // parser streams per SM, literal coders in shared memory (shared memory decides, see parse_geometry)
constexpr int kEncMaxWarps = 14;           // ... with the literal coders in global memory (15 and 16 measured no faster: profiles/r02_parse_residency_ab.log)

// where the match finder left the lists of the first block of a batch (trace tap)
struct MfTrace {
    const uint32_t* idx = nullptr;    // [len + 1], 1-based: offset into pairs or 0xFFFFFFFF (the match finder's temporary lists)
    const uint32_t* pairs = nullptr;  // count, then count x pair_word (lzb_encode.cuh)
    const uint16_t* pairs2 = nullptr; // pair2_word per pair
    uint32_t pair_words = 0;          // words in use
};

// Enqueue the whole encode pipeline for the batch on `st`.  With `mf_only` the match finder
// alone runs (one wave) and *mf_only describes the lists of block 0; the stream is synchronised.
cudaError_t run_encode(const EncodeArgs& a, EncScratch& scratch, int num_sms, cudaStream_t st, int* launches,
                       MfTrace* mf_only = nullptr);

}  // namespace lzb
