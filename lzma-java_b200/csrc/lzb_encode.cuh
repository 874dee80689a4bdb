// lzb_encode.cuh -- device-side layout shared by the encoder kernels.
//
// The encoder is a three-kernel pipeline per wave of blocks (DESIGN.md
// section "encoder"):
//   1. lzb_mf_link_kernel   one warp per block: replays the three hash-head
//                           tables of BinTree (LZ/BinTree.java:171-187,210)
//                           32 positions at a time and records, per position,
//                           the hash-2 / hash-3 candidates and the successor
//                           in its hash-4 bucket;
//   2. lzb_mf_tree_kernel   one thread per hash-4 bucket: inserts the
//                           bucket's positions in order into its own binary
//                           tree (BinTree.java:212-270) and emits every
//                           position's (length, distance) list;
//   3. lzb_parse_kernel     one warp per block: getOptimum / Backward /
//                           encodeOne (LZMA/Encoder.java:335-1024) over the
//                           precomputed lists, range coder included.
// Steps 1-2 exploit that the match list of a position is a pure function of
// the data: buckets own disjoint trees.
#pragma once
#include "lzb_common.cuh"
#include "lzb_kernels.h"

namespace lzb {

constexpr uint32_t kHash2Size = 1u << 10;
constexpr uint32_t kHash3Size = 1u << 16;
constexpr uint32_t kBT2HashSize = 1u << 16;
constexpr uint32_t kMfEmpty = 0xFFFFFFFFu;  // idx[] value of a position with no pairs
// A match pair (BinTree.LengthAndDistance, BinTree.java:22-39) is 6 bytes in global memory:
//   pairs[k]  (u32) = distance | (len & 7) << 29        distance < 2^29 = the largest dictionary (Encoder.java:1135-1146)
//   pairs2[k] (u16) = cont | (len >> 3) << 9            cont = GetMatchLen after "match + literal" (Encoder.java:769), <= fb <= 273
// len <= 273 needs 9 bits: the low three ride above the distance, the high six above the continuation.
constexpr int kPairDistBits = 29;
constexpr uint32_t kPairDistMask = (1u << kPairDistBits) - 1;
__host__ __device__ __forceinline__ uint32_t pair_word(uint32_t len, uint32_t dist) { return dist | ((len & 7u) << kPairDistBits); }
__host__ __device__ __forceinline__ uint16_t pair2_word(uint32_t len, uint32_t cont) { return (uint16_t)(cont | ((len >> 3) << 9)); }
__host__ __device__ __forceinline__ uint32_t pair_len(uint32_t w, uint32_t w2) { return (w >> kPairDistBits) | ((w2 >> 9) << 3); }
__host__ __device__ __forceinline__ uint32_t pair_dist(uint32_t w) { return w & kPairDistMask; }
__host__ __device__ __forceinline__ uint32_t pair_cont(uint32_t w2) { return w2 & 511u; }
// Positions are 32-bit and 2 * position + 1 indexes the tree; BinTree.Normalize (2^30 positions) is not built.
constexpr uint64_t kEncMaxBlock = (1ull << 30) - 1;
constexpr uint32_t kHeadFlag = 0x80000000u;

// Scratch of one match-finder GROUP of blocks; every per-block array is strided by the group's largest block.
// The lists it produces are temporary (bump-allocated in the order the tree threads finish); lzb_list_* then
// compacts them, position-ordered and exactly sized, into the wave's list pool (BlockLists) and this scratch
// is reused by the next group.
struct MfWave {
    const uint8_t* in;
    const uint64_t* in_off;  // already offset to the wave's first block
    const uint64_t* in_len;
    uint32_t n_blocks;
    uint32_t np;             // stride of per-position arrays = max block length + 1 (1-based positions)
    uint32_t hash_stride;    // ints per block in `heads`
    uint32_t pair_cap;       // u32 slots per block in `pairs`
    uint32_t hash_mask;
    uint32_t cyclic_size;    // dict + 1
    int32_t fb, cut;
    bool bt4;
    uint32_t* heads;         // [n_blocks][hash_stride]  zeroed
    uint32_t* next;          // [n_blocks][np]           zeroed; successor in the hash bucket
    uint32_t* prev2;         // [n_blocks][np]           hash-2 candidate | kHeadFlag if first of its bucket
    uint32_t* prev3;         // [n_blocks][np]
    uint32_t* son;           // [n_blocks][2*np]         absolute-indexed tree links
    uint32_t* idx;           // [n_blocks][np]           offset of the position's list in `pairs` or kMfEmpty
    uint16_t* cnt;           // [n_blocks][np]           number of pairs of the position (0: no list)
    uint32_t* pairs;         // [n_blocks][pair_cap]     lists: count, then count pair words (pair_word)
    uint16_t* pairs2;        // [n_blocks][pair_cap]     per pair: pair2_word -- the rest of the length and the
                             //                          "match + literal + rep0" continuation, a function of the data only
    uint32_t* pair_used;     // [n_blocks]               zeroed; bump allocator
    uint32_t* overflow;      // [1]                      zeroed; set when a block ran out of pair slots
    uint4* long_items;       // buckets longer than kLongChain: (block, last inserted position, next position, -)
    uint32_t* long_count;    // [1] zeroed
    uint32_t* long_ticket;   // [1] zeroed
};
constexpr uint32_t kLongChain = 48;  // positions a bucket's thread inserts itself before handing over

// Where the lists of one block live in the wave's list pool (byte offsets): idx[n + 1] (1-based: offset of the
// position's list in `pairs`, or kMfEmpty), then the lists in POSITION ORDER -- count, then count pair words --
// and the parallel pair2 halfwords.  Consecutive positions are neighbours in memory, so the parser's reads of a
// block walk forward through three arrays instead of hopping between sectors.
struct BlockLists {
    uint64_t idx_off, pairs_off, pairs2_off;
};
constexpr uint32_t kListTile = 2048;   // positions per CTA of the compaction kernels (256 threads x 8)
constexpr uint32_t kListSlack = 64;    // pair slots after a block's lists: the parser prefetches 32 slots blindly

struct ParseArgs {
    const uint8_t* in;       // inputs of the wave
    const uint64_t* in_off;  // already offset to the wave's first block
    const uint64_t* in_len;
    uint32_t n_blocks;
    const uint8_t* pool;     // list pool of the wave
    const BlockLists* lists; // [n_blocks]
    uint8_t* out;
    const uint64_t* out_off; // already offset to the wave's first block
    const uint64_t* out_cap;
    uint64_t* out_len;
    uint32_t* ticket;        // zeroed
    const uint32_t* order;   // ticket -> block of the wave, or nullptr for the identity
    void* opt_scratch;       // [grid * warps][4096] packed _optimum nodes
    uint16_t* lit_scratch;   // [grid * warps][0x300 << (lc+lp)] when the literal model does not fit shared memory
    int32_t dict_size, dist_table_size;
    int32_t lc, lp, pb, fb;
    bool eos, with_header;
    uint32_t slice_bytes;    // shared memory per warp
    bool lit_in_smem;        // literal coders in the warp's slice (else in lit_scratch)
    // ICodeProgress (Encoder.java:929-933, 1070-1072): when set, every stream adds what it has consumed / produced
    // to progress[0] / progress[1] (pinned host memory) each time it has advanced by kProgressStep input bytes
    unsigned long long* progress;
};
constexpr uint32_t kProgressStep = 1u << 16;

struct ParseGeometry {
    uint32_t slice_bytes, cta_table_bytes;
    int max_warps;           // streams resident per SM
    bool lit_in_smem;
};

cudaError_t upload_mf_tables();
cudaError_t launch_mf(const MfWave& w, uint32_t max_len, int num_sms, cudaStream_t st, cudaEvent_t* ev = nullptr);
// list compaction: per-tile sizes + per-block scan (tile_sum becomes the exclusive tile offsets, w_total[b] the
// block's pair words), then the gather into the pool once the host has placed the blocks
cudaError_t launch_list_scan(const MfWave& w, uint32_t max_len, uint32_t* tile_sum, uint32_t* w_total, cudaStream_t st);
cudaError_t launch_list_gather(const MfWave& w, uint32_t max_len, const uint32_t* tile_off, const BlockLists* lists, uint8_t* pool,
                               cudaStream_t st);
cudaError_t launch_parse(const ParseArgs& a, int grid, int warps, cudaStream_t st);
ParseGeometry parse_geometry(int lc, int lp, int pb, int fb, uint32_t blocks_per_sm, int force_lit);
size_t parse_opt_bytes_per_slot();

}  // namespace lzb
