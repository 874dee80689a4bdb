// lzb_common.cuh -- constants and small helpers shared by the sm_100a kernels.
//
// Constants restate LZMA/Base.java:5-86 and RangeCoder/RangeBase.java:4-7 of
// rfalke/lzma-java (paths relative to src/main/java/SevenZip/Compression/).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace lzb {

constexpr int kNumRepDistances = 4;
constexpr int kNumStates = 12;
constexpr int kNumPosSlotBits = 6;
constexpr int kNumLenToPosStates = 4;
constexpr int kMatchMinLen = 2;
constexpr int kNumAlignBits = 4;
constexpr int kAlignTableSize = 1 << kNumAlignBits;
constexpr int kAlignMask = kAlignTableSize - 1;
constexpr int kStartPosModelIndex = 4;
constexpr int kEndPosModelIndex = 14;
constexpr int kNumFullDistances = 1 << (kEndPosModelIndex / 2);  // 128
constexpr int kNumLowLenBits = 3;
constexpr int kNumMidLenBits = 3;
constexpr int kNumHighLenBits = 8;
constexpr int kNumLowLenSymbols = 1 << kNumLowLenBits;
constexpr int kNumMidLenSymbols = 1 << kNumMidLenBits;
constexpr int kNumLenSymbols = kNumLowLenSymbols + kNumMidLenSymbols + (1 << kNumHighLenBits);  // 272
constexpr int kMatchMaxLen = kMatchMinLen + kNumLenSymbols - 1;                                  // 273
constexpr int kNumBitModelTotalBits = 11;
constexpr int kBitModelTotal = 1 << kNumBitModelTotalBits;
constexpr int kNumMoveBits = 5;
constexpr uint32_t kTopValue = 1u << 24;
constexpr uint16_t kProbInit = kBitModelTotal >> 1;

// Base.java:16-40 -- the 12-state machine
__host__ __device__ __forceinline__ int st_lit(int s) { return s < 4 ? 0 : (s < 10 ? s - 3 : s - 6); }
__host__ __device__ __forceinline__ int st_match(int s) { return s < 7 ? 7 : 10; }
__host__ __device__ __forceinline__ int st_shortrep(int s) { return s < 7 ? 9 : 11; }
__host__ __device__ __forceinline__ int st_longrep(int s) { return s < 7 ? 8 : 11; }
__host__ __device__ __forceinline__ bool st_is_char(int s) { return s < 7; }
// Base.java:52-58
__host__ __device__ __forceinline__ int len_to_pos_state(int len) {
    len -= kMatchMinLen;
    return len < kNumLenToPosStates ? len : kNumLenToPosStates - 1;
}

// Probability-model layout of one stream, in 16-bit slots.  Unlike the
// reference (fixed state<<4 strides, Base.kNumPosStatesBitsMax) the
// per-posState arrays are strided by the stream's own pb, which shrinks the
// pb=2 model from 1846 to 1176 slots and lets 15 streams share one SM's
// shared memory.
struct ModelLayout {
    int is_match;      // [12 << pb]   index (state << pb) + posState
    int is_rep0_long;  // [12 << pb]
    int is_rep;        // [12]
    int is_rep_g0;     // [12]
    int is_rep_g1;     // [12]
    int is_rep_g2;     // [12]
    int pos_slot;      // [4][64]
    int pos_dec;       // [115]  (kNumFullDistances - kEndPosModelIndex = 114, +1 pad)
    int pos_align;     // [16]
    int len;           // choice[2] low[(1<<pb)][8] mid[(1<<pb)][8] high[256]
    int rep_len;       // same
    int literal;       // [0x300 << (lc+lp)]  (may live outside shared memory)
    int n_fixed;       // slots before `literal`
    int n_literal;
};

__host__ __device__ inline ModelLayout make_layout(int lc, int lp, int pb) {
    ModelLayout L;
    int o = 0;
    L.is_match = o;     o += kNumStates << pb;
    L.is_rep0_long = o; o += kNumStates << pb;
    L.is_rep = o;       o += kNumStates;
    L.is_rep_g0 = o;    o += kNumStates;
    L.is_rep_g1 = o;    o += kNumStates;
    L.is_rep_g2 = o;    o += kNumStates;
    L.pos_slot = o;     o += kNumLenToPosStates << kNumPosSlotBits;
    L.pos_dec = o;      o += 116;
    L.pos_align = o;    o += kAlignTableSize;
    int len_size = 2 + (16 << pb) + 256;
    L.len = o;          o += len_size;
    L.rep_len = o;      o += len_size;
    o = (o + 7) & ~7;   // 16-byte align the literal block
    L.literal = o;
    L.n_fixed = o;
    L.n_literal = 0x300 << (lc + lp);
    return L;
}
// offsets inside a length coder block
__host__ __device__ __forceinline__ int len_low(int pb_unused, int pos_state) { (void)pb_unused; return 2 + (pos_state << 3); }
__host__ __device__ __forceinline__ int len_mid(int pb, int pos_state) { return 2 + (8 << pb) + (pos_state << 3); }
__host__ __device__ __forceinline__ int len_high(int pb) { return 2 + (16 << pb); }

}  // namespace lzb
