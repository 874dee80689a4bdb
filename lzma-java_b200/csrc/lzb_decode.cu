// lzb_decode.cu -- batch LZMA decoder for sm_100a: one warp per .lzma stream.
//
// Replaces Decoder.Code (LZMA/Decoder.java:205-301) with its helpers
// RangeDecoder (RangeCoder/RangeDecoder.java:19-64), BitTreeDecoder
// (RangeCoder/BitTreeDecoder.java:19-37), LenDecoder / LiteralDecoder
// (Decoder.java:25-127) and OutWindow.CopyBlock (LZ/OutWindow.java:53-67) of
// rfalke/lzma-java.  Design (DESIGN.md section "decoder"):
//   * one persistent CTA per SM, 15 warps, each warp owns one stream at a
//     time and pulls the next one from a global ticket counter;
//   * the stream's whole probability model (7 320 16-bit slots at lc3 lp0
//     pb2) lives in the warp's private 15 488-byte slice of shared memory;
//   * lane 0 runs the serial range-decoder chain with range/code and a
//     one-byte input lookahead in registers;
//   * matches are copied by all 32 lanes: out[pos+k] = out[pos-d+(k mod d)],
//     which is order-free even when the match overlaps itself; the output
//     buffer doubles as the dictionary window (block <= dictionary);
//   * the lane that loads index k = len also supplies the next match byte,
//     the lane of k = len-1 the new previous byte, so lane 0 never re-reads
//     global memory after a match.
#include "lzb_common.cuh"
#include "lzb_kernels.h"

namespace lzb {

constexpr unsigned kFull = 0xFFFFFFFFu;

struct RangeDec {
    uint32_t range, code, nextb;
    const uint8_t* p;
    const uint8_t* end;

    // InputStream.read(): bytes past the end read as -1, which the reference
    // ORs into _code as all ones (RangeDecoder.java:23,36).
    __device__ __forceinline__ void prefetch() { nextb = (p < end) ? (uint32_t)__ldg(p) : 0xFFFFFFFFu; }
    __device__ __forceinline__ uint32_t read() {
        uint32_t b = nextb;
        ++p;
        prefetch();
        return b;
    }
    __device__ __forceinline__ void init(const uint8_t* in, const uint8_t* e) {  // RangeDecoder.java:19-25
        p = in;
        end = e;
        code = 0;
        range = 0xFFFFFFFFu;
        prefetch();
#pragma unroll
        for (int i = 0; i < 5; i++) code = (code << 8) | read();
    }
    __device__ __forceinline__ void normalize() {
        if (range < kTopValue) {
            range <<= 8;
            code = (code << 8) | read();
        }
    }
    // RangeDecoder.DecodeBit (:43-64)
    template <typename P>
    __device__ __forceinline__ uint32_t bit(P* prob) {
        uint32_t p0 = *prob;
        uint32_t bound = (range >> kNumBitModelTotalBits) * p0;
        uint32_t b;
        if (code < bound) {
            range = bound;
            *prob = (uint16_t)(p0 + ((kBitModelTotal - p0) >> kNumMoveBits));
            b = 0;
        } else {
            range -= bound;
            code -= bound;
            *prob = (uint16_t)(p0 - (p0 >> kNumMoveBits));
            b = 1;
        }
        normalize();
        return b;
    }
    // RangeDecoder.DecodeDirectBits (:27-41)
    __device__ __forceinline__ uint32_t direct(int nbits) {
        uint32_t result = 0;
        for (int i = nbits; i != 0; i--) {
            range >>= 1;
            uint32_t t = (code - range) >> 31;
            code -= range & (t - 1);
            result = (result << 1) | (1 - t);
            normalize();
        }
        return result;
    }
    // BitTreeDecoder.Decode (BitTreeDecoder.java:19-25)
    template <int NBITS>
    __device__ __forceinline__ uint32_t tree(uint16_t* probs) {
        uint32_t m = 1;
#pragma unroll
        for (int i = 0; i < NBITS; i++) m = (m << 1) + bit(probs + m);
        return m - (1u << NBITS);
    }
    // BitTreeDecoder.ReverseDecode (:27-37) / Decoder.ReverseDecode (Decoder.java:13-23)
    __device__ __forceinline__ uint32_t reverse(uint16_t* probs, int nbits) {
        uint32_t m = 1, symbol = 0;
        for (int i = 0; i < nbits; i++) {
            uint32_t b = bit(probs + m);
            m = (m << 1) + b;
            symbol |= b << i;
        }
        return symbol;
    }
};

// LenDecoder.Decode (Decoder.java:48-59) on the pb-strided layout
__device__ __forceinline__ uint32_t decode_len(RangeDec& rd, uint16_t* lenp, int pb, uint32_t pos_state) {
    if (rd.bit(lenp + 0) == 0) return rd.tree<kNumLowLenBits>(lenp + len_low(pb, pos_state));
    if (rd.bit(lenp + 1) == 0) return kNumLowLenSymbols + rd.tree<kNumMidLenBits>(lenp + len_mid(pb, pos_state));
    return kNumLowLenSymbols + kNumMidLenSymbols + rd.tree<kNumHighLenBits>(lenp + len_high(pb));
}

// LiteralDecoder.Decoder2.DecodeNormal / DecodeWithMatchByte (Decoder.java:70-95)
template <typename P>
__device__ __forceinline__ uint32_t decode_literal(RangeDec& rd, P* probs, bool matched, uint32_t match_byte) {
    uint32_t symbol = 1;
    if (matched) {
        do {
            uint32_t match_bit = (match_byte >> 7) & 1;
            match_byte <<= 1;
            uint32_t b = rd.bit(probs + ((1 + match_bit) << 8) + symbol);
            symbol = (symbol << 1) | b;
            if (match_bit != b) break;
        } while (symbol < 0x100);
    }
    while (symbol < 0x100) symbol = (symbol << 1) | rd.bit(probs + symbol);
    return symbol & 0xFF;
}

enum : int { EV_MATCH = 0, EV_DONE = 1, EV_DATA_ERROR = 2, EV_CAPACITY = 3 };

template <bool LIT_SMEM>
__device__ void decode_stream(const DecodeArgs& a, uint32_t s, uint16_t* model, uint16_t* lit_global, int lane) {
    const uint64_t in_len = a.in_len[s];
    const uint8_t* in = a.in + a.in_off[s];
    uint8_t* out = a.out + a.out_off[s];
    const uint64_t cap = a.out_cap[s];

    // LzmaAlone.java:220-236 -- 5 property bytes + LE64 size; Decoder.java:303-318
    int status = 1;
    uint64_t pos = 0;
    if (in_len < LZB_KERNEL_HEADER) {
        status = 0;  // "input .lzma file is too short" / "Can't read stream size"
    } else {
        const uint32_t v = in[0];
        const int lc = v % 9, rem = v / 9, lp = rem % 5, pb = rem / 5;
        uint32_t dict = 0;
        uint64_t usize = 0;
        for (int i = 0; i < 4; i++) dict |= (uint32_t)in[1 + i] << (8 * i);
        for (int i = 0; i < 8; i++) usize |= (uint64_t)in[5 + i] << (8 * i);
        const int64_t out_size = (int64_t)usize;
        // SetLcLpPb (:172-182) rejects pb > 4 (lc, lp are bounded by the
        // arithmetic); SetDictionarySize (:160-170) rejects a negative size.
        if (pb > 4 || (int32_t)dict < 0) {
            status = 0;
        } else {
            const ModelLayout L = make_layout(lc, lp, pb);
            uint16_t* lit = LIT_SMEM ? model + L.literal : lit_global;
            for (int i = lane; i < L.n_fixed; i += 32) model[i] = kProbInit;  // Decoder.Init :184-203
            for (int i = lane; i < L.n_literal; i += 32) lit[i] = kProbInit;
            __syncwarp();

            const uint32_t dict_check = dict > 1 ? dict : 1;  // m_DictionarySizeCheck :166
            const uint32_t pos_mask = (1u << pb) - 1, lp_mask = (1u << lp) - 1;
            const uint64_t limit = out_size < 0 ? ~0ull : (uint64_t)out_size;

            RangeDec rd;
            int state = 0;
            uint32_t rep0 = 0, rep1 = 0, rep2 = 0, rep3 = 0;
            uint32_t prev_byte = 0, match_byte = 0;
            if (lane == 0) rd.init(in + LZB_KERNEL_HEADER, in + in_len);

            for (;;) {
                int ev = EV_DONE;
                uint32_t len = 0;
                if (lane == 0) {
                    while (pos < limit) {  // Decoder.Code :219
                        const uint32_t pos_state = (uint32_t)pos & pos_mask;
                        if (rd.bit(model + L.is_match + (state << pb) + pos_state) == 0) {
                            uint16_t* probs = lit + 0x300u * ((((uint32_t)pos & lp_mask) << lc) + (prev_byte >> (8 - lc)));
                            prev_byte = decode_literal(rd, probs, !st_is_char(state), match_byte);
                            if (pos >= cap) { ev = EV_CAPACITY; break; }
                            out[pos] = (uint8_t)prev_byte;
                            // the byte the next matched literal would compare with (GetByte(rep0), :227)
                            // is only needed in state >= 7, i.e. never directly after a literal
                            state = st_lit(state);
                            pos++;
                            continue;
                        }
                        if (rd.bit(model + L.is_rep + state) == 1) {  // :233-259
                            len = 0;
                            if (rd.bit(model + L.is_rep_g0 + state) == 0) {
                                if (rd.bit(model + L.is_rep0_long + (state << pb) + pos_state) == 0) {
                                    state = st_shortrep(state);
                                    len = 1;
                                }
                            } else {
                                uint32_t distance;
                                if (rd.bit(model + L.is_rep_g1 + state) == 0) {
                                    distance = rep1;
                                } else {
                                    if (rd.bit(model + L.is_rep_g2 + state) == 0) {
                                        distance = rep2;
                                    } else {
                                        distance = rep3;
                                        rep3 = rep2;
                                    }
                                    rep2 = rep1;
                                }
                                rep1 = rep0;
                                rep0 = distance;
                            }
                            if (len == 0) {
                                len = decode_len(rd, model + L.rep_len, pb, pos_state) + kMatchMinLen;
                                state = st_longrep(state);
                            }
                        } else {  // :260-286
                            rep3 = rep2;
                            rep2 = rep1;
                            rep1 = rep0;
                            len = kMatchMinLen + decode_len(rd, model + L.len, pb, pos_state);
                            state = st_match(state);
                            const uint32_t pos_slot = rd.tree<kNumPosSlotBits>(model + L.pos_slot + (len_to_pos_state(len) << kNumPosSlotBits));
                            if (pos_slot >= kStartPosModelIndex) {
                                const int num_direct_bits = (int)(pos_slot >> 1) - 1;
                                rep0 = (2 | (pos_slot & 1)) << num_direct_bits;
                                if (pos_slot < kEndPosModelIndex) {
                                    rep0 += rd.reverse(model + L.pos_dec + rep0 - pos_slot - 1, num_direct_bits);
                                } else {
                                    rep0 += rd.direct(num_direct_bits - kNumAlignBits) << kNumAlignBits;
                                    rep0 += rd.reverse(model + L.pos_align, kNumAlignBits);
                                    if ((int32_t)rep0 < 0) {
                                        ev = (rep0 == 0xFFFFFFFFu) ? EV_DONE : EV_DATA_ERROR;  // end marker :277-282
                                        break;
                                    }
                                }
                            } else {
                                rep0 = pos_slot;
                            }
                        }
                        if ((uint64_t)rep0 >= pos || rep0 >= dict_check) { ev = EV_DATA_ERROR; break; }  // :288-291
                        if (pos + len > cap) { ev = EV_CAPACITY; break; }
                        ev = EV_MATCH;
                        break;
                    }
                }
                ev = __shfl_sync(kFull, ev, 0);
                if (ev != EV_MATCH) {
                    status = ev == EV_DONE ? 1 : (ev == EV_DATA_ERROR ? 0 : LZB_KERNEL_E_CAPACITY);
                    pos = (uint64_t)__shfl_sync(kFull, (uint32_t)pos, 0) | ((uint64_t)__shfl_sync(kFull, (uint32_t)(pos >> 32), 0) << 32);
                    break;
                }
                // OutWindow.CopyBlock (OutWindow.java:53-67), all lanes.
                len = __shfl_sync(kFull, len, 0);
                const uint32_t d = __shfl_sync(kFull, rep0, 0) + 1;
                pos = (uint64_t)__shfl_sync(kFull, (uint32_t)pos, 0) | ((uint64_t)__shfl_sync(kFull, (uint32_t)(pos >> 32), 0) << 32);
                __syncwarp();  // order lane 0's literal stores before the lanes' loads
                const uint8_t* src = out + pos - d;
                uint8_t* dst = out + pos;
                uint32_t last = 0, next = 0;
                // index k = len is loaded but not stored: it is the byte GetByte(rep0) will
                // return for a matched literal that follows (Decoder.java:227).
                for (uint32_t k = lane; k <= len; k += 32) {
                    const uint32_t b = src[k < d ? k : k % d];
                    if (k < len) dst[k] = (uint8_t)b;
                    if (k == len - 1) last = b;
                    if (k == len) next = b;
                }
                prev_byte = __shfl_sync(kFull, last, (len - 1) & 31);  // GetByte(0) :294
                match_byte = __shfl_sync(kFull, next, len & 31);
                pos += len;
                __syncwarp();
            }
        }
    }
    if (lane == 0) {
        a.out_len[s] = pos;
        a.status[s] = status;
    }
}

template <bool LIT_SMEM>
__global__ void __launch_bounds__(kDecMaxWarps * 32, 1) lzb_decode_kernel(DecodeArgs a) {
    extern __shared__ __align__(16) uint16_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint16_t* model = smem + (size_t)warp * (kDecSliceBytes / 2);
    uint16_t* lit_global = LIT_SMEM ? nullptr
                                    : a.lit_scratch + ((size_t)blockIdx.x * (blockDim.x >> 5) + warp) * a.lit_stride;
    for (;;) {
        uint32_t s = 0;
        if (lane == 0) s = atomicAdd(a.ticket, 1u);
        s = __shfl_sync(kFull, s, 0);
        if (s >= a.n) break;
        decode_stream<LIT_SMEM>(a, s, model, lit_global, lane);
    }
}

// Header pre-pass: the largest literal model (in 16-bit slots) among streams
// whose model does not fit a warp's shared-memory slice; 0 if all fit.
__global__ void lzb_decode_scan_headers(const uint8_t* in, const uint64_t* in_off, const uint64_t* in_len, uint32_t n,
                                        uint32_t* max_spill) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (in_len[i] < LZB_KERNEL_HEADER) return;
    const uint32_t v = in[in_off[i]];
    const int lc = v % 9, rem = v / 9, lp = rem % 5, pb = rem / 5;
    if (pb > 4) return;
    const ModelLayout L = make_layout(lc, lp, pb);
    if ((size_t)(L.n_fixed + L.n_literal) * 2 > kDecSliceBytes) atomicMax(max_spill, (uint32_t)L.n_literal);
}

cudaError_t launch_decode_scan(const uint8_t* in, const uint64_t* in_off, const uint64_t* in_len, uint32_t n,
                               uint32_t* d_max_spill, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    lzb_decode_scan_headers<<<(n + 255) / 256, 256, 0, st>>>(in, in_off, in_len, n, d_max_spill);
    return cudaGetLastError();
}

cudaError_t launch_decode(const DecodeArgs& a, bool lit_in_smem, int num_sms, cudaStream_t st, int* grid_out, int* warps_out) {
    if (a.n == 0) return cudaSuccess;
    int warps = (int)((a.n + (uint32_t)num_sms - 1) / (uint32_t)num_sms);
    if (warps > kDecMaxWarps) warps = kDecMaxWarps;
    if (warps < 1) warps = 1;
    int grid = (int)((a.n + (uint32_t)warps - 1) / (uint32_t)warps);
    if (grid > num_sms) grid = num_sms;
    const size_t smem = (size_t)warps * kDecSliceBytes;
    auto kern = lit_in_smem ? lzb_decode_kernel<true> : lzb_decode_kernel<false>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kDecMaxWarps * kDecSliceBytes));
    if (err != cudaSuccess) return err;
    kern<<<grid, warps * 32, smem, st>>>(a);
    if (grid_out) *grid_out = grid;
    if (warps_out) *warps_out = warps;
    return cudaGetLastError();
}

}  // namespace lzb
