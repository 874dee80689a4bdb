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
//     buffer doubles as the dictionary window (a distance never exceeds the
//     position, and the reference's rep0 >= dictionary check is kept);
//   * the lane that loads index k = len also supplies the next match byte,
//     the lane of k = len-1 the new previous byte, so lane 0 never re-reads
//     global memory after a match.
#include "lzb_common.cuh"
#include "lzb_kernels.h"

namespace lzb {

constexpr unsigned kFull = 0xFFFFFFFFu;

// Range decoder state of one stream, in lane 0's registers.  The kernel is
// issue-bound (profiles/: 63-75 % of issue slots busy with 15 streams per SM), so
// the one thing that matters is the instruction count of a bit decode.  `bit_s`
// is written in PTX against a shared-memory byte address: 13 instructions
// (LDS, SHF, IMAD, ISETP, IADD, SEL, @IADD, SEL, IADD, SHF, IADD, STS, SEL).
struct RangeDec {
    uint32_t range, code, nextb, ip, len;
    const uint8_t* in;

    // InputStream.read(): bytes past the end read as -1, which the reference
    // ORs into _code as all ones (RangeDecoder.java:23,36).
    __device__ __forceinline__ void fetch() {
        nextb = 0xFFFFFFFFu;
        if (ip < len) nextb = __ldg(in + ip);
        ip++;
    }
    __device__ __forceinline__ void init(const uint8_t* p, uint32_t n) {  // RangeDecoder.java:19-25
        in = p;
        len = n;
        ip = 0;
        code = 0;
        range = 0xFFFFFFFFu;
        fetch();
#pragma unroll 1
        for (int i = 0; i < 5; i++) {
            code = (code << 8) | nextb;
            fetch();
        }
    }
    __device__ __forceinline__ void normalize() {
        if (range < kTopValue) {
            range <<= 8;
            code = (code << 8) | nextb;
            fetch();
        }
    }
    // RangeDecoder.DecodeBit (:43-64) on a shared-memory probability, branch-free:
    //   bit 0: p += (2048 - p) >> 5      bit 1: p -= p >> 5
    // are both  p -= (p - k) >> 5 (arithmetic shift) with k = 2017 (= 2048 - 31) resp. 0.
    __device__ __forceinline__ uint32_t bit_s(uint32_t saddr) {
        uint32_t b;
        asm volatile(
            "{\n\t"
            ".reg .pred z;\n\t"
            ".reg .u16 ph;\n\t"
            ".reg .u32 p, t, bound, r1, k;\n\t"
            ".reg .s32 d;\n\t"
            "ld.shared.u16 ph, [%3];\n\t"
            "cvt.u32.u16 p, ph;\n\t"
            "shr.u32 t, %0, 11;\n\t"
            "mul.lo.u32 bound, t, p;\n\t"
            "setp.lt.u32 z, %1, bound;\n\t"
            "sub.u32 r1, %0, bound;\n\t"
            "selp.u32 %0, bound, r1, z;\n\t"
            "@!z sub.u32 %1, %1, bound;\n\t"
            "selp.u32 k, 2017, 0, z;\n\t"
            "sub.s32 d, p, k;\n\t"
            "shr.s32 d, d, 5;\n\t"
            "sub.u32 p, p, d;\n\t"
            "cvt.u16.u32 ph, p;\n\t"
            "st.shared.u16 [%3], ph;\n\t"
            "selp.u32 %2, 0, 1, z;\n\t"
            "}"
            : "+r"(range), "+r"(code), "=r"(b)
            : "r"(saddr)
            : "memory");
        normalize();
        return b;
    }
    // the same on a generic pointer (literal coders spilled to global memory)
    __device__ __forceinline__ uint32_t bit_g(uint16_t* prob) {
        const uint32_t p0 = *prob;
        const uint32_t bound = (range >> kNumBitModelTotalBits) * p0;
        const bool one = code >= bound;
        const int32_t k = one ? 0 : (kBitModelTotal - 31);
        range = one ? range - bound : bound;
        if (one) code -= bound;
        *prob = (uint16_t)((int32_t)p0 - (((int32_t)p0 - k) >> kNumMoveBits));
        normalize();
        return one ? 1u : 0u;
    }
    // RangeDecoder.DecodeDirectBits (:27-41)
    __device__ __forceinline__ uint32_t direct(int nbits) {
        uint32_t result = 0;
#pragma unroll 1
        for (int i = nbits; i != 0; i--) {
            range >>= 1;
            const uint32_t t = (code - range) >> 31;
            code -= range & (t - 1);
            result = (result << 1) | (1 - t);
            normalize();
        }
        return result;
    }
    // BitTreeDecoder.Decode (BitTreeDecoder.java:19-25); `m2` walks the tree as a byte offset (2 * m)
    template <int NBITS>
    __device__ __forceinline__ uint32_t tree(uint32_t sbase) {
        uint32_t m2 = 2;
#pragma unroll
        for (int i = 0; i < NBITS; i++) m2 = (m2 << 1) + (bit_s(sbase + m2) << 1);
        return (m2 >> 1) - (1u << NBITS);
    }
    __device__ __forceinline__ uint32_t tree_n(uint32_t sbase, int nbits) {
        uint32_t m2 = 2;
#pragma unroll 1
        for (int i = 0; i < nbits; i++) m2 = (m2 << 1) + (bit_s(sbase + m2) << 1);
        return (m2 >> 1) - (1u << nbits);
    }
    // BitTreeDecoder.ReverseDecode (:27-37) / Decoder.ReverseDecode (Decoder.java:13-23)
    __device__ __forceinline__ uint32_t reverse(uint32_t sbase, int nbits) {
        uint32_t m2 = 2, symbol = 0;
#pragma unroll 1
        for (int i = 0; i < nbits; i++) {
            const uint32_t b = bit_s(sbase + m2);
            m2 = (m2 << 1) + (b << 1);
            symbol |= b << i;
        }
        return symbol;
    }
};

// LenDecoder.Decode (Decoder.java:48-59) on the pb-strided layout; `slen` = shared address of the coder
__device__ __forceinline__ uint32_t decode_len(RangeDec& rd, uint32_t slen, int pb, uint32_t pos_state) {
    if (rd.bit_s(slen) == 0) return rd.tree<kNumLowLenBits>(slen + 2 * len_low(pb, pos_state));
    if (rd.bit_s(slen + 2) == 0) return kNumLowLenSymbols + rd.tree<kNumMidLenBits>(slen + 2 * len_mid(pb, pos_state));
    return kNumLowLenSymbols + kNumMidLenSymbols + rd.tree_n(slen + 2 * len_high(pb), kNumHighLenBits);
}

// LiteralDecoder.Decoder2.DecodeNormal (Decoder.java:70-77), unrolled
__device__ __forceinline__ uint32_t decode_literal_s(RangeDec& rd, uint32_t sprobs) {
    uint32_t m2 = 2;
#pragma unroll
    for (int i = 0; i < 8; i++) m2 = (m2 << 1) + (rd.bit_s(sprobs + m2) << 1);
    return (m2 >> 1) & 0xFF;
}
// DecodeWithMatchByte (Decoder.java:79-95): `offs` is 0x100 while the decoded bits still agree
// with the match byte (probability index ((1 + matchBit) << 8) + symbol), 0 afterwards.
__device__ __forceinline__ uint32_t decode_literal_matched_s(RangeDec& rd, uint32_t sprobs, uint32_t match_byte) {
    uint32_t symbol = 1, offs = 0x100u;
#pragma unroll 1
    do {
        match_byte <<= 1;
        const uint32_t mb = match_byte & offs;
        const uint32_t b = rd.bit_s(sprobs + 2 * (offs + mb + symbol));
        symbol = (symbol << 1) | b;
        offs &= b ? mb : ~mb;
    } while (symbol < 0x100);
    return symbol & 0xFF;
}
// both forms on a generic pointer
__device__ __forceinline__ uint32_t decode_literal_g(RangeDec& rd, uint16_t* probs, bool matched, uint32_t match_byte) {
    uint32_t symbol = 1, offs = matched ? 0x100u : 0u;
#pragma unroll 1
    do {
        match_byte <<= 1;
        const uint32_t mb = match_byte & offs;
        const uint32_t b = rd.bit_g(probs + offs + mb + symbol);
        symbol = (symbol << 1) | b;
        offs &= b ? mb : ~mb;
    } while (symbol < 0x100);
    return symbol & 0xFF;
}

enum : int { EV_MATCH = 0, EV_DONE = 1, EV_DATA_ERROR = 2, EV_CAPACITY = 3 };

template <bool LIT_SMEM>
__device__ void decode_stream(const DecodeArgs& a, uint32_t s, uint16_t* model, uint16_t* lit_global, int lane) {
    const uint64_t in_len = a.in_len[s];
    const uint8_t* in = a.in + a.in_off[s];
    uint8_t* out = a.out + a.out_off[s];
    const uint64_t cap64 = a.out_cap[s];

    // LzmaAlone.java:220-236 -- 5 property bytes + LE64 size; Decoder.java:303-318
    int status = 1;
    uint32_t pos = 0;
    if (in_len < LZB_KERNEL_HEADER) {
        status = 0;  // "input .lzma file is too short" / "Can't read stream size"
    } else if (in_len - LZB_KERNEL_HEADER >= 0xFFFFFFF0ull || cap64 >= 0xFFFFFFF0ull) {
        status = LZB_KERNEL_E_UNSUPPORTED;  // positions are 32-bit in this kernel
    } else {
        const uint32_t cap = (uint32_t)cap64;
        const uint32_t v = in[0];
        const int lc = v % 9, rem = v / 9, lp = rem % 5, pb = rem / 5;
        uint32_t dict = 0;
        uint64_t usize = 0;
        for (int i = 0; i < 4; i++) dict |= (uint32_t)in[1 + i] << (8 * i);
        for (int i = 0; i < 8; i++) usize |= (uint64_t)in[5 + i] << (8 * i);
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
            // outSize < 0 decodes until the end marker; a size beyond the capacity ends in EV_CAPACITY
            const uint32_t limit = usize > (uint64_t)cap ? 0xFFFFFFFFu : (uint32_t)usize;

            RangeDec rd;
            const uint32_t sm = (uint32_t)__cvta_generic_to_shared(model);  // shared byte address of the model
            int state = 0;
            uint32_t rep0 = 0, rep1 = 0, rep2 = 0, rep3 = 0;
            uint32_t prev_byte = 0, match_byte = 0;
            if (lane == 0) rd.init(in + LZB_KERNEL_HEADER, (uint32_t)(in_len - LZB_KERNEL_HEADER));

            for (;;) {
                uint32_t evlen = EV_DONE;  // event | len << 2
                if (lane == 0) {
                    int ev = EV_DONE;
                    uint32_t len = 0;
#pragma unroll 1
                    while (pos < limit) {  // Decoder.Code :219
                        const uint32_t pos_state = pos & pos_mask;
                        if (rd.bit_s(sm + 2 * (L.is_match + (state << pb) + pos_state)) == 0) {
                            const uint32_t coder = 0x300u * (((pos & lp_mask) << lc) + (prev_byte >> (8 - lc)));
                            if (LIT_SMEM) {
                                const uint32_t sprobs = sm + 2 * (L.literal + coder);
                                prev_byte = state >= 7 ? decode_literal_matched_s(rd, sprobs, match_byte) : decode_literal_s(rd, sprobs);
                            } else {
                                prev_byte = decode_literal_g(rd, lit_global + coder, state >= 7, match_byte);
                            }
                            if (pos >= cap) { ev = EV_CAPACITY; break; }
                            out[pos] = (uint8_t)prev_byte;
                            state = st_lit(state);
                            pos++;
                            continue;
                        }
                        const uint32_t is_rep = rd.bit_s(sm + 2 * (L.is_rep + state));
                        len = 0;
                        if (is_rep) {  // :233-259
                            if (rd.bit_s(sm + 2 * (L.is_rep_g0 + state)) == 0) {
                                if (rd.bit_s(sm + 2 * (L.is_rep0_long + (state << pb) + pos_state)) == 0) {
                                    state = st_shortrep(state);
                                    len = 1;
                                }
                            } else {
                                uint32_t distance;
                                if (rd.bit_s(sm + 2 * (L.is_rep_g1 + state)) == 0) {
                                    distance = rep1;
                                } else {
                                    if (rd.bit_s(sm + 2 * (L.is_rep_g2 + state)) == 0) {
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
                        }
                        if (len == 0) {  // one length decoder site for both coders (:255-258, :264)
                            len = kMatchMinLen + decode_len(rd, sm + 2 * (is_rep ? L.rep_len : L.len), pb, pos_state);
                            state = is_rep ? st_longrep(state) : st_match(state);
                        }
                        if (!is_rep) {  // :260-286
                            rep3 = rep2;
                            rep2 = rep1;
                            rep1 = rep0;
                            const uint32_t pos_slot = rd.tree<kNumPosSlotBits>(sm + 2 * (L.pos_slot + (len_to_pos_state(len) << kNumPosSlotBits)));
                            if (pos_slot >= kStartPosModelIndex) {
                                const int num_direct_bits = (int)(pos_slot >> 1) - 1;
                                rep0 = (2 | (pos_slot & 1)) << num_direct_bits;
                                uint32_t rprobs = sm + 2 * (L.pos_dec + rep0 - pos_slot - 1);
                                int rbits = num_direct_bits;
                                if (pos_slot >= kEndPosModelIndex) {
                                    rep0 += rd.direct(num_direct_bits - kNumAlignBits) << kNumAlignBits;
                                    rprobs = sm + 2 * L.pos_align;
                                    rbits = kNumAlignBits;
                                }
                                rep0 += rd.reverse(rprobs, rbits);
                                if (pos_slot >= kEndPosModelIndex && (int32_t)rep0 < 0) {
                                    ev = (rep0 == 0xFFFFFFFFu) ? EV_DONE : EV_DATA_ERROR;  // end marker :277-282
                                    break;
                                }
                            } else {
                                rep0 = pos_slot;
                            }
                        }
                        if (rep0 >= pos || rep0 >= dict_check) { ev = EV_DATA_ERROR; break; }  // :288-291
                        if (len > cap - pos) { ev = EV_CAPACITY; break; }
                        ev = EV_MATCH;
                        break;
                    }
                    evlen = (uint32_t)ev | (len << 2);
                }
                evlen = __shfl_sync(kFull, evlen, 0);
                const int ev = (int)(evlen & 3);
                if (ev != EV_MATCH) {
                    status = ev == EV_DONE ? 1 : (ev == EV_DATA_ERROR ? 0 : LZB_KERNEL_E_CAPACITY);
                    break;
                }
                // OutWindow.CopyBlock (OutWindow.java:53-67), all lanes: out[pos+k] = out[pos-d+(k mod d)].
                const uint32_t len = evlen >> 2;
                const uint32_t d = __shfl_sync(kFull, rep0, 0) + 1;
                pos = __shfl_sync(kFull, pos, 0);
                __syncwarp();  // order earlier stores (lane 0's literals, other lanes' copies) before these loads
                const uint8_t* src = out + pos - d;
                uint8_t* dst = out + pos;
                uint32_t last = 0, next = 0;
                // index k = len is loaded but not stored: it is the byte GetByte(rep0) will
                // return for a matched literal that follows (Decoder.java:227).
                if (d > len) {
                    for (uint32_t k = lane; k <= len; k += 32) {
                        const uint32_t b = src[k];
                        if (k < len) dst[k] = (uint8_t)b;
                        if (k == len - 1) last = b;
                        if (k == len) next = b;
                    }
                } else {
                    for (uint32_t k = lane; k <= len; k += 32) {
                        const uint32_t b = src[k % d];
                        if (k < len) dst[k] = (uint8_t)b;
                        if (k == len - 1) last = b;
                        if (k == len) next = b;
                    }
                }
                prev_byte = __shfl_sync(kFull, last, (len - 1) & 31);  // GetByte(0) :294
                match_byte = __shfl_sync(kFull, next, len & 31);
                pos += len;
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
