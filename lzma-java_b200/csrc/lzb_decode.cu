// lzb_decode.cu -- batch LZMA decoder for sm_100a: one warp per .lzma stream.
//
// Replaces Decoder.Code (LZMA/Decoder.java:205-301) with its helpers
// RangeDecoder (RangeCoder/RangeDecoder.java:19-64), BitTreeDecoder
// (RangeCoder/BitTreeDecoder.java:19-37), LenDecoder / LiteralDecoder
// (Decoder.java:25-127) and OutWindow.CopyBlock (LZ/OutWindow.java:53-67) of
// rfalke/lzma-java.  Design (DESIGN.md section "decoder"):
//   * one persistent CTA per SM; each warp owns one stream at a time (the
//     first by position, the following ones from a global ticket counter);
//   * the stream's probability model lives in the warp's private slice of
//     shared memory -- all of it (kDecSmem, 15 streams per SM), or all but the
//     matched-literal tables, which then sit in global memory (kDecHybrid,
//     28 streams per SM), or all but the literal coders (kDecGlobal, lc+lp > 3);
//     see DecMode in lzb_kernels.h.  Residency is what the throughput hangs on;
//   * lane 0 runs the serial range-decoder chain with range/code and a
//     one-byte input lookahead in registers; the bit decode is PTX;
//   * matches are copied by all 32 lanes: out[pos+k] = out[pos-d+(k mod d)],
//     which is order-free even when the match overlaps itself; the output
//     buffer doubles as the dictionary window (a distance never exceeds the
//     position, and the reference's rep0 >= dictionary check is kept);
//   * nothing waits for a copy: its last 32 bytes stay in registers until the
//     next copy, and lane 0 prefetches the two bytes a following literal needs.
#include "lzb_common.cuh"
#include "lzb_kernels.h"

namespace lzb {

constexpr unsigned kFull = 0xFFFFFFFFu;

// ---- compressed input: a per-warp shared-memory ring filled by TMA bulk copies -------------------
// The reference reads its input one byte per normalisation (RangeDecoder.java:23,36,50,59).  Here a
// stream's compressed bytes are staged through a small ring in shared memory: lane 0 issues
// cp.async.bulk (the TMA unit's 1-D copy, UBLKCP in SASS) a quarter of the ring at a time, completion
// is counted on one mbarrier per quarter, and normalize() takes its byte with one LDS from an address
// that is a single LOP3 away from the read position (the ring is aligned to its size).
// The ring is checked once per SYMBOL, not per byte: a symbol consumes at most 48 bytes (a match with
// 26 direct bits), a quarter is at least 64 bytes, and at every symbol start the quarter under the
// read position and the two after it have landed while the fourth is in flight.  So the bit decoder
// itself carries no bounds test at all: past the end of the payload the ring simply holds 0xFF, which
// is what the reference's InputStream.read() = -1 ORs into the code (RangeDecoder.java:23,36).
// A quarter lasts a stream ~100 us, which hides HBM latency -- and PCIe latency too: the source may be
// pinned HOST memory, so a host-buffer batch needs no input copy before the launch.
// Ring position `rp` counts bytes from g0, the 16-byte-aligned address at or below the payload start.
struct __align__(16) InRing {
    uint64_t mbar[4];   // one per quarter
    uint64_t g0;        // global address of ring position 0 (16-byte aligned)
    uint32_t end_rp;    // ring position of the first byte past the payload
    uint32_t total;     // bytes to fetch from g0 on (a multiple of 16)
    uint32_t state;     // bits 0-3: phase parity of mbar[q]; bits 4-7: a copy into quarter q is pending
    uint32_t pad[3];
};
static_assert(sizeof(InRing) == 64, "InRing header");
constexpr uint32_t kInRingHeader = (uint32_t)sizeof(InRing);

__device__ __forceinline__ void mbar_init(uint32_t mbar_s) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_s) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar_s, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_LOOP;\n\t"
        "}" ::"r"(mbar_s), "r"(parity) : "memory");
}
// one TMA bulk copy global -> shared of `bytes` (a multiple of 16), completion counted on the mbarrier
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst_s, uint64_t src_g, uint32_t bytes, uint32_t mbar_s) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic accesses of this quarter precede the async write
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_s), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_s), "l"(src_g), "r"(bytes), "r"(mbar_s) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// Request absolute chunk `c` of the stream (ring positions [c * chunk, (c + 1) * chunk)) into its
// quarter, or fill the quarter with 0xFF when the payload ends before it.
__device__ __forceinline__ void ring_request(uint32_t hdr_s, uint32_t ring_s, uint32_t chunk, uint32_t c) {
    const uint32_t q = c & 3u, from = c * chunk, total = lds_u32(hdr_s + 44);
    const uint32_t dst = ring_s + q * chunk;
    if (from < total) {
        uint32_t bytes = total - from;
        if (bytes > chunk) bytes = chunk;
        uint64_t g0;
        asm volatile("ld.shared.u64 %0, [%1];" : "=l"(g0) : "r"(hdr_s + 32) : "memory");
        bulk_copy_g2s(dst, g0 + from, bytes, hdr_s + 8 * q);
        sts_u32(hdr_s + 48, lds_u32(hdr_s + 48) | (16u << q));
    } else {
#pragma unroll 1
        for (uint32_t i = 0; i < chunk; i += 4) sts_u32(dst + i, 0xFFFFFFFFu);
    }
}
// Make absolute chunk `c` readable: wait for its copy, then blank whatever follows the payload inside it.
__device__ __forceinline__ void ring_land(uint32_t hdr_s, uint32_t ring_s, uint32_t chunk, uint32_t c) {
    const uint32_t q = c & 3u, st = lds_u32(hdr_s + 48);
    if (!(st & (16u << q))) return;  // nothing pending: landed earlier, or a 0xFF quarter
    mbar_wait(hdr_s + 8 * q, (st >> q) & 1u);
    sts_u32(hdr_s + 48, (st ^ (1u << q)) & ~(16u << q));
    const uint32_t end_rp = lds_u32(hdr_s + 40), lo = c * chunk, hi = lo + chunk;
    if (end_rp < hi) {
#pragma unroll 1
        for (uint32_t i = end_rp > lo ? end_rp : lo; i < hi; i++) sts_u8(ring_s + q * chunk + (i - lo), 0xFFu);
    }
}
// The read position crossed into absolute chunk c = rp / chunk (checked at symbol starts): chunk c - 1 is
// used up, so its quarter takes chunk c + 3, and chunk c + 2 must have landed before the next symbol.
// Once per 64-256 compressed bytes.  Returns the next position to call it at.
template <uint32_t RING>
__device__ __forceinline__ uint32_t ring_advance(uint32_t hdr_s, uint32_t ring_s, uint32_t rp) {
    constexpr uint32_t chunk = RING / 4;
    const uint32_t c = rp / chunk;
    ring_request(hdr_s, ring_s, chunk, c + 3);
    ring_land(hdr_s, ring_s, chunk, c + 2);
    return (c + 1) * chunk;
}

// Range decoder state of one stream, in lane 0's registers.  With 28 streams per
// SM the kernel is issue-bound (profiles/r01_decode_hybrid_ncu.txt: 85 % of issue
// slots busy), so what matters is the instruction count of a bit decode.  `bit_s`
// is written in PTX against a shared-memory byte address: 12 instructions
// (LDS, SHF, IMAD, ISETP, IADD, SEL, @IADD, SEL, IMAD, SHF, STS, SEL).
template <uint32_t RING>
struct RangeDec {
    uint32_t range, code, nextb, beyond, rp, check, end_rp;
    uint32_t ring_s, hdr_s;  // shared addresses of the ring (aligned to RING) and of its header

    // InputStream.read(): past the end of the payload it returns -1, which the reference ORs into _code as
    // all ones (RangeDecoder.java:23,36) -- so such a "byte" is 0xFFFFFFFF here, not 0xFF.
    __device__ __forceinline__ void fetch() {
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(nextb) : "r"(ring_s | (rp & (RING - 1))) : "memory");
        beyond = rp < end_rp ? 0u : 0xFFFFFFFFu;  // ORed in with the byte; kept apart so that nothing waits for the load
        rp++;
    }
    // the reader crossed into the next quarter (seen at a symbol start, Decoder.Code's loop head)
    __device__ __forceinline__ void advance() { check = ring_advance<RING>(hdr_s, ring_s, rp); }
    // RangeDecoder.java:19-25.  `p` / `n`: the payload; the mbarriers of the header at `hdr` were
    // initialised once per warp and carry their phase from stream to stream.
    __device__ __forceinline__ void init(const uint8_t* p, uint32_t n, uint32_t hdr, uint32_t ring) {
        constexpr uint32_t chunk = RING / 4;
        hdr_s = hdr;
        ring_s = ring;
        const uint64_t addr = reinterpret_cast<uint64_t>(p);
        const uint32_t skew = (uint32_t)(addr & 15u);
        asm volatile("st.shared.u64 [%0], %1;" ::"r"(hdr_s + 32), "l"(addr - skew) : "memory");
        end_rp = skew + n;
        sts_u32(hdr_s + 40, end_rp);
        sts_u32(hdr_s + 44, (end_rp + 15u) & ~15u);
#pragma unroll 1
        for (uint32_t c = 0; c < 4; c++) ring_request(hdr_s, ring_s, chunk, c);
#pragma unroll 1
        for (uint32_t c = 0; c < 3; c++) ring_land(hdr_s, ring_s, chunk, c);
        rp = skew;
        check = chunk;
        code = 0;
        range = 0xFFFFFFFFu;
        fetch();
#pragma unroll 1
        for (int i = 0; i < 5; i++) {
            code = (code << 8) | nextb | beyond;
            fetch();
        }
    }
    // the stream is over: no copy may still be in flight when the next stream reuses the ring
    __device__ __forceinline__ void drain() {
        const uint32_t c = rp / (RING / 4);
#pragma unroll 1
        for (uint32_t k = 0; k < 4; k++) ring_land(hdr_s, ring_s, RING / 4, c + k);
    }
    // RangeDecoder.java:33-37,58-62.  A real branch around the refill, in PTX: left to itself the compiler
    // predicates these eight instructions, and predicated-off instructions still take issue slots in a
    // kernel whose only active lane is issue-bound (measured: 85 -> 102 ms per C2 pass).
    __device__ __forceinline__ void normalize() {
        asm volatile(
            "{\n\t"
            ".reg .pred q, h;\n\t"
            ".reg .u32 t;\n\t"
            "setp.gt.u32 q, %0, 0xFFFFFF;\n\t"
            "@q bra.uni NORM_DONE;\n\t"
            "shl.b32 %0, %0, 8;\n\t"
            "shl.b32 %1, %1, 8;\n\t"
            "lop3.b32 %1, %1, %2, %4, 0xFE;\n\t"
            "and.b32 t, %3, %7;\n\t"
            "or.b32 t, t, %5;\n\t"
            "ld.shared.u8 %2, [t];\n\t"
            "setp.ge.u32 h, %3, %6;\n\t"
            "selp.u32 %4, 0xFFFFFFFF, 0, h;\n\t"
            "add.u32 %3, %3, 1;\n\t"
            "NORM_DONE:\n\t"
            "}"
            : "+r"(range), "+r"(code), "+r"(nextb), "+r"(rp), "+r"(beyond)
            : "r"(ring_s), "r"(end_rp), "n"(RING - 1)
            : "memory");
    }
    // RangeDecoder.DecodeBit (:43-64) on a shared-memory probability, branch-free:
    //   bit 0: p += (2048 - p) >> 5 = (31 p + 2048) >> 5      bit 1: p -= p >> 5 = (31 p + 31) >> 5
    // (p + floor(x / 32) = floor((32 p + x) / 32), and p - floor(p / 32) = ceil(31 p / 32)).
    __device__ __forceinline__ uint32_t bit_s(uint32_t saddr) {
        uint32_t b;
        asm volatile(
            "{\n\t"
            ".reg .pred z;\n\t"
            ".reg .u16 ph;\n\t"
            ".reg .u32 p, t, bound, r1, k;\n\t"
            "ld.shared.u16 ph, [%3];\n\t"
            "cvt.u32.u16 p, ph;\n\t"
            "shr.u32 t, %0, 11;\n\t"
            "mul.lo.u32 bound, t, p;\n\t"
            "setp.lt.u32 z, %1, bound;\n\t"
            "sub.u32 r1, %0, bound;\n\t"
            "selp.u32 %0, bound, r1, z;\n\t"
            "@!z sub.u32 %1, %1, bound;\n\t"
            "selp.u32 k, 2048, 31, z;\n\t"
            "mad.lo.u32 p, p, 31, k;\n\t"
            "shr.u32 p, p, 5;\n\t"
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
    // One level of a bit tree: decode the bit at shared address `a` and move to the child,
    //   a' = sbase + 2 * (2 m + bit) = 2 a - sbase + 2 bit      (a = sbase + 2 m)
    // with `nsb` = -sbase; the predicate feeds the address directly (no 0/1 materialised):
    // 18 instructions per level against 20 for `m = (m << 1) + bit_s(...)`.
    __device__ __forceinline__ void tree_step(uint32_t& a, uint32_t nsb) {
        asm volatile(
            "{\n\t"
            ".reg .pred z;\n\t"
            ".reg .u16 ph;\n\t"
            ".reg .u32 p, t, bound, r1, k, a2;\n\t"
            "ld.shared.u16 ph, [%2];\n\t"
            "cvt.u32.u16 p, ph;\n\t"
            "shr.u32 t, %0, 11;\n\t"
            "mul.lo.u32 bound, t, p;\n\t"
            "setp.lt.u32 z, %1, bound;\n\t"
            "sub.u32 r1, %0, bound;\n\t"
            "selp.u32 %0, bound, r1, z;\n\t"
            "@!z sub.u32 %1, %1, bound;\n\t"
            "selp.u32 k, 2048, 31, z;\n\t"
            "mad.lo.u32 p, p, 31, k;\n\t"
            "shr.u32 p, p, 5;\n\t"
            "cvt.u16.u32 ph, p;\n\t"
            "st.shared.u16 [%2], ph;\n\t"
            "add.u32 a2, %2, %2;\n\t"
            "add.u32 %2, a2, %3;\n\t"
            "@!z add.u32 %2, %2, 2;\n\t"
            "}"
            : "+r"(range), "+r"(code), "+r"(a)
            : "r"(nsb)
            : "memory");
        normalize();
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
    // RangeDecoder.DecodeDirectBits (:27-41).  The reference tests for normalisation after every
    // bit; a range in [2^(24+j), 2^(25+j)) falls below 2^24 exactly at its (j+1)-th halving, so the
    // bits are taken in runs of 8 - clz(range) with one test per run (same arithmetic, fewer instructions).
    __device__ __forceinline__ uint32_t direct(int nbits) {
        uint32_t result = 0;
#pragma unroll 1
        do {
            int m = 8 - __clz(range);
            if (m > nbits) m = nbits;
            nbits -= m;
#pragma unroll 1
            do {
                range >>= 1;
                // the reference tests the SIGN of code - range (:31-32); on a corrupt stream, where
                // code may exceed range by 2^31 or more, that is not the unsigned comparison
                const bool one = (int32_t)(code - range) >= 0;
                if (one) code -= range;
                result += result;
                if (one) result++;
            } while (--m);
            normalize();
        } while (nbits);
        return result;
    }
    // BitTreeDecoder.Decode (BitTreeDecoder.java:19-25); `m2` walks the tree as a byte offset (2 * m)
    template <int NBITS>
    __device__ __forceinline__ uint32_t tree(uint32_t sbase) {
        uint32_t a = sbase + 2;
        const uint32_t nsb = 0u - sbase;
#pragma unroll
        for (int i = 0; i < NBITS; i++) tree_step(a, nsb);
        return ((a + nsb) >> 1) - (1u << NBITS);
    }
    __device__ __forceinline__ uint32_t tree_n(uint32_t sbase, int nbits) {
        uint32_t a = sbase + 2;
        const uint32_t nsb = 0u - sbase;
#pragma unroll 1
        for (int i = 0; i < nbits; i++) tree_step(a, nsb);
        return ((a + nsb) >> 1) - (1u << nbits);
    }
    // BitTreeDecoder.ReverseDecode (:27-37) / Decoder.ReverseDecode (Decoder.java:13-23): a tree
    // level that also ORs `mask` (1 << level) into the symbol when the bit is 1
    __device__ __forceinline__ void rev_step(uint32_t& a, uint32_t nsb, uint32_t& symbol, uint32_t mask) {
        asm volatile(
            "{\n\t"
            ".reg .pred z;\n\t"
            ".reg .u16 ph;\n\t"
            ".reg .u32 p, t, bound, r1, k, a2;\n\t"
            "ld.shared.u16 ph, [%2];\n\t"
            "cvt.u32.u16 p, ph;\n\t"
            "shr.u32 t, %0, 11;\n\t"
            "mul.lo.u32 bound, t, p;\n\t"
            "setp.lt.u32 z, %1, bound;\n\t"
            "sub.u32 r1, %0, bound;\n\t"
            "selp.u32 %0, bound, r1, z;\n\t"
            "@!z sub.u32 %1, %1, bound;\n\t"
            "selp.u32 k, 2048, 31, z;\n\t"
            "mad.lo.u32 p, p, 31, k;\n\t"
            "shr.u32 p, p, 5;\n\t"
            "cvt.u16.u32 ph, p;\n\t"
            "st.shared.u16 [%2], ph;\n\t"
            "add.u32 a2, %2, %2;\n\t"
            "add.u32 %2, a2, %4;\n\t"
            "@!z add.u32 %2, %2, 2;\n\t"
            "@!z or.b32 %3, %3, %5;\n\t"
            "}"
            : "+r"(range), "+r"(code), "+r"(a), "+r"(symbol)
            : "r"(nsb), "r"(mask)
            : "memory");
        normalize();
    }
    __device__ __forceinline__ uint32_t reverse(uint32_t sbase, int nbits) {
        uint32_t a = sbase + 2, symbol = 0, mask = 1;
        const uint32_t nsb = 0u - sbase;
#pragma unroll 1
        for (int i = 0; i < nbits; i++, mask <<= 1) rev_step(a, nsb, symbol, mask);
        return symbol;
    }
};

// LenDecoder.Decode (Decoder.java:48-59) on the pb-strided layout; `slen` = shared address of the coder
template <class RD>
__device__ __forceinline__ uint32_t decode_len(RD& rd, uint32_t slen, int pb, uint32_t pos_state) {
    if (rd.bit_s(slen) == 0) return rd.tree<kNumLowLenBits>(slen + 2 * len_low(pb, pos_state));
    if (rd.bit_s(slen + 2) == 0) return kNumLowLenSymbols + rd.tree<kNumMidLenBits>(slen + 2 * len_mid(pb, pos_state));
    return kNumLowLenSymbols + kNumMidLenSymbols + rd.tree_n(slen + 2 * len_high(pb), kNumHighLenBits);
}

// LiteralDecoder.Decoder2.DecodeNormal (Decoder.java:70-77), unrolled
template <class RD>
__device__ __forceinline__ uint32_t decode_literal_s(RD& rd, uint32_t sprobs) {
    uint32_t a = sprobs + 2;
    const uint32_t nsb = 0u - sprobs;
#pragma unroll
    for (int i = 0; i < 8; i++) rd.tree_step(a, nsb);
    return ((a + nsb) >> 1) & 0xFF;
}
// DecodeWithMatchByte (Decoder.java:79-95): `offs` is 0x100 while the decoded bits still agree
// with the match byte (probability index ((1 + matchBit) << 8) + symbol), 0 afterwards.
template <class RD>
__device__ __forceinline__ uint32_t decode_literal_matched_s(RD& rd, uint32_t sprobs, uint32_t match_byte) {
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
// kDecHybrid: the matched tables (indices 0x100..0x2FF of a coder) are `gprobs` in global memory,
// 0x200 slots per coder; once a bit disagrees with the match byte the walk continues in the
// coder's normal tree in shared memory (0x100 slots per coder at `sprobs`).  Most matched
// literals part from their match byte within a few bits, so the top three levels of the matched
// trees (symbol < 8: 16 slots per coder at `stop`, index (matchBit << 3) + symbol) stay in shared
// memory as well and only the deeper, rarer nodes cost an L2 round trip.
template <class RD>
__device__ __forceinline__ uint32_t decode_literal_matched_h(RD& rd, uint32_t sprobs, uint32_t stop, uint16_t* gprobs,
                                                             uint32_t match_byte) {
    uint32_t symbol = 1;
#pragma unroll 1
    do {
        match_byte <<= 1;
        const uint32_t mb = match_byte & 0x100u;
        const uint32_t b = symbol < 8 ? rd.bit_s(stop + 2 * ((mb >> 5) + symbol)) : rd.bit_g(gprobs + mb + symbol);
        symbol = (symbol << 1) | b;
        if ((mb >> 8) != b) break;
    } while (symbol < 0x100);
#pragma unroll 1
    while (symbol < 0x100) symbol = (symbol << 1) | rd.bit_s(sprobs + 2 * symbol);
    return symbol & 0xFF;
}
// both forms on a generic pointer
template <class RD>
__device__ __forceinline__ uint32_t decode_literal_g(RD& rd, uint16_t* probs, bool matched, uint32_t match_byte) {
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

// DecodeArgs::progress, marks [from, to) of one stream.  Lane 0 publishes the warp's output stores at device
// scope (__syncwarp orders every lane's stores before its fence, which is cumulative) and counts the stream
// in a device-memory counter per mark; the stream that completes a mark -- it has then observed every other
// stream's count, hence their stores -- publishes system-wide and tells the host.  The host only ever starts
// a DMA read of the output after it has seen progress[m] = n.
__device__ __forceinline__ void report_progress(const DecodeArgs& a, uint32_t from, uint32_t to, int lane) {
    __syncwarp();
    if (lane == 0 && from < to) {
        __threadfence();
        uint32_t full = 0;  // marks (at most kDecMaxMarks = 32) that this stream completed for the whole batch
        for (uint32_t m = from; m < to; m++)
            if (atomicAdd(a.progress_dev + m, 1u) == a.n - 1) full |= 1u << m;
        if (full) {
            __threadfence_system();
            for (uint32_t m = from; m < to; m++)
                if (full >> m & 1u) *reinterpret_cast<volatile uint32_t*>(a.progress + m) = a.n;
        }
    }
}

template <int MODE, bool PROGRESS>
__device__ void decode_stream(const DecodeArgs& a, uint32_t s, uint16_t* model, uint16_t* lit_global, uint32_t hdr_s, uint32_t ring_s, int lane) {
    const uint64_t in_len = a.in_len[s];
    const uint8_t* in = a.in + a.in_off[s];
    uint8_t* out = a.out + a.out_off[s];
    const uint64_t cap64 = a.out_cap[s];

    // LzmaAlone.java:220-236 -- 5 property bytes + LE64 size; Decoder.java:303-318
    int status = 1;
    uint32_t pos = 0;
    uint32_t marks_done = 0;  // progress marks already reported for this stream
    uint32_t next_mark = PROGRESS && a.marks > 1 ? a.mark_at[0] : 0xFFFFFFFFu;  // output position of the next mark to report
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
        const ModelLayout L = make_layout(lc, lp, pb);
        // literal slots in shared / global memory for this mode
        const int n_lit_s = MODE == kDecSmem ? L.n_literal : MODE == kDecHybrid ? 0x110 << (lc + lp) : 0;  // hybrid: normal trees + matched tops
        const int n_lit_g = MODE == kDecSmem ? 0 : MODE == kDecHybrid ? 0x200 << (lc + lp) : L.n_literal;
        if (pb > 4 || (int32_t)dict < 0) {
            status = 0;
        } else if ((size_t)(L.n_fixed + n_lit_s) * 2 > dec_mode_model(MODE) || (MODE != kDecSmem && (size_t)n_lit_g > a.lit_stride)) {
            status = LZB_KERNEL_E_UNSUPPORTED;  // the host picked a mode this stream's lc/lp/pb does not fit
        } else {
            for (int i = lane; i < L.n_fixed + n_lit_s; i += 32) model[i] = kProbInit;  // Decoder.Init :184-203
            for (int i = lane; i < n_lit_g; i += 32) lit_global[i] = kProbInit;
            __syncwarp();

            const uint32_t dict_check = dict > 1 ? dict : 1;  // m_DictionarySizeCheck :166
            const uint32_t pos_mask = (1u << pb) - 1, lp_mask = (1u << lp) - 1;
            // outSize < 0 decodes until the end marker; a size beyond the capacity ends in EV_CAPACITY
            const uint32_t limit = usize > (uint64_t)cap ? 0xFFFFFFFFu : (uint32_t)usize;

            RangeDec<dec_mode_ring(MODE)> rd;
            const uint32_t sm = (uint32_t)__cvta_generic_to_shared(model);  // shared byte address of the model
            int state = 0;
            uint32_t rep0 = 0, rep1 = 0, rep2 = 0, rep3 = 0;
            uint32_t prev_byte = 0, match_byte = 0;
            uint32_t pend_b = 0;  // deferred tail of the last match copy (one byte per lane)
            uint8_t* pend_dst = out;
            bool pend_valid = false;
            if (lane == 0) rd.init(in + LZB_KERNEL_HEADER, (uint32_t)(in_len - LZB_KERNEL_HEADER), hdr_s, ring_s);

            for (;;) {
                uint32_t evlen = EV_DONE;  // event | len << 2
                if (lane == 0) {
                    int ev = EV_DONE;
                    uint32_t len = 0;
#pragma unroll 1
                    while (pos < limit) {  // Decoder.Code :219
                        // the one place the input ring is looked after, once per symbol: leave the symbol loop with
                        // the "refill" event (a match of length 0) so that no call sits inside it
                        if (rd.rp >= rd.check) { ev = EV_MATCH; len = 0; break; }
                        const uint32_t pos_state = pos & pos_mask;
                        if (rd.bit_s(sm + 2 * (L.is_match + (state << pb) + pos_state)) == 0) {
                            const uint32_t ctx = ((pos & lp_mask) << lc) + (prev_byte >> (8 - lc));
                            if (MODE == kDecSmem) {
                                const uint32_t sprobs = sm + 2 * (L.literal + 0x300u * ctx);
                                prev_byte = state >= 7 ? decode_literal_matched_s(rd, sprobs, match_byte) : decode_literal_s(rd, sprobs);
                            } else if (MODE == kDecHybrid) {
                                const uint32_t sprobs = sm + 2 * (L.literal + 0x100u * ctx);
                                const uint32_t stop = sm + 2 * (L.literal + (0x100u << (lc + lp)) + 16u * ctx);
                                prev_byte = state >= 7 ? decode_literal_matched_h(rd, sprobs, stop, lit_global + 0x200u * ctx, match_byte)
                                                       : decode_literal_s(rd, sprobs);
                            } else {
                                prev_byte = decode_literal_g(rd, lit_global + 0x300u * ctx, state >= 7, match_byte);
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
                    // the "a progress mark was passed" bit rides in the event word: the other lanes then need not wait
                    // for the shuffled position to find out (a shared-memory-latency stall per match otherwise)
                    if (PROGRESS && ev == EV_MATCH && len != 0 && pos >= next_mark) evlen |= 0x80000000u;
                }
                evlen = __shfl_sync(kFull, evlen, 0);
                // the tail of the previous match is still in registers: store it now that its loads
                // have landed (lane 0 decoded a whole symbol under their L2 latency)
                if (pend_valid) *pend_dst = (uint8_t)pend_b;
                pend_valid = false;
                if (evlen == 0) {  // refill event: the read position crossed into the next quarter of the input ring
                    if (lane == 0) rd.advance();
                    continue;
                }
                const int ev = (int)(evlen & 3);
                if (ev != EV_MATCH) {
                    status = ev == EV_DONE ? 1 : (ev == EV_DATA_ERROR ? 0 : LZB_KERNEL_E_CAPACITY);
                    break;
                }
                // OutWindow.CopyBlock (OutWindow.java:53-67), all lanes: out[pos+k] = out[pos-d+(k mod d)].
                const uint32_t len = (evlen >> 2) & 0x1FFu;
                const uint32_t d = __shfl_sync(kFull, rep0, 0) + 1;
                pos = __shfl_sync(kFull, pos, 0);
                if (PROGRESS && (evlen >> 31)) {
                    // everything below `pos` is stored (the pending tail went out above)
                    uint32_t reached = marks_done;
                    while (reached + 1 < a.marks && a.mark_at[reached] <= pos) reached++;
                    report_progress(a, marks_done, reached, lane);
                    marks_done = reached;
                    next_mark = marks_done + 1 < a.marks ? a.mark_at[marks_done] : 0xFFFFFFFFu;
                }
                __syncwarp();  // order earlier stores (lane 0's literals, other lanes' copies) before these loads
                const uint8_t* src = out + pos - d;
                uint8_t* dst = out + pos;
                // Lane 0 also fetches the two bytes a literal after this match would need, GetByte(0)
                // (:294) and GetByte(rep0) (:227); nothing waits for them unless that literal comes.
                // All 32-byte chunks but the last are stored at once; the last stays pending.
                uint32_t k = lane;
                if (d > len) {  // the common case: source and destination do not overlap, no modulo
                    if (lane == 0) {
                        prev_byte = src[len - 1];
                        match_byte = src[len];
                    }
                    for (uint32_t chunks = (len - 1) >> 5; chunks != 0; chunks--, k += 32) dst[k] = src[k];
                    pend_valid = k < len;
                    if (pend_valid) pend_b = src[k];
                } else {
                    if (lane == 0) {
                        prev_byte = src[(len - 1) % d];
                        match_byte = src[len % d];
                    }
                    for (uint32_t chunks = (len - 1) >> 5; chunks != 0; chunks--, k += 32) dst[k] = src[k % d];
                    pend_valid = k < len;
                    if (pend_valid) pend_b = src[k % d];
                }
                pend_dst = dst + k;
                pos += len;
            }
            if (lane == 0) rd.drain();  // no input copy may be in flight when the next stream takes the ring
        }
    }
    if (lane == 0) {
        a.out_len[s] = pos;
        a.status[s] = status;
    }
    if (PROGRESS) report_progress(a, marks_done, a.marks, lane);
}

template <int MODE, bool PROGRESS>
__global__ void __launch_bounds__(dec_mode_warps(MODE) * 32, 1) lzb_decode_kernel(DecodeArgs a) {
    // dynamic shared memory: [input rings, one per warp, each aligned to its size][ring headers][model slices]
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr uint32_t RING = dec_mode_ring(MODE);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t warps = blockDim.x >> 5;
    const uint32_t base_s = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t ring_s = base_s + (uint32_t)warp * RING;
    const uint32_t hdr_s = base_s + warps * RING + (uint32_t)warp * kInRingHeader;
    uint16_t* model = reinterpret_cast<uint16_t*>(smem_raw + warps * (RING + kInRingHeader) + (size_t)warp * dec_mode_model(MODE));
    if (lane == 0) {  // the input ring's mbarriers, once per warp (InRing)
        for (uint32_t q = 0; q < 4; q++) mbar_init(hdr_s + 8 * q);
        sts_u32(hdr_s + 48, 0);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint16_t* lit_global = MODE == kDecSmem ? nullptr
                                            : a.lit_scratch + ((size_t)blockIdx.x * (blockDim.x >> 5) + warp) * a.lit_stride;
    // first stream by position (warp-major, so a small batch spreads over all SMs), then by ticket
    const uint32_t slots = gridDim.x * (blockDim.x >> 5);
    uint32_t s = (uint32_t)warp * gridDim.x + blockIdx.x;
    while (s < a.n) {
        decode_stream<MODE, PROGRESS>(a, s, model, lit_global, hdr_s, ring_s, lane);
        if (lane == 0) s = slots + atomicAdd(a.ticket, 1u);
        s = __shfl_sync(kFull, s, 0);
    }
}

// Header pre-pass: max over well-formed streams of lc + lp + 1 (0 if there is none) and of pb + 1.
__global__ void lzb_decode_scan_headers(const uint8_t* in, const uint64_t* in_off, const uint64_t* in_len, uint32_t n,
                                        uint32_t* max_lclp1) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (in_len[i] < LZB_KERNEL_HEADER) return;
    const uint32_t v = in[in_off[i]];
    const int lc = v % 9, rem = v / 9, lp = rem % 5, pb = rem / 5;
    if (pb > 4) return;
    atomicMax(max_lclp1, (uint32_t)(lc + lp + 1));
    atomicMax(max_lclp1 + 1, (uint32_t)(pb + 1));
}

cudaError_t launch_decode_scan(const uint8_t* in, const uint64_t* in_off, const uint64_t* in_len, uint32_t n,
                               uint32_t* d_max_lclp1, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    lzb_decode_scan_headers<<<(n + 255) / 256, 256, 0, st>>>(in, in_off, in_len, n, d_max_lclp1);
    return cudaGetLastError();
}

cudaError_t launch_decode(const DecodeArgs& a, int mode, int num_sms, cudaStream_t st, int* grid_out, int* warps_out) {
    if (a.n == 0) return cudaSuccess;
    const int max_warps = dec_mode_warps(mode);
    const size_t slice = dec_mode_slice(mode);
    int warps = (int)((a.n + (uint32_t)num_sms - 1) / (uint32_t)num_sms);
    if (warps > max_warps) warps = max_warps;
    if (warps < 1) warps = 1;
    // streams are handed out by ticket, so spread the warps over every SM (4096 streams: 148 CTAs
    // of 28 warps with 48 idle warps, not 147 full CTAs and an idle SM)
    const int grid = a.n < (uint32_t)num_sms ? (int)a.n : num_sms;
    const bool prog = a.progress != nullptr;  // the reporting variant only when somebody listens
    auto kern = mode == kDecSmem     ? (prog ? lzb_decode_kernel<kDecSmem, true> : lzb_decode_kernel<kDecSmem, false>)
                : mode == kDecHybrid ? (prog ? lzb_decode_kernel<kDecHybrid, true> : lzb_decode_kernel<kDecHybrid, false>)
                                     : (prog ? lzb_decode_kernel<kDecGlobal, true> : lzb_decode_kernel<kDecGlobal, false>);
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(max_warps * slice));
    if (err != cudaSuccess) return err;
    kern<<<grid, warps * 32, (size_t)warps * slice, st>>>(a);
    if (grid_out) *grid_out = grid;
    if (warps_out) *warps_out = warps;
    return cudaGetLastError();
}

}  // namespace lzb
