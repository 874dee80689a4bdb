// lzma_b200.hpp -- header-only C++ mirror of the reference's Encoder / Decoder classes over the
// C ABI of lzma_b200.h.  Same method names, argument meaning and error behaviour as
// SevenZip/Compression/LZMA/Encoder.java:1064-1184 and Decoder.java:205-318, with
// std::istream / std::ostream standing in for InputStream / OutputStream and
// std::runtime_error for IOException.  (The reference is Java; the build image has no JDK, so
// this is the compiled-language host side the parity tests and native callers use.  The Java
// drop-in source is under java/.)
#pragma once
#include <cstdint>
#include <istream>
#include <iterator>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "lzma_b200.h"

namespace SevenZip {

struct ICodeProgress {  // ICodeProgress.java:4
    virtual void SetProgress(int64_t inSize, int64_t outSize) = 0;
    virtual ~ICodeProgress() = default;
};

namespace Compression {
namespace LZMA {

class Encoder {
public:
    explicit Encoder(int device = 0) : h_(lzb_enc_create(device)) {
        if (!h_) throw std::runtime_error(std::string("lzb_enc_create: ") + lzb_last_error());  // no CPU fallback
    }
    ~Encoder() { lzb_enc_destroy(h_); }
    Encoder(const Encoder&) = delete;
    Encoder& operator=(const Encoder&) = delete;

    static bool SetAlgorithm(int algorithm) { return lzb_enc_set_algorithm(algorithm) == LZB_OK; }
    bool SetDictionarySize(int dictionarySize) { return check(lzb_enc_set_dictionary_size(h_, dictionarySize)); }
    bool SetNumFastBytes(int numFastBytes) { return check(lzb_enc_set_num_fast_bytes(h_, numFastBytes)); }
    bool SetMatchFinder(int matchFinderIndex) { return check(lzb_enc_set_match_finder(h_, matchFinderIndex)); }
    bool SetLcLpPb(int lc, int lp, int pb) { return check(lzb_enc_set_lc_lp_pb(h_, lc, lp, pb)); }
    void SetEndMarkerMode(bool endMarkerMode) { check(lzb_enc_set_end_marker_mode(h_, endMarkerMode ? 1 : 0)); }

    void WriteCoderProperties(std::ostream& outStream) {
        uint8_t props[5];
        check(lzb_enc_write_coder_properties(h_, props));
        outStream.write(reinterpret_cast<const char*>(props), 5);
    }

    // inSize / outSize are ignored like in the reference (Encoder.java:1046-1062)
    void Code(std::istream& inStream, std::ostream& outStream, int64_t /*inSize*/, int64_t /*outSize*/, ICodeProgress* progress) {
        std::vector<uint8_t> in((std::istreambuf_iterator<char>(inStream)), std::istreambuf_iterator<char>());
        std::vector<uint8_t> out(lzb_enc_bound(in.size()));
        uint64_t n = 0;
        const int rc = lzb_enc_code(h_, in.data(), in.size(), out.data(), out.size(), &n);
        if (rc != LZB_OK) throw std::runtime_error(std::string("lzb_enc_code: ") + lzb_last_error());
        outStream.write(reinterpret_cast<const char*>(out.data()), (std::streamsize)n);
        outStream.flush();
        if (progress) progress->SetProgress((int64_t)in.size(), (int64_t)n);
    }

private:
    static bool check(int rc) {
        if (rc < 0) throw std::runtime_error(std::string("lzma_b200: ") + lzb_last_error());
        return rc == LZB_OK;
    }
    lzb_enc* h_;
};

class Decoder {
public:
    explicit Decoder(int device = 0) : h_(lzb_dec_create(device)) {
        if (!h_) throw std::runtime_error(std::string("lzb_dec_create: ") + lzb_last_error());  // no CPU fallback
    }
    ~Decoder() { lzb_dec_destroy(h_); }
    Decoder(const Decoder&) = delete;
    Decoder& operator=(const Decoder&) = delete;

    bool SetDecoderProperties(const std::vector<uint8_t>& properties) {
        static const uint8_t none = 0;
        const int rc = lzb_dec_set_decoder_properties(h_, properties.empty() ? &none : properties.data(), (uint32_t)properties.size());
        if (rc < 0) throw std::runtime_error(std::string("lzma_b200: ") + lzb_last_error());
        return rc == LZB_OK;
    }

    // outSize < 0: decode until the end marker.  false = the reference's `return false` (corrupt data).
    bool Code(std::istream& inStream, std::ostream& outStream, int64_t outSize) {
        std::vector<uint8_t> in((std::istreambuf_iterator<char>(inStream)), std::istreambuf_iterator<char>());
        uint64_t cap = outSize >= 0 ? (uint64_t)outSize + 273 : (in.size() * 8 > (1u << 20) ? in.size() * 8 : (1u << 20));
        for (;;) {
            std::vector<uint8_t> out(cap);
            uint64_t n = 0;
            const int rc = lzb_dec_code(h_, in.data(), in.size(), out.data(), cap, outSize, &n);
            if (rc == LZB_E_CAPACITY && outSize < 0) {
                cap *= 4;
                continue;
            }
            if (rc < 0) throw std::runtime_error(std::string("lzb_dec_code: ") + lzb_last_error());
            if (rc == LZB_FALSE) return false;  // nothing is flushed (Decoder.java:281,290)
            outStream.write(reinterpret_cast<const char*>(out.data()), (std::streamsize)n);
            return true;
        }
    }

private:
    lzb_dec* h_;
};

}  // namespace LZMA
}  // namespace Compression
}  // namespace SevenZip
