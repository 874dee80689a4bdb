// LzmaAlone.java:190-239 ("e" / "d") written against the C++ mirror classes of include/lzma_b200.hpp.
// usage: lzma_alone e|d in out [dictLog fb]
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "lzma_b200.hpp"

using namespace SevenZip::Compression::LZMA;

int main(int argc, char** argv) {
    if (argc < 4) {
        std::fprintf(stderr, "usage: %s e|d in out [dictLog fb]\n", argv[0]);
        return 2;
    }
    try {
        std::ifstream in(argv[2], std::ios::binary);
        std::ofstream out(argv[3], std::ios::binary);
        if (argv[1][0] == 'e') {
            Encoder encoder;
            const int dict_log = argc > 4 ? std::atoi(argv[4]) : 23, fb = argc > 5 ? std::atoi(argv[5]) : 128;
            if (!Encoder::SetAlgorithm(2)) throw std::runtime_error("Incorrect compression mode");
            if (!encoder.SetDictionarySize(1 << dict_log)) throw std::runtime_error("Incorrect dictionary size");
            if (!encoder.SetNumFastBytes(fb)) throw std::runtime_error("Incorrect -fb value");
            if (!encoder.SetMatchFinder(1)) throw std::runtime_error("Incorrect -mf value");
            if (!encoder.SetLcLpPb(3, 0, 2)) throw std::runtime_error("Incorrect -lc or -lp or -pb value");
            encoder.SetEndMarkerMode(false);
            encoder.WriteCoderProperties(out);
            in.seekg(0, std::ios::end);
            const int64_t file_size = in.tellg();
            in.seekg(0);
            for (int i = 0; i < 8; i++) out.put((char)((file_size >> (8 * i)) & 0xFF));
            encoder.Code(in, out, -1, -1, nullptr);
        } else {
            char props[5];
            if (!in.read(props, 5)) throw std::runtime_error("input .lzma file is too short");
            Decoder decoder;
            if (!decoder.SetDecoderProperties(std::vector<uint8_t>(props, props + 5))) throw std::runtime_error("Incorrect stream properties");
            int64_t out_size = 0;
            for (int i = 0; i < 8; i++) {
                const int v = in.get();
                if (v < 0) throw std::runtime_error("Can't read stream size");
                out_size |= (int64_t)v << (8 * i);
            }
            if (!decoder.Code(in, out, out_size)) throw std::runtime_error("Error in data stream");
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
