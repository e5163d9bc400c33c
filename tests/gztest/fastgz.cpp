// Test-only driver of snk::fast_gzip_member: fastgz <file> <piece_bytes> [repeat] -> concatenated gzip members on stdout,
// "MB/s=<rate> ratio=<r>" on stderr.
#include "../../soapnuke_b200/host/fast_deflate.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>
int main(int argc, char** argv)
{
    if (argc < 3) return 64;
    std::ifstream f(argv[1], std::ios::binary);
    std::vector<char> data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    const size_t piece = (size_t)atol(argv[2]);
    const int repeat = argc > 3 ? atoi(argv[3]) : 1;
    std::string out;
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < repeat; r++) {
        out.clear();
        if (data.empty()) snk::fast_gzip_member(nullptr, 0, out);
        for (size_t a = 0; a < data.size(); a += piece)
            snk::fast_gzip_member((const uint8_t*)data.data() + a, std::min(piece, data.size() - a), out);
    }
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    fwrite(out.data(), 1, out.size(), stdout);
    fprintf(stderr, "MB/s=%.1f ratio=%.4f\n", data.size() * (double)repeat / s / 1e6, data.empty() ? 0.0 : (double)out.size() / data.size());
    return 0;
}
