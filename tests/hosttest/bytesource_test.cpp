// Test-only driver for the reader of the host pipeline (ByteSource in soapnuke_b200/host/process.cpp, compiled in here):
// bytesource_test <file> <seed>. Reads the file back through read() with random request sizes, thread counts and read-ahead
// limits - stopped before reading, or still running while read() is called - and checks the bytes, the coverage of the
// per-thread shares and their newline counts. Exit 0 = all trials identical to the file.
#include "../../soapnuke_b200/host/process.cpp"
#include <fstream>
#include <iterator>
#include <random>
using namespace snk;
int main(int argc, char** argv)
{
    if (argc < 3) return 64;
    const char* path = argv[1];
    std::ifstream f(path, std::ios::binary);
    std::vector<char> ref((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    std::mt19937 rng((unsigned)atoi(argv[2]));
    for (int trial = 0; trial < 6; trial++) {
        ByteSource src(path, false);
        src.set_read_threads(1 + (int)(rng() % 8));
        const size_t pf = (size_t)(rng() % 3 == 0 ? 0 : (1 + rng() % 64)) << 20;
        src.start_prefetch(pf, 1 + (int)(rng() % 8));
        if (trial % 2) usleep(1000 * (rng() % 40));
        if (trial % 3 != 2) src.stop_prefetch();       // trial % 3 == 2: read while the read-ahead is still running
        std::vector<char> out(ref.size() + 100);
        size_t pos = 0, nl = 0;
        std::vector<ByteSource::Part> parts;
        for (;;) {
            const size_t want = (rng() % 4 == 0) ? (1 + rng() % 100000) : ((size_t)(1 + rng() % 24) << 20);
            const size_t got = src.read(out.data() + pos, std::min(want, out.size() - pos), &parts);
            if (!got) break;
            size_t covered = 0;
            for (const ByteSource::Part& p : parts) { if (p.off != covered) { printf("shares are not contiguous\n"); return 1; } covered += p.len; nl += p.newlines; }
            if (!parts.empty() && covered != got) { printf("shares do not cover the read\n"); return 1; }
            if (parts.empty()) nl += count_newlines(out.data() + pos, got);
            pos += got;
        }
        src.stop_prefetch();
        if (pos != ref.size() || memcmp(out.data(), ref.data(), pos)) { printf("MISMATCH in trial %d at %zu bytes\n", trial, pos); return 1; }
        if (nl != count_newlines(ref.data(), ref.size())) { printf("newline count mismatch in trial %d\n", trial); return 1; }
        printf("trial %d ok: %zu MiB read ahead (limit %zu MiB)\n", trial, src.prefetched() >> 20, pf >> 20);
    }
    // nth_newline: offset just behind the need-th newline
    for (int k = 0; k < 200; k++) {
        const size_t off = rng() % (ref.size() / 2), len = 1 + rng() % 5000, need = 1 + rng() % 40;
        size_t c = 0;
        const size_t at = nth_newline(ref.data() + off, len, need, &c);
        size_t c2 = 0, at2 = len;
        for (size_t i = 0; i < len; i++) if (ref[off + i] == '\n' && ++c2 == need) { at2 = i + 1; break; }
        if (at != at2 || c != c2) { printf("nth_newline mismatch\n"); return 1; }
    }
    printf("nth_newline ok\n");
    return 0;
}
