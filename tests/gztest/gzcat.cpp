// Test-only driver of snk::GzMemberReader: gzcat <file.gz> <threads> [read_size] -> decoded bytes on stdout,
// "members=<n> cancelled=<n>" on stderr; exit 2 on a decode error (after writing what was decoded before it).
#include "../../soapnuke_b200/host/gz_members.h"
#include <cstdio>
#include <cstdlib>
#include <vector>
int main(int argc, char** argv)
{
    if (argc < 3) return 64;
    snk::GzMemberReader* r = snk::GzMemberReader::open(argv[1], atoi(argv[2]));
    if (!r) { fprintf(stderr, "not-a-member-file\n"); return 3; }
    const size_t rs = argc > 3 ? (size_t)atol(argv[3]) : (1u << 20);
    std::vector<char> buf(rs);
    int rc = 0;
    for (;;) {
        const size_t got = r->read(buf.data(), rs);
        if (got == snk::GzMemberReader::kError) { rc = 2; break; }
        if (got == 0) break;
        fwrite(buf.data(), 1, got, stdout);
    }
    const auto c = r->counters();
    fprintf(stderr, "members=%llu cancelled=%llu\n", (unsigned long long)c.members, (unsigned long long)c.cancelled);
    delete r;
    return rc;
}
