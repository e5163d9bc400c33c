// gz_members.h — parallel decode of multi-member gzip input on host threads.
//
// The reference inflates every input file once per worker thread and keeps only the lines of its own blocks
// (peprocess.cpp:2014-2050, 2088-2131: each of the T sub_threads gzopen()s fq1/fq2 and gzgets() through the whole
// file); its unused mGzip.cpp:41-105 sketches the alternative this reader implements: a .gz file that is a
// concatenation of members (bgzip / BGZF blocks, pigz -i, `cat a.gz b.gz`, SOAPnuke's own per-thread outputs, and the
// <= 4 MiB members this engine writes) can be inflated member by member on different threads.
//
// No index is needed and nothing about the file is assumed:
//   * a scanner thread lists CANDIDATE member starts (the bytes 1f 8b 08 + a flag byte without reserved bits);
//   * workers inflate candidates speculatively, in file order, each into its own queue of decoded chunks;
//   * the consumer follows the CHAIN: the member at offset 0, then the member that starts exactly where the previous
//     one ended (zlib has checked its CRC-32 and length by then), and so on. Candidates that are not on the chain
//     (magic bytes inside compressed data) are cancelled and their output dropped, chain positions the scanner did
//     not list are queued when the chain reaches them. The decoded byte stream is therefore exactly what gzread()
//     returns for the same file, including "trailing garbage is ignored" — with one stream per member instead of one
//     per file. A single-member file degrades to one inflate stream (the reference's speed).
#ifndef SNK_GZ_MEMBERS_H
#define SNK_GZ_MEMBERS_H
#include <cstddef>
#include <cstdint>
#include <string>

namespace snk {

class GzMemberReader {
public:
    // nullptr when the file is not a regular, non-empty file that starts with a gzip header (pipes, plain text
    // under a .gz name): the caller keeps its gzread() stream for those.
    static GzMemberReader* open(const std::string& path, int nthreads);
    ~GzMemberReader();
    // Up to n decoded bytes into dst (blocks until at least one is there); 0 at the end of the data;
    // kError when the stream is corrupt or truncated (what gzread() reports as -1).
    static constexpr size_t kError = ~(size_t)0;
    size_t read(char* dst, size_t n);
    struct Counters { uint64_t members = 0, cancelled = 0, bytes_out = 0, bytes_dropped = 0; };
    Counters counters() const;
    GzMemberReader(const GzMemberReader&) = delete;
    GzMemberReader& operator=(const GzMemberReader&) = delete;
private:
    GzMemberReader() = default;
    struct Impl;
    Impl* d_ = nullptr;
};

} // namespace snk
#endif
