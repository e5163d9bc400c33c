// fast_deflate.h — a small, fast gzip member encoder for the clean / trim FASTQ outputs.
//
// The reference writes its .gz outputs with zlib level 2 (gzopen(...,"wb") + gzsetparams(.., 2, ..),
// peprocess.cpp:1803-1810), about 80 MB/s per host thread; with 16 threads that stage, not the GPU, bounded the
// drop-in CLI at ~2 M reads/s. Only the DECOMPRESSED bytes are part of the parity contract (the reference's own member
// boundaries depend on its thread count), so the members are produced by this encoder instead: greedy LZ77 with a
// one-entry-per-bucket hash table of 4-byte prefixes (window 32 KiB, matches 4..258) and one dynamic-Huffman block per
// 256 Ki tokens, plain RFC 1951 / RFC 1952 output that any inflate reads. SNK_GZ_CODEC=zlib selects zlib level 2 again.
#ifndef SNK_FAST_DEFLATE_H
#define SNK_FAST_DEFLATE_H
#include <cstddef>
#include <cstdint>
#include <string>

namespace snk {

// appends one complete gzip member holding in[0, n) to out
void fast_gzip_member(const uint8_t* in, size_t n, std::string& out);

}
#endif
