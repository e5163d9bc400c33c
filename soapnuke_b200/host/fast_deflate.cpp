// fast_deflate.cpp — see fast_deflate.h. Plain RFC 1951 (deflate) / RFC 1952 (gzip) encoder written for speed.
#include "fast_deflate.h"
#include <zlib.h>            // crc32() only
#include <algorithm>
#include <cstring>
#include <vector>

namespace snk {

namespace {

constexpr int kLitLenSyms = 286, kDistSyms = 30, kClSyms = 19;
constexpr uint32_t kWindow = 32768;
constexpr int kMinMatch = 8, kMaxMatch = 258;     // short matches cost more bits than 2-bit bases and 3-bit qualities as literals
constexpr size_t kBlockTokens = 1u << 18;
constexpr int kHashBits = 15;

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193,
                                12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct Tables {
    uint8_t len_sym[256];      // match length - 3 -> length symbol - 257
    uint8_t dist_sym[512];     // distance - 1 (< 256), or 256 + ((distance - 1) >> 7)
    Tables()
    {
        for (int c = 0; c < 29; c++)
            for (int l = kLenBase[c]; l < kLenBase[c] + (1 << kLenExtra[c]) && l <= kMaxMatch; l++) len_sym[l - 3] = (uint8_t)c;
        len_sym[kMaxMatch - 3] = 28;
        for (int c = 0; c < 30; c++)
            for (uint32_t d = kDistBase[c]; d < (uint32_t)kDistBase[c] + (1u << kDistExtra[c]); d++) {
                const uint32_t i = d - 1;
                if (i < 256) dist_sym[i] = (uint8_t)c; else dist_sym[256 + (i >> 7)] = (uint8_t)c;
            }
    }
};
const Tables kT;
inline int dist_symbol(uint32_t dist) { const uint32_t i = dist - 1; return i < 256 ? kT.dist_sym[i] : kT.dist_sym[256 + (i >> 7)]; }

inline uint32_t load32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint64_t load64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

// ---- length-limited Huffman code lengths (two-queue construction, then the usual overflow repair)
void build_lengths(const uint32_t* freq, int n, int max_bits, uint8_t* lens)
{
    struct Sym { uint32_t f; int s; };
    Sym syms[kLitLenSyms];
    int m = 0;
    for (int i = 0; i < n; i++) { lens[i] = 0; if (freq[i]) syms[m++] = {freq[i], i}; }
    if (m == 0) return;
    if (m == 1) { lens[syms[0].s] = 1; return; }
    std::sort(syms, syms + m, [](const Sym& a, const Sym& b) { return a.f != b.f ? a.f < b.f : a.s < b.s; });
    uint64_t w[2 * kLitLenSyms];
    int parent[2 * kLitLenSyms], depth[2 * kLitLenSyms];
    for (int i = 0; i < m; i++) w[i] = syms[i].f;
    int li = 0, ii = m;
    for (int next = m; next < 2 * m - 1; next++) {
        int pick[2];
        for (int k = 0; k < 2; k++) pick[k] = (li < m && (ii >= next || w[li] <= w[ii])) ? li++ : ii++;
        w[next] = w[pick[0]] + w[pick[1]];
        parent[pick[0]] = parent[pick[1]] = next;
    }
    depth[2 * m - 2] = 0;
    int count[64] = {0};
    for (int k = 2 * m - 3; k >= 0; k--) {
        depth[k] = depth[parent[k]] + 1;
        if (k < m) count[std::min(depth[k], 63)]++;
    }
    for (int i = max_bits + 1; i < 64; i++) { count[max_bits] += count[i]; count[i] = 0; }
    uint64_t total = 0;
    for (int i = max_bits; i > 0; i--) total += (uint64_t)count[i] << (max_bits - i);
    while (total != (1ull << max_bits)) {
        count[max_bits]--;
        for (int i = max_bits - 1; i > 0; i--)
            if (count[i]) { count[i]--; count[i + 1] += 2; break; }
        total--;
    }
    // the least frequent symbols get the longest codes
    int k = 0;
    for (int len = max_bits; len >= 1; len--)
        for (int c = 0; c < count[len]; c++) lens[syms[k++].s] = (uint8_t)len;
}

// canonical codes, bit-reversed for the LSB-first deflate stream
void build_codes(const uint8_t* lens, int n, uint16_t* codes)
{
    int bl_count[16] = {0};
    for (int i = 0; i < n; i++) bl_count[lens[i]]++;
    bl_count[0] = 0;
    uint32_t next_code[16];
    uint32_t code = 0;
    for (int bits = 1; bits <= 15; bits++) { code = (code + (uint32_t)bl_count[bits - 1]) << 1; next_code[bits] = code; }
    for (int i = 0; i < n; i++) {
        const int len = lens[i];
        if (!len) { codes[i] = 0; continue; }
        uint32_t c = next_code[len]++, r = 0;
        for (int b = 0; b < len; b++) { r = (r << 1) | (c & 1u); c >>= 1; }
        codes[i] = (uint16_t)r;
    }
}

struct BitWriter {
    uint8_t* p;
    uint64_t acc = 0;
    int nbits = 0;
    inline void put(uint32_t v, int n)          // n <= 32, v < 2^n
    {
        acc |= (uint64_t)v << nbits;
        nbits += n;
        if (nbits >= 32) { const uint32_t lo = (uint32_t)acc; memcpy(p, &lo, 4); p += 4; acc >>= 32; nbits -= 32; }
    }
    inline void finish() { while (nbits > 0) { *p++ = (uint8_t)acc; acc >>= 8; nbits -= 8; } nbits = 0; acc = 0; }
};

struct Block {
    std::vector<uint32_t> tok;                  // literal: byte value; match: 1<<31 | (len-3) << 16 | (dist-1)
    size_t ntok = 0;
    uint32_t ll_freq[kLitLenSyms], d_freq[kDistSyms];
    void reset() { ntok = 0; memset(ll_freq, 0, sizeof ll_freq); memset(d_freq, 0, sizeof d_freq); }
};

void write_block(const Block& b, bool final, BitWriter& bw)
{
    uint32_t ll_freq[kLitLenSyms], d_freq[kDistSyms];
    memcpy(ll_freq, b.ll_freq, sizeof ll_freq); memcpy(d_freq, b.d_freq, sizeof d_freq);
    ll_freq[256] = 1;
    int used = 0;
    for (int i = 0; i < kDistSyms; i++) used += d_freq[i] != 0;
    for (int i = 0; used < 2 && i < 2; i++) if (!d_freq[i]) { d_freq[i] = 1; used++; }     // inflate wants a complete distance code
    uint8_t ll_len[kLitLenSyms], d_len[kDistSyms];
    uint16_t ll_code[kLitLenSyms], d_code[kDistSyms];
    build_lengths(ll_freq, kLitLenSyms, 15, ll_len); build_codes(ll_len, kLitLenSyms, ll_code);
    build_lengths(d_freq, kDistSyms, 15, d_len); build_codes(d_len, kDistSyms, d_code);
    int hlit = kLitLenSyms, hdist = kDistSyms;
    while (hlit > 257 && !ll_len[hlit - 1]) hlit--;
    while (hdist > 1 && !d_len[hdist - 1]) hdist--;
    // run-length code of the two length tables (symbols 16 / 17 / 18)
    uint8_t all[kLitLenSyms + kDistSyms];
    memcpy(all, ll_len, (size_t)hlit); memcpy(all + hlit, d_len, (size_t)hdist);
    const int nall = hlit + hdist;
    struct Rle { uint8_t sym, extra; };
    Rle rle[kLitLenSyms + kDistSyms];
    int nr = 0;
    uint32_t cl_freq[kClSyms] = {0};
    for (int i = 0; i < nall;) {
        const uint8_t v = all[i];
        int run = 1;
        while (i + run < nall && all[i + run] == v) run++;
        int left = run;
        if (v == 0) {
            while (left >= 11) { const int r = std::min(left, 138); rle[nr++] = {18, (uint8_t)(r - 11)}; cl_freq[18]++; left -= r; }
            if (left >= 3) { rle[nr++] = {17, (uint8_t)(left - 3)}; cl_freq[17]++; left = 0; }
            while (left-- > 0) { rle[nr++] = {0, 0}; cl_freq[0]++; }
        } else {
            rle[nr++] = {v, 0}; cl_freq[v]++; left--;
            while (left >= 3) { const int r = std::min(left, 6); rle[nr++] = {16, (uint8_t)(r - 3)}; cl_freq[16]++; left -= r; }
            while (left-- > 0) { rle[nr++] = {v, 0}; cl_freq[v]++; }
        }
        i += run;
    }
    uint8_t cl_len[kClSyms];
    uint16_t cl_code[kClSyms];
    build_lengths(cl_freq, kClSyms, 7, cl_len); build_codes(cl_len, kClSyms, cl_code);
    int hclen = kClSyms;
    while (hclen > 4 && !cl_len[kClOrder[hclen - 1]]) hclen--;
    bw.put(final ? 1u : 0u, 1); bw.put(2u, 2);
    bw.put((uint32_t)(hlit - 257), 5); bw.put((uint32_t)(hdist - 1), 5); bw.put((uint32_t)(hclen - 4), 4);
    for (int i = 0; i < hclen; i++) bw.put(cl_len[kClOrder[i]], 3);
    for (int i = 0; i < nr; i++) {
        bw.put(cl_code[rle[i].sym], cl_len[rle[i].sym]);
        if (rle[i].sym == 16) bw.put(rle[i].extra, 2);
        else if (rle[i].sym == 17) bw.put(rle[i].extra, 3);
        else if (rle[i].sym == 18) bw.put(rle[i].extra, 7);
    }
    uint32_t ll_packed[kLitLenSyms];             // code | length << 16: one load per token
    for (int i = 0; i < kLitLenSyms; i++) ll_packed[i] = ll_code[i] | ((uint32_t)ll_len[i] << 16);
    const uint32_t* tk = b.tok.data();
    for (size_t i = 0; i < b.ntok; i++) {
        const uint32_t t = tk[i];
        if (!(t & 0x80000000u)) {
            const uint32_t e = ll_packed[t];
            const uint32_t t2 = tk[i + 1];             // (one slot of slack behind the last token)
            if (i + 1 < b.ntok && !(t2 & 0x80000000u)) {   // two literals in one put (<= 30 bits)
                const uint32_t e2 = ll_packed[t2];
                bw.put((e & 0xFFFFu) | ((e2 & 0xFFFFu) << (e >> 16)), (int)((e >> 16) + (e2 >> 16)));
                i++;
                continue;
            }
            bw.put(e & 0xFFFFu, (int)(e >> 16));
            continue;
        }
        const uint32_t l3 = (t >> 16) & 0xFFu, d1 = t & 0xFFFFu;
        const int ls = kT.len_sym[l3], ds = dist_symbol(d1 + 1);
        bw.put(ll_code[257 + ls], ll_len[257 + ls]);
        if (kLenExtra[ls]) bw.put(l3 + 3 - kLenBase[ls], kLenExtra[ls]);
        bw.put(d_code[ds], d_len[ds]);
        if (kDistExtra[ds]) bw.put(d1 + 1 - kDistBase[ds], kDistExtra[ds]);
    }
    bw.put(ll_code[256], ll_len[256]);
}

} // namespace

void fast_gzip_member(const uint8_t* in, size_t n, std::string& out)
{
    const size_t start = out.size();
    // worst case: every token a literal with a 9-bit code plus the block headers
    size_t cap = start + 64 + n + n / 7 + (n / kBlockTokens + 1) * 512;
    out.resize(cap);
    uint8_t* const o0 = (uint8_t*)&out[0] + start;
    static const uint8_t hdr[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 0xff};
    memcpy(o0, hdr, 10);
    BitWriter bw;
    bw.p = o0 + 10;
    static thread_local std::vector<uint32_t> table;
    table.assign((size_t)1 << kHashBits, 0u);
    static thread_local Block blk;
    blk.reset();
    size_t pos = 0;
    blk.tok.resize(kBlockTokens + 8);
    uint32_t* const tk = blk.tok.data();
    uint32_t* const tab = table.data();          // raw pointers: the thread_local objects are not re-resolved per byte
    uint32_t* const llf = blk.ll_freq;
    uint32_t llf2[256] = {0};                    // second literal histogram: back-to-back increments of one counter would serialise
    uint32_t* const df = blk.d_freq;
    size_t nt = 0;
    auto flush = [&](bool final) {
        for (int i = 0; i < 256; i++) { llf[i] += llf2[i]; llf2[i] = 0; }
        blk.ntok = nt; write_block(blk, final, bw); blk.reset(); nt = 0;
    };
    while (pos + kMinMatch <= n) {
        const uint64_t cur = load64(in + pos);
        const uint32_t h = (uint32_t)((cur * 0x9E3779B185EBCA87ull) >> (64 - kHashBits));
        const uint32_t cand = tab[h];
        tab[h] = (uint32_t)pos;
        const uint32_t dist = (uint32_t)pos - cand;
        // cand < pos always, so the load is safe; testing the (rarely true) equality first keeps the branch predictable
        if (load64(in + cand) == cur && dist - 1u < kWindow) {
            const size_t maxlen = std::min<size_t>(kMaxMatch, n - pos);
            size_t len = 8;
            while (len + 8 <= maxlen) {
                const uint64_t x = load64(in + cand + len) ^ load64(in + pos + len);
                if (x) { len += (size_t)(__builtin_ctzll(x) >> 3); goto matched; }
                len += 8;
            }
            while (len < maxlen && in[cand + len] == in[pos + len]) len++;
        matched:
            tk[nt++] = 0x80000000u | ((uint32_t)(len - 3) << 16) | (dist - 1u);
            llf[257 + kT.len_sym[len - 3]]++;
            df[dist_symbol(dist)]++;
            pos += len;
        } else {                                   // probe every second position: two literals per miss
            const uint32_t a = in[pos], b = in[pos + 1];
            tk[nt] = a; tk[nt + 1] = b; nt += 2;
            llf[a]++; llf2[b]++;
            pos += 2;
        }
        if (nt >= kBlockTokens) flush(false);
    }
    for (; pos < n; pos++) { tk[nt++] = in[pos]; llf[in[pos]]++; }
    flush(true);
    bw.finish();
    uLong c = crc32(0L, Z_NULL, 0);
    for (size_t a = 0; a < n; a += (size_t)1 << 30) c = crc32(c, in + a, (uInt)std::min<size_t>((size_t)1 << 30, n - a));
    const uint32_t crc = (uint32_t)c, isize = (uint32_t)n;
    memcpy(bw.p, &crc, 4); memcpy(bw.p + 4, &isize, 4);
    out.resize((size_t)(bw.p + 8 - (uint8_t*)&out[0]));
}

} // namespace snk
