// text_core.cuh — per-record device functions of the FASTQ text path: the
// steps right before and right after the filter kernel, done on the device so that the host only
// moves bytes.
//
//   line index / row packing   what sub_thread's getline loop + C_fastq fill do
//                              peprocess.cpp:2090-2131 (.gz: every line loses spaceNum trailing
//                              characters), :2198-2239 (plain: every line loses 1)
//   id transform               read_filter.cpp:357-382 (index removal, only when `index` is set)
//   record formatting          peprocess.cpp:3383-3433 output_fastqs, :1617-1629 preOutput (/1 /2),
//                              seprocess.cpp:2302-2352, :919
//
// Like filter_core.cuh everything here is `__host__ __device__` so that tests/coretest can replay
// it on a CPU (test only); the product user is text_kernels.cuh.
#pragma once
#include "filter_core.cuh"

namespace snkcore {

struct TextFormat {
    int32_t strip;      // characters every input line loses at its end, the newline included
    int32_t pe_info;    // number of "/1" ("/2") suffixes preOutput appends: gp.whether_add_pe_info, twice for the clean
                        // records when the trim files are written too (peprocess.cpp:1460-1475 calls it on the same record again)
    int32_t fasta;      // gp.output_file_type == "fasta"
    int32_t id_mode;    // 0 keep, 1 index removal with seqType "0", 2 index removal otherwise
    int32_t qshift;     // outputQualityPhred - qualityPhred
};

enum TextFlags : uint32_t {
    TEXT_STRIDE_OVERFLOW = 1,   // a read is longer than the row stride: nothing was filtered, resubmit with a larger stride
    TEXT_LEN_MISMATCH = 2,      // sequence and quality lines of a record differ in length
    TEXT_LINE_COUNT = 4,        // the text does not hold 4 lines per record
    TEXT_TOO_LONG = 8           // a read exceeds SNK_MAX_READ_LEN
};

// 0x80 in every byte of w that equals '\n' (exact, no borrow artefacts)
SNK_HD uint32_t newline_bytes(uint32_t w)
{
    const uint32_t x = w ^ 0x0A0A0A0Au;
    return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);
}
// bit b set when byte b of the 16-byte chunk is '\n'; bytes at or beyond `valid` are ignored
SNK_HD uint32_t newline_mask16(const U4& v, int valid)
{
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t z = newline_bytes(w[k]);
        m |= (((z >> 7) & 1u) | ((z >> 14) & 2u) | ((z >> 21) & 4u) | ((z >> 28) & 8u)) << (4 * k);
    }
    return valid >= 16 ? m : (valid <= 0 ? 0u : (m & ((1u << valid) - 1u)));
}

// visible length of line k: its bytes up to and including the terminator, minus `strip`
SNK_HD uint32_t line_visible(const uint32_t* off, uint32_t k, uint32_t strip)
{
    const uint32_t raw = off[k + 1] - off[k];
    return raw > strip ? raw - strip : 0u;
}

// four bytes starting at any byte offset of a 4-byte aligned buffer (reads the two aligned words
// that cover them: the buffer must be readable up to 7 bytes past the last byte asked for)
SNK_HD uint32_t load4_unaligned(const uint8_t* base, size_t byte_off)
{
    const size_t a = byte_off & ~(size_t)3;
    const uint32_t lo = load4(base + a), hi = load4(base + a + 4);
    return funnel_r(lo, hi, (uint32_t)(byte_off & 3) * 8u);
}

// bytes [16c, 16c+16) of a packed row: the line's visible bytes, zero beyond `len`
SNK_HD U4 pack_chunk(const uint8_t* text, size_t line_start, uint32_t len, uint32_t c)
{
    U4 v = {0u, 0u, 0u, 0u};
    const uint32_t b0 = 16u * c;
    if (b0 >= len) return v;
    const uint32_t left = len - b0;
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t at = 4u * (uint32_t)k;
        if (at >= left) { w[k] = 0u; continue; }
        uint32_t x = load4_unaligned(text, line_start + b0 + at);
        const uint32_t nb = left - at;
        if (nb < 4u) x &= (1u << (8u * nb)) - 1u;
        w[k] = x;
    }
    v.x = w[0]; v.y = w[1]; v.z = w[2]; v.w = w[3];
    return v;
}

// read_filter.cpp:357-382 on the record's id (n visible bytes). Returns the new length; writes the
// new id to dst when dst != nullptr.
SNK_HD uint32_t id_transform(const uint8_t* id, uint32_t n, int mode, uint8_t* dst)
{
    if (mode == 1) {
        uint32_t o = 0;
        bool cp = true;
        for (uint32_t k = 0; k < n; k++) {
            const uint8_t ch = id[k];
            if (ch == '#') cp = false;
            if (cp) { if (dst) dst[o] = ch; o++; }
            else if (ch == '/') { cp = true; if (dst) dst[o] = ch; o++; }
        }
        return o;
    }
    uint32_t keep = n;
    if (mode == 2)
        for (uint32_t k = n; k > 0; k--)
            if (id[k - 1] == ':') { keep = k - 1; break; }      // substr(0, find_last_of(':')); no ':' keeps everything
    if (dst) for (uint32_t k = 0; k < keep; k++) dst[k] = id[k];
    return keep;
}

// ---- tile / fov removal lists (config keys `tile`, `fov`). stat_read parses the tile / fov out of the record
// id (read_filter.cpp:86-148) and check_tile_or_fov (:14-79) compares it with the comma separated list:
// only exact equality with an entry ever selects a read. Entries and the parsed strings are held as 8
// bytes, NUL padded, in one 64-bit word.
struct IdFilter {
    int32_t n_tile, n_fov, seq_type1, pad_;
    unsigned long long tile[SNK_MAX_ID_FILTERS], fov[SNK_MAX_ID_FILTERS];
};
SNK_HD void make_id_filter(const snk_params& p, IdFilter& F)
{
    F.n_tile = p.n_tile; F.n_fov = p.n_fov; F.seq_type1 = p.seq_type1; F.pad_ = 0;
    for (int e = 0; e < SNK_MAX_ID_FILTERS; e++) {
        unsigned long long t = 0, f = 0;
        for (int k = SNK_ID_FILTER_LEN - 1; k >= 0; k--) { t = (t << 8) | (uint8_t)p.tile[e][k]; f = (f << 8) | (uint8_t)p.fov[e][k]; }
        F.tile[e] = t; F.fov[e] = f;
    }
}
// SNK_PRE_TILE / SNK_PRE_FOV bits of one record id (n visible bytes)
SNK_HD uint32_t id_prefilter(const uint8_t* id, uint32_t n, const IdFilter& F)
{
    uint32_t flags = 0;
    if (F.n_tile > 0) {
        uint32_t i = 0;
        int num = 0;
        const int want = F.seq_type1 ? 4 : 2;
        for (; i < n; i++) {
            if (id[i] == ':') num++;
            if (num >= want) break;
        }
        unsigned long long tile = 0;
        int tn = 0;
        for (uint32_t j = 0; j != 4; j++) {
            const uint32_t k = i + j + 1;
            const uint8_t ch = k < n ? id[k] : (uint8_t)0;
            if (ch >= '0' && ch <= '9') { tile |= (unsigned long long)ch << (8 * tn); tn++; }
        }
        for (int e = 0; e < F.n_tile; e++)
            if (F.tile[e] == tile) { flags |= SNK_PRE_TILE; break; }
    }
    if (F.n_fov > 0) {
        uint32_t i = 0;
        for (; i < n; i++)
            if (id[i] == 'C' && i + 8 < n && id[i + 4] == 'R') break;
        unsigned long long fov = 0;
        for (uint32_t k = 0; k < 8 && i + k < n; k++) fov |= (unsigned long long)id[i + k] << (8 * k);      // substr(i, 8)
        for (int e = 0; e < F.n_fov; e++)
            if (F.fov[e] == fov) { flags |= SNK_PRE_FOV; break; }
    }
    return flags;
}

// bytes a kept record adds to the clean file
SNK_HD uint32_t record_out_len(uint32_t id_out, uint32_t clean_len, const TextFormat& F)
{
    const uint32_t head = id_out + 2u * (uint32_t)F.pe_info + 1u;
    return F.fasta ? head + clean_len + 1u : head + 2u * clean_len + 4u;
}

// Everything of the output record after the id: ["/1"] \n seq \n [+ \n qual \n]. Lane `lane` of
// `nl` writes bytes lane, lane+nl, ... of that tail.
SNK_HD void format_tail(uint8_t* dst, const uint8_t* seq, const uint8_t* qual, uint32_t clean_len, int mate, const TextFormat& F,
                        uint32_t lane, uint32_t nl)
{
    const uint32_t sfx = 2u * (uint32_t)F.pe_info;
    const uint32_t seq0 = sfx + 1u, seq1 = seq0 + clean_len;               // [seq0, seq1) = bases
    const uint32_t q0 = seq1 + 3u, q1 = q0 + clean_len;                    // [q0, q1) = qualities (fastq)
    const uint32_t total = F.fasta ? seq1 + 1u : q1 + 1u;
    for (uint32_t j = lane; j < total; j += nl) {
        uint8_t ch;
        if (j >= seq0 && j < seq1) ch = seq[j - seq0];
        else if (j >= q0 && j < q1) ch = (uint8_t)((int)qual[j - q0] + F.qshift);
        else if (j < sfx) ch = (j & 1u) == 0 ? (uint8_t)'/' : (uint8_t)(mate ? '2' : '1');
        else if (j == seq1 + 1u) ch = '+';
        else ch = '\n';
        dst[j] = ch;
    }
}

// output_fastqs for fasta: the first '@' of the finished id (suffix included) becomes '>'
SNK_HD void fasta_fix(uint8_t* id, uint32_t n)
{
    for (uint32_t k = 0; k < n; k++)
        if (id[k] == '@') { id[k] = '>'; return; }
}

} // namespace snkcore
