// text_kernels.cuh — device side of the FASTQ text path (sm_100a): line index, row packing into the
// filter kernel's fixed-stride SoA, and compaction of the surviving records into clean FASTQ text.
//
//   newline_count_kernel   '\n' per 32 KB segment of one mate's text
//   segment_scan_kernel    exclusive scan of the segment counts (one CTA per mate)
//   line_offset_kernel     line_off[k] = byte where line k starts, k = 0..4n
//   pack_rows_kernel       lines 4i+1 / 4i+3 -> seq[i][stride] / qual[i][stride], len[i]; checks
//   (filter_kernel)        unchanged; skipped on the device when a read did not fit the stride
//   out_len_kernel         bytes each record adds to the clean file + per-CTA sums
//   block_scan_kernel      exclusive scan of the per-CTA sums (one CTA per mate)
//   out_offset_kernel      rec_off[i] = byte where record i starts in the clean text, rec_off[n] = total
//   format_kernel          one warp per kept record writes id, bases, '+', qualities
//
// All of it is byte traffic read once / written once (HBM-bound, << PCIe time of the same batch).
#pragma once
#include <cuda_runtime.h>
#include "text_core.cuh"

namespace snkcore {

constexpr uint32_t kSegThreads = 256;
constexpr uint32_t kSegIters = 8;
constexpr uint32_t kSegBytes = kSegThreads * 16u * kSegIters;      // 32 KB
constexpr uint32_t kScanThreads = 1024;
constexpr uint32_t kRecThreads = 256;                              // records per CTA in out_len / out_offset

struct TextMeta {                 // device copy of snk_text_meta (same layout)
    unsigned long long out_bytes[2];
    uint32_t kept;
    uint32_t max_len;
    uint32_t flags;
    uint32_t bad_record;
    uint32_t newlines[2];
    uint32_t pad_[2];
};

struct TextArgs {
    const uint8_t* text[2];       // 16-byte aligned, readable 32 bytes past `bytes`
    uint32_t bytes[2];
    uint32_t nseg[2];
    uint32_t* seg_count[2];       // [nseg]
    uint32_t* seg_base[2];        // [nseg]
    uint32_t* line_off[2];        // [4n + 1]
    uint8_t* seq[2];
    uint8_t* qual[2];
    uint16_t* len[2];
    const snk_read_result* res[2];
    uint32_t* rec_off[2];         // [n + 1]; holds the record lengths between out_len and out_offset
    uint32_t* blk_sum[2];         // [ceil(n / kRecThreads)]
    uint8_t* out[2];
    TextMeta* meta;
    uint32_t n;                   // records
    uint32_t stride;
    int mates;
    TextFormat fmt;
    IdFilter idf;                 // tile / fov removal lists (ids are parsed while packing)
};

#ifdef __CUDACC__

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}
// exclusive scan over the CTA (blockDim.x multiple of 32, <= 1024); *total = CTA sum. `ws` = 33 words of shared memory
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* ws, uint32_t* total)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint32_t inc = warp_incl_scan(v, lane);
    __syncthreads();                       // ws may still be read from the previous call
    if (lane == 31) ws[w] = inc;
    __syncthreads();
    if (w == 0) {
        const uint32_t s = lane < nw ? ws[lane] : 0u;
        const uint32_t si = warp_incl_scan(s, lane);
        ws[lane] = si - s;
        if (lane == 31) ws[32] = si;
    }
    __syncthreads();
    *total = ws[32];
    return ws[w] + inc - v;
}

__device__ __forceinline__ U4 seg_load(const uint8_t* text, uint32_t bytes, uint32_t off)
{
    if (off >= bytes) return U4{0u, 0u, 0u, 0u};
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(text + off));
    return U4{v.x, v.y, v.z, v.w};
}

__global__ void __launch_bounds__(kSegThreads) newline_count_kernel(const __grid_constant__ TextArgs A)
{
    const int m = blockIdx.y;
    if (blockIdx.x >= A.nseg[m]) return;
    __shared__ uint32_t ws[33];
    const uint32_t base = blockIdx.x * kSegBytes;
    uint32_t c = 0;
#pragma unroll
    for (uint32_t it = 0; it < kSegIters; it++) {
        const uint32_t off = base + it * kSegThreads * 16u + threadIdx.x * 16u;
        c += popc32(newline_mask16(seg_load(A.text[m], A.bytes[m], off), (int)min((long long)A.bytes[m] - (long long)off, 16ll)));
    }
    uint32_t total;
    block_excl_scan(c, ws, &total);
    if (threadIdx.x == 0) A.seg_count[m][blockIdx.x] = total;
}

// one CTA per mate: seg_base = exclusive scan of seg_count; meta.newlines = total
__global__ void __launch_bounds__(kScanThreads) segment_scan_kernel(const __grid_constant__ TextArgs A)
{
    const int m = blockIdx.x;
    __shared__ uint32_t ws[33];
    uint32_t carry = 0;
    for (uint32_t i0 = 0; i0 < A.nseg[m]; i0 += kScanThreads) {
        const uint32_t i = i0 + threadIdx.x;
        const uint32_t v = i < A.nseg[m] ? A.seg_count[m][i] : 0u;
        uint32_t total;
        const uint32_t ex = block_excl_scan(v, ws, &total);
        if (i < A.nseg[m]) A.seg_base[m][i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) {
        A.meta->newlines[m] = carry;
        const uint32_t want = 4u * A.n;
        // 4n lines: 4n newlines, or 4n-1 when the last line of the input has none
        if (carry == want) { /* line_off[4n] comes from the last newline */ }
        else if (carry + 1u == want && A.bytes[m] > 0) A.line_off[m][want] = A.bytes[m];
        else atomicOr(&A.meta->flags, (uint32_t)TEXT_LINE_COUNT);
        A.line_off[m][0] = 0;
    }
}

__global__ void __launch_bounds__(kSegThreads) line_offset_kernel(const __grid_constant__ TextArgs A)
{
    const int m = blockIdx.y;
    if (blockIdx.x >= A.nseg[m]) return;
    __shared__ uint32_t ws[33];
    const uint32_t base = blockIdx.x * kSegBytes;
    uint32_t k0 = A.seg_base[m][blockIdx.x];          // newlines before this segment
    const uint32_t max_line = 4u * A.n;
    uint32_t* lo = A.line_off[m];
    for (uint32_t it = 0; it < kSegIters; it++) {
        const uint32_t off = base + it * kSegThreads * 16u + threadIdx.x * 16u;
        uint32_t mask = newline_mask16(seg_load(A.text[m], A.bytes[m], off), (int)min((long long)A.bytes[m] - (long long)off, 16ll));
        uint32_t total;
        uint32_t k = k0 + block_excl_scan(popc32(mask), ws, &total);
        while (mask) {
            const int b = ctz32(mask);
            mask &= mask - 1u;
            k++;                                         // this newline ends line k-1, line k starts behind it
            if (k <= max_line) lo[k] = off + (uint32_t)b + 1u;
        }
        k0 += total;
    }
}

// thread -> (record, seq|qual, 16-byte chunk of the row)
__global__ void __launch_bounds__(256) pack_rows_kernel(const __grid_constant__ TextArgs A)
{
    const int m = blockIdx.y;
    // wrong number of lines: line_off[] is only partly written, nothing below may follow it (the batch is rejected)
    if (A.meta->flags & TEXT_LINE_COUNT) return;
    const uint32_t cpr = A.stride / 16u;                 // chunks per row
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t rec64 = g / (2u * cpr);
    if (rec64 >= A.n) return;
    const uint32_t i = (uint32_t)rec64, k = (uint32_t)(g % (2u * cpr));
    const uint32_t which = k / cpr, c = k % cpr;
    const uint32_t* off = A.line_off[m] + 4u * (size_t)i;
    const uint32_t strip = (uint32_t)A.fmt.strip;
    const uint32_t sn = line_visible(off, 1, strip), qn = line_visible(off, 3, strip);
    if (k == 0) {
        uint32_t bad = 0;
        if (sn != qn) bad |= TEXT_LEN_MISMATCH;
        if (sn > SNK_MAX_READ_LEN) bad |= TEXT_TOO_LONG;
        if (sn > A.stride) { bad |= TEXT_STRIDE_OVERFLOW; }
        atomicMax(&A.meta->max_len, sn);
        if (bad) {
            atomicOr(&A.meta->flags, bad);
            if (bad & (TEXT_LEN_MISMATCH | TEXT_TOO_LONG)) atomicMin(&A.meta->bad_record, i);
        }
        uint32_t pre = 0;           // only mate 1's id decides for a pair (sequence.cpp:213-230)
        if (m == 0 && (A.idf.n_tile > 0 || A.idf.n_fov > 0)) pre = id_prefilter(A.text[m] + off[0], line_visible(off, 0, strip), A.idf);
        A.len[m][i] = (uint16_t)((sn < A.stride ? sn : A.stride) | pre);
    }
    const uint32_t len = which ? (qn < sn ? qn : sn) : sn;
    const U4 v = pack_chunk(A.text[m], off[which ? 3 : 1], len < A.stride ? len : A.stride, c);
    uint8_t* row = (which ? A.qual[m] : A.seq[m]) + (size_t)i * A.stride;
    *reinterpret_cast<uint4*>(row + 16u * c) = make_uint4(v.x, v.y, v.z, v.w);
}

__device__ __forceinline__ uint32_t record_len_of(const TextArgs& A, int m, uint32_t i)
{
    const snk_read_result r = A.res[m][i];
    if (r.category != SNK_KEEP) return 0u;
    const uint32_t* off = A.line_off[m] + 4u * (size_t)i;
    uint32_t idn = line_visible(off, 0, (uint32_t)A.fmt.strip);
    if (A.fmt.id_mode) idn = id_transform(A.text[m] + off[0], idn, A.fmt.id_mode, nullptr);
    return record_out_len(idn, r.clean_len, A.fmt);
}

__global__ void __launch_bounds__(kRecThreads) out_len_kernel(const __grid_constant__ TextArgs A)
{
    const int m = blockIdx.y;
    __shared__ uint32_t ws[33];
    const uint32_t i = blockIdx.x * kRecThreads + threadIdx.x;
    const uint32_t v = (i < A.n && !(A.meta->flags & TEXT_STRIDE_OVERFLOW)) ? record_len_of(A, m, i) : 0u;
    if (i < A.n) A.rec_off[m][i] = v;
    uint32_t total;
    block_excl_scan(v, ws, &total);
    if (threadIdx.x == 0) A.blk_sum[m][blockIdx.x] = total;
    if (m == 0) {
        const unsigned kept = __syncthreads_count(v != 0u);
        if (threadIdx.x == 0 && kept) atomicAdd(&A.meta->kept, kept);
    }
}

// one CTA per mate: blk_sum -> exclusive scan in place; totals -> meta.out_bytes, rec_off[n]
__global__ void __launch_bounds__(kScanThreads) block_scan_kernel(const __grid_constant__ TextArgs A)
{
    const int m = blockIdx.x;
    __shared__ uint32_t ws[33];
    const uint32_t nb = (A.n + kRecThreads - 1) / kRecThreads;
    unsigned long long carry = 0;
    for (uint32_t i0 = 0; i0 < nb; i0 += kScanThreads) {
        const uint32_t i = i0 + threadIdx.x;
        const uint32_t v = i < nb ? A.blk_sum[m][i] : 0u;
        uint32_t total;
        const uint32_t ex = block_excl_scan(v, ws, &total);
        if (i < nb) A.blk_sum[m][i] = (uint32_t)carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) {
        A.meta->out_bytes[m] = carry;
        A.rec_off[m][A.n] = (uint32_t)carry;
    }
}

__global__ void __launch_bounds__(kRecThreads) out_offset_kernel(const __grid_constant__ TextArgs A)
{
    const int m = blockIdx.y;
    __shared__ uint32_t ws[33];
    const uint32_t i = blockIdx.x * kRecThreads + threadIdx.x;
    const uint32_t v = i < A.n ? A.rec_off[m][i] : 0u;
    uint32_t total;
    const uint32_t ex = block_excl_scan(v, ws, &total);
    if (i < A.n) A.rec_off[m][i] = A.blk_sum[m][blockIdx.x] + ex;
}

// one warp per (record, mate)
__global__ void __launch_bounds__(256) format_kernel(const __grid_constant__ TextArgs A)
{
    const int m = blockIdx.y;
    const uint32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31;
    if (i >= A.n) return;
    const uint32_t o0 = A.rec_off[m][i], o1 = A.rec_off[m][i + 1];
    if (o1 == o0) return;                                 // dropped
    const snk_read_result r = A.res[m][i];
    const uint32_t* off = A.line_off[m] + 4u * (size_t)i;
    const uint8_t* id = A.text[m] + off[0];
    const uint32_t idn = line_visible(off, 0, (uint32_t)A.fmt.strip);
    uint8_t* dst = A.out[m] + o0;
    uint32_t id_out = idn;
    if (A.fmt.id_mode == 0) {
        for (uint32_t j = lane; j < idn; j += 32) dst[j] = id[j];
    } else {
        if (lane == 0) id_out = id_transform(id, idn, A.fmt.id_mode, dst);
        id_out = __shfl_sync(0xFFFFFFFFu, id_out, 0);
    }
    const size_t row = (size_t)i * A.stride + r.head_cut;
    format_tail(dst + id_out, A.seq[m] + row, A.qual[m] + row, r.clean_len, m, A.fmt, lane, 32u);
    if (A.fmt.fasta) {
        __syncwarp();
        if (lane == 0) fasta_fix(dst, id_out + 2u * (uint32_t)A.fmt.pe_info);
    }
}

#endif // __CUDACC__

} // namespace snkcore
