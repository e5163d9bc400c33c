// engine.cu — host side of the CUDA engine and its extern "C" ABI (include/snk_engine.h).
//
// Owns the device statistics tables, the per-lane device staging buffers and streams, prepares the
// device parameter block (adapter budgets evaluated once with the reference's expression types,
// read_filter.cpp:714-724,769) and launches filter_kernel. There is no CPU implementation of the
// hot path in this library: without a CUDA device every compute entry point fails.
#include <cuda_runtime.h>
#include <cmath>
#include <climits>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <mutex>
#include "filter_kernel.cuh"
#include "ws_kernel.cuh"
#include <nvtx3/nvToolsExt.h>
#include "text_kernels.cuh"
#include "dev_params.h"
#include "../host/host_common.h"

using namespace snkcore;

namespace {

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            snk::set_error(std::string("CUDA error: ") + cudaGetErrorString(e__) + " at " #expr);   \
            return 1;                                                                               \
        }                                                                                           \
    } while (0)

constexpr int kLanes = 3;
constexpr size_t kSmemLimit = 227 * 1024;

struct Lane {
    cudaStream_t stream = nullptr;
    uint8_t* d_buf = nullptr;      // one allocation: seq1, qual1, seq2, qual2, len1, len2, out1, out2
    size_t cap = 0;
    // results are copied back after the kernel; remember where
    bool pending = false;
    // text path: pinned copy of the batch's TextMeta and where the outputs of the last submission live
    TextMeta* h_meta = nullptr;
    TextMeta* d_meta = nullptr;
    int text_mates = 0;
    uint32_t text_n = 0;
    const uint8_t* d_out[2] = {nullptr, nullptr};
    const uint32_t* d_rec_off[2] = {nullptr, nullptr};
    const snk_read_result* d_res[2] = {nullptr, nullptr};
    // per-stage device timing of the text path: events on the lane's stream around each stage of the last submission
    // (0 start, 1 text copied in, 2 line index + row packing done, 3 filter kernel done, 4 clean text formatted) and of
    // the last fetch (5 start, 6 copied out); folded into snk_engine::stage_ms when the lane is next synchronised
    cudaEvent_t ev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool ev_submit = false, ev_fetch = false;
};

struct LaunchPlan {
    int maxc;          // template instantiation
    uint32_t R, X, W, threads, ntiles;
    int qb;
    size_t smem;
    int grid;
};

} // namespace

struct snk_engine {
    int device = 0;
    int num_sms = 0;
    snk_params params;
    DevParams dev;
    unsigned long long* d_stats = nullptr;
    size_t stats_words = 0;
    unsigned int* d_err = nullptr;           // [0] flags
    unsigned long long* d_err_index = nullptr;
    ContamDev* d_contams = nullptr;          // [2][SNK_MAX_CONTAMS] when contaminants are configured
    GContamDev* d_gcontams = nullptr;        // [SNK_MAX_CONTAMS] when global contaminants are configured
    Lane lanes[kLanes];
    uint64_t launches = 0;
    double stage_ms[SNK_STAGE_COUNT] = {0};  // device time per stage of the text path, summed over batches and lanes
    int kernel_choice = 1;                   // 1 = filter_kernel (phase-synchronous CTAs, the production kernel), 0 = the warp-specialised
                                             // kernel where its shape fits (SNK_KERNEL=ws; measured slower)
    uint32_t ws_wpg = 0;                     // SNK_WS_WPG: force the scan group size (tuning)
    uint32_t ws_interleave = 1;              // SNK_WS_MAP=block: consecutive warps form a group (tuning)
    std::mutex mu;
    std::mutex host_mu;                      // the synchronous *_host entry points share lane 0: one caller at a time
    // per engine (= per device) launch cache, keyed by kernel instantiation: cudaFuncSetAttribute and the
    // occupancy query are per device, so they must not be remembered in function-local statics
    struct KernelCache { const void* fn; int threads; size_t smem; int blocks; bool smem_opt_in; };
    std::vector<KernelCache> kcache;
};

namespace {

struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// the lane's stream is idle: add the event intervals of its last submission / fetch to the engine's stage clocks
void fold_stage_times(snk_engine* e, Lane& L)
{
    auto ms = [&](int a, int b) { float t = 0; return cudaEventElapsedTime(&t, L.ev[a], L.ev[b]) == cudaSuccess ? (double)t : 0.0; };
    std::lock_guard<std::mutex> g(e->mu);
    if (L.ev_submit) {
        e->stage_ms[SNK_STAGE_H2D] += ms(0, 1); e->stage_ms[SNK_STAGE_INDEX_PACK] += ms(1, 2);
        e->stage_ms[SNK_STAGE_FILTER] += ms(2, 3); e->stage_ms[SNK_STAGE_FORMAT] += ms(3, 4);
        L.ev_submit = false;
    }
    if (L.ev_fetch) { e->stage_ms[SNK_STAGE_D2H] += ms(5, 6); L.ev_fetch = false; }
}

template <int MAXC, int MATES, int J>
int launch_one(snk_engine* e, const DevParams& dp, const KernelArgs& ka, const LaunchPlan& lp, cudaStream_t stream)
{
    auto kern = filter_kernel<MAXC, MATES, J>;
    snk_engine::KernelCache* kc = nullptr;
    for (auto& c : e->kcache) if (c.fn == (const void*)kern) kc = &c;
    if (!kc) { e->kcache.push_back({(const void*)kern, 0, 0, 1, false}); kc = &e->kcache.back(); }
    if (!kc->smem_opt_in) {                   // once per device and instantiation: raise the opt-in shared-memory limit
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
        kc->smem_opt_in = true;
    }
    // persistent grid: resident CTAs per SM (shared memory / registers / threads) x SMs
    if (kc->threads != (int)lp.threads || kc->smem != lp.smem) {
        int nb = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, (int)lp.threads, lp.smem));
        kc->threads = (int)lp.threads; kc->smem = lp.smem; kc->blocks = nb < 1 ? 1 : nb;
    }
    const uint32_t max_grid = (uint32_t)e->num_sms * (uint32_t)kc->blocks;
    const int grid = (int)(lp.ntiles < max_grid ? (lp.ntiles ? lp.ntiles : 1) : max_grid);
    kern<<<grid, lp.threads, lp.smem, stream>>>(dp, ka);
    CUDA_TRY(cudaGetLastError());
    e->launches++;
    return 0;
}

// ---- warp-specialised kernel (ws_kernel.cuh): one CTA per SM, rows up to kWsMaxStride bytes
template <int MAXC, int MATES>
int launch_ws_one(snk_engine* e, const DevParams& dp, const WsArgs& wa, cudaStream_t stream)
{
    auto kern = filter_ws_kernel<MAXC, MATES>;
    snk_engine::KernelCache* kc = nullptr;
    for (auto& c : e->kcache) if (c.fn == (const void*)kern) kc = &c;
    if (!kc) { e->kcache.push_back({(const void*)kern, 0, 0, 1, false}); kc = &e->kcache.back(); }
    if (!kc->smem_opt_in) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
        kc->smem_opt_in = true;
    }
    const uint32_t nt = wa.k.tm.ntiles;
    const int grid = (int)(nt < (uint32_t)e->num_sms ? (nt ? nt : 1) : (uint32_t)e->num_sms);
    kern<<<grid, kWsThreads, wa.s.total, stream>>>(dp, wa);
    CUDA_TRY(cudaGetLastError());
    e->launches++;
    return 0;
}
template <int MATES>
int launch_ws(snk_engine* e, const DevParams& dp, const WsArgs& wa, cudaStream_t stream)
{
    const uint32_t chunks = wa.k.stride / 16;
    if (chunks <= 4) return launch_ws_one<4, MATES>(e, dp, wa, stream);
    if (chunks <= 7) return launch_ws_one<7, MATES>(e, dp, wa, stream);
    if (chunks <= 10) return launch_ws_one<10, MATES>(e, dp, wa, stream);
    return launch_ws_one<16, MATES>(e, dp, wa, stream);
}

template <int MATES>
int launch_mates(snk_engine* e, const DevParams& dp, const KernelArgs& ka, const LaunchPlan& lp, cudaStream_t stream)
{
    switch (lp.maxc) {
        case 4: return launch_one<4, MATES, 4>(e, dp, ka, lp, stream);
        case 7: return launch_one<7, MATES, 4>(e, dp, ka, lp, stream);
        case 10: return launch_one<10, MATES, 4>(e, dp, ka, lp, stream);
        case 16: return launch_one<16, MATES, 4>(e, dp, ka, lp, stream);
        case 32: return launch_one<32, MATES, 4>(e, dp, ka, lp, stream);
        default: return launch_one<63, MATES, 4>(e, dp, ka, lp, stream);
    }
}

int make_plan(snk_engine* e, int mates, uint32_t stride, uint32_t n, uint64_t first, LaunchPlan& lp, TileMap& tm)
{
    if (stride == 0 || stride % 16 != 0 || stride > 1008) { snk::set_error("batch stride must be a multiple of 16 in [16,1008]"); return 1; }
    const uint32_t chunks = stride / 16;
    lp.maxc = chunks <= 4 ? 4 : chunks <= 7 ? 7 : chunks <= 10 ? 10 : chunks <= 16 ? 16 : chunks <= 32 ? 32 : 63;
    lp.W = stride / hist_j(stride);
    lp.threads = cta_threads(mates, stride);
    lp.X = align_up(hist_items(mates, stride), 32);
    int qb = e->dev.qb;
    const uint32_t maxR = lp.threads / (mates * kNT);  // kNT threads per read in phase A
    // Prefer a tile that lets two CTAs share an SM (latency hiding across the phase barriers);
    // fall back to one CTA per SM, then to fewer shared-memory quality bins.
    // Largest tile (<= one read per kNT threads) whose shared-memory plan fits; if even a small tile does
    // not fit, keep fewer quality bins in shared memory (the rest go through the checked global path).
    uint32_t R = 0;
    // two CTAs per SM hide each other's phase barriers and staging waits: accept a tile up to a quarter
    // smaller than the largest one if that is what it takes to fit twice (228 KB per SM, 1 KB reserved per CTA)
    constexpr size_t kSmemTwoPerSm = (228 * 1024) / 2 - 1024;
    for (uint32_t r = maxR; r >= 8 && 4 * r >= 3 * maxR && !R; r -= (r > 16 ? 8 : 4))
        if (plan_smem(mates, r, stride, lp.X, qb, ada_slots(e->dev.n_adapters)).total <= kSmemTwoPerSm) R = r;
    for (; !R;) {
        for (uint32_t r = maxR; r >= 8 && !R; r -= (r > 16 ? 8 : 4))
            if (plan_smem(mates, r, stride, lp.X, qb, ada_slots(e->dev.n_adapters)).total <= kSmemLimit) R = r;
        if (R) break;
        if (qb <= 0) { snk::set_error("read stride too large for the shared-memory tile"); return 1; }
        qb = qb > 4 ? qb - 4 : 0;
    }
    lp.R = R; lp.qb = qb;
    lp.smem = plan_smem(mates, R, stride, lp.X, qb, ada_slots(e->dev.n_adapters)).total;
    tm = make_tile_map(first, n, R, (uint64_t)e->params.slot_block);
    lp.ntiles = tm.ntiles;
    lp.grid = 0;       // resolved in launch_one from the kernel's real occupancy
    return 0;
}

int launch_filter(snk_engine* e, int mates, const snk_batch* d1, const snk_batch* d2, snk_read_result* o1,
                  snk_read_result* o2, uint64_t first, cudaStream_t stream, const unsigned int* skip_word = nullptr,
                  unsigned int skip_mask = 0)
{
    if (!d1 || (mates == 2 && !d2)) { snk::set_error("null batch"); return 1; }
    if (mates == 2 && (d1->n != d2->n || d1->stride != d2->stride)) { snk::set_error("reads number in fq1 and fq2 are different"); return 1; }
    if (d1->n == 0) return 0;
    if ((mates == 2) != (e->params.is_pe != 0)) { snk::set_error("engine was created for the other read layout (PE/SE)"); return 1; }
    LaunchPlan lp; TileMap tm;
    WsArgs wa;
    memset(&wa, 0, sizeof(wa));
    const bool use_ws = e->kernel_choice == 0 &&
        ws_make_shape(mates, d1->stride, e->dev.qb, ada_slots(e->dev.n_adapters), (uint32_t)kSmemLimit, e->ws_wpg, wa.s);
    if (use_ws) {
        wa.s.interleave = e->ws_interleave;
        lp.R = wa.s.R; lp.W = wa.s.W; lp.X = wa.s.X; lp.qb = wa.s.qb;
        tm = make_tile_map(first, d1->n, wa.s.R, (uint64_t)e->params.slot_block);
    } else if (make_plan(e, mates, d1->stride, d1->n, first, lp, tm)) return 1;
    KernelArgs ka;
    memset(&ka, 0, sizeof(ka));
    ka.seq[0] = d1->seq; ka.qual[0] = d1->qual; ka.len[0] = d1->len; ka.out[0] = o1;
    if (mates == 2) { ka.seq[1] = d2->seq; ka.qual[1] = d2->qual; ka.len[1] = d2->len; ka.out[1] = o2; }
    for (int m = 0; m < mates; m++)
        if (((uintptr_t)ka.seq[m] | (uintptr_t)ka.qual[m]) & 15 || ((uintptr_t)ka.out[m] & 7) || ((uintptr_t)ka.len[m] & 1)) {
            snk::set_error("device buffers must be 16-byte aligned (seq/qual) and 8-byte aligned (results)");
            return 1;
        }
    ka.stats = e->d_stats; ka.err_flags = e->d_err; ka.err_index = e->d_err_index;
    ka.skip_word = skip_word; ka.skip_mask = skip_mask;
    ka.stride = d1->stride; ka.R = lp.R; ka.items_w = lp.W; ka.X = lp.X; ka.tm = tm;
    DevParams dp = e->dev;
    dp.qb = lp.qb;               // may have been lowered so that this stride's histograms fit
    if (use_ws) {
        wa.k = ka;
        wa.magic = stride_magic(d1->stride);
        return (mates == 2) ? launch_ws<2>(e, dp, wa, stream) : launch_ws<1>(e, dp, wa, stream);
    }
    return (mates == 2) ? launch_mates<2>(e, dp, ka, lp, stream) : launch_mates<1>(e, dp, ka, lp, stream);
}

size_t batch_bytes(int mates, uint32_t n, uint32_t stride)
{
    const size_t rows = (size_t)n * stride;
    const size_t lens = ((size_t)n * 2 + 255) / 256 * 256;
    const size_t outs = ((size_t)n * sizeof(snk_read_result) + 255) / 256 * 256;
    return (size_t)mates * (2 * (rows + 256) + lens + outs);
}

int lane_reserve(Lane& L, size_t bytes)
{
    if (L.cap >= bytes) return 0;
    if (L.d_buf) CUDA_TRY(cudaFree(L.d_buf));
    L.d_buf = nullptr; L.cap = 0;
    const size_t want = bytes + bytes / 8;
    CUDA_TRY(cudaMalloc(&L.d_buf, want));
    L.cap = want;
    return 0;
}

int filter_host_async(snk_engine* e, int lane, int mates, const snk_batch* r1, const snk_batch* r2,
                      snk_read_result* out1, snk_read_result* out2, uint64_t first)
{
    if (!e) { snk::set_error("null engine"); return 1; }
    if (lane < 0 || lane >= kLanes) { snk::set_error("lane out of range"); return 1; }
    if (!r1 || !out1 || (mates == 2 && (!r2 || !out2))) { snk::set_error("null batch or result buffer"); return 1; }
    if (mates == 2 && (r1->n != r2->n || r1->stride != r2->stride)) { snk::set_error("reads number in fq1 and fq2 are different"); return 1; }
    CUDA_TRY(cudaSetDevice(e->device));
    Lane& L = e->lanes[lane];
    const uint32_t n = r1->n, stride = r1->stride;
    if (n == 0) return 0;
    CUDA_TRY(cudaStreamSynchronize(L.stream));        // the lane's buffers are free again
    if (lane_reserve(L, batch_bytes(mates, n, stride))) return 1;
    const size_t rows = (size_t)n * stride;
    const size_t rows_al = (rows + 255) / 256 * 256 + 256;
    const size_t lens = ((size_t)n * 2 + 255) / 256 * 256;
    const size_t outs = ((size_t)n * sizeof(snk_read_result) + 255) / 256 * 256;
    uint8_t* p = L.d_buf;
    snk_batch d[2];
    snk_read_result* dout[2] = {nullptr, nullptr};
    const snk_batch* h[2] = {r1, r2};
    for (int m = 0; m < mates; m++) {
        uint8_t* dseq = p; p += rows_al;
        uint8_t* dqual = p; p += rows_al;
        uint16_t* dlen = reinterpret_cast<uint16_t*>(p); p += lens;
        dout[m] = reinterpret_cast<snk_read_result*>(p); p += outs;
        CUDA_TRY(cudaMemcpyAsync(dseq, h[m]->seq, rows, cudaMemcpyHostToDevice, L.stream));
        CUDA_TRY(cudaMemcpyAsync(dqual, h[m]->qual, rows, cudaMemcpyHostToDevice, L.stream));
        CUDA_TRY(cudaMemcpyAsync(dlen, h[m]->len, (size_t)n * 2, cudaMemcpyHostToDevice, L.stream));
        d[m].seq = dseq; d[m].qual = dqual; d[m].len = dlen; d[m].n = n; d[m].stride = stride;
    }
    {
        std::lock_guard<std::mutex> g(e->mu);
        if (launch_filter(e, mates, &d[0], mates == 2 ? &d[1] : nullptr, dout[0], dout[1], first, L.stream)) return 1;
    }
    snk_read_result* hout[2] = {out1, out2};
    for (int m = 0; m < mates; m++)
        CUDA_TRY(cudaMemcpyAsync(hout[m], dout[m], (size_t)n * sizeof(snk_read_result), cudaMemcpyDeviceToHost, L.stream));
    L.pending = true;
    return 0;
}

// ------------------------------------------------------------------ FASTQ text path
__global__ void text_meta_init_kernel(TextMeta* m)
{
    m->out_bytes[0] = m->out_bytes[1] = 0;
    m->kept = 0; m->max_len = 0; m->flags = 0; m->bad_record = 0xFFFFFFFFu;
    m->newlines[0] = m->newlines[1] = 0; m->pad_[0] = m->pad_[1] = 0;
}

inline size_t al256(size_t v) { return (v + 255) / 256 * 256; }

int filter_text_async(snk_engine* e, int lane, int mates, const char* const text[2], const size_t bytes[2], uint32_t n,
                      uint32_t stride, const snk_text_format* fmt, uint64_t first)
{
    if (!e) { snk::set_error("null engine"); return 1; }
    if (lane < 0 || lane >= kLanes) { snk::set_error("lane out of range"); return 1; }
    if (!fmt || !text[0] || (mates == 2 && !text[1])) { snk::set_error("null text or format"); return 1; }
    if ((mates == 2) != (e->params.is_pe != 0)) { snk::set_error("engine was created for the other read layout (PE/SE)"); return 1; }
    if (stride == 0 || stride % 16 != 0 || stride > 1008) { snk::set_error("batch stride must be a multiple of 16 in [16,1008]"); return 1; }
    if (n == 0) { snk::set_error("empty text batch"); return 1; }
    if (fmt->strip < 0 || fmt->id_mode < 0 || fmt->id_mode > 2 || fmt->pe_info < 0 || fmt->pe_info > 2) { snk::set_error("bad text format"); return 1; }
    for (int m = 0; m < mates; m++)
        if (bytes[m] == 0 || bytes[m] > 0xF0000000ull) { snk::set_error("a text batch must hold 1 byte .. 3.75 GiB per mate"); return 1; }
    CUDA_TRY(cudaSetDevice(e->device));
    Lane& L = e->lanes[lane];
    CUDA_TRY(cudaStreamSynchronize(L.stream));        // the lane's buffers are free again
    if (!L.h_meta) {
        CUDA_TRY(cudaHostAlloc((void**)&L.h_meta, sizeof(TextMeta), cudaHostAllocDefault));
        CUDA_TRY(cudaMalloc((void**)&L.d_meta, sizeof(TextMeta)));
    }
    // ---- carve the lane buffer
    const size_t rows = al256((size_t)n * stride) + 256;
    const uint32_t nblk = (n + kRecThreads - 1) / kRecThreads;
    size_t need = 0;
    size_t o_text[2], o_segc[2], o_segb[2], o_line[2], o_seq[2], o_qual[2], o_len[2], o_res[2], o_rec[2], o_blk[2], o_out[2];
    uint32_t nseg[2] = {0, 0};
    for (int m = 0; m < mates; m++) {
        nseg[m] = (uint32_t)((bytes[m] + kSegBytes - 1) / kSegBytes);
        o_text[m] = need; need += al256(bytes[m] + 64);
        o_segc[m] = need; need += al256((size_t)nseg[m] * 4);
        o_segb[m] = need; need += al256((size_t)nseg[m] * 4);
        o_line[m] = need; need += al256(((size_t)4 * n + 2) * 4);
        o_seq[m] = need; need += rows;
        o_qual[m] = need; need += rows;
        o_len[m] = need; need += al256((size_t)n * 2);
        o_res[m] = need; need += al256((size_t)n * sizeof(snk_read_result));
        o_rec[m] = need; need += al256(((size_t)n + 1) * 4);
        o_blk[m] = need; need += al256((size_t)nblk * 4);
        // clean text <= raw text + the "/1" "/2" id suffixes (pe_info 2 = the reference's double suffix: 4 bytes per record)
        o_out[m] = need; need += al256(bytes[m] + 4 * (size_t)n + 64);
    }
    if (lane_reserve(L, need)) return 1;
    TextArgs ta;
    memset(&ta, 0, sizeof ta);
    snk_batch d[2];
    snk_read_result* dres[2] = {nullptr, nullptr};
    NvtxRange nvtx_submit("snk:text_submit");
    if (L.ev_submit || L.ev_fetch) { CUDA_TRY(cudaStreamSynchronize(L.stream)); fold_stage_times(e, L); }
    CUDA_TRY(cudaEventRecord(L.ev[0], L.stream));
    for (int m = 0; m < mates; m++) {
        uint8_t* b = L.d_buf;
        CUDA_TRY(cudaMemcpyAsync(b + o_text[m], text[m], bytes[m], cudaMemcpyHostToDevice, L.stream));
        ta.text[m] = b + o_text[m]; ta.bytes[m] = (uint32_t)bytes[m]; ta.nseg[m] = nseg[m];
        ta.seg_count[m] = reinterpret_cast<uint32_t*>(b + o_segc[m]);
        ta.seg_base[m] = reinterpret_cast<uint32_t*>(b + o_segb[m]);
        ta.line_off[m] = reinterpret_cast<uint32_t*>(b + o_line[m]);
        ta.seq[m] = b + o_seq[m]; ta.qual[m] = b + o_qual[m];
        ta.len[m] = reinterpret_cast<uint16_t*>(b + o_len[m]);
        dres[m] = reinterpret_cast<snk_read_result*>(b + o_res[m]);
        ta.res[m] = dres[m];
        ta.rec_off[m] = reinterpret_cast<uint32_t*>(b + o_rec[m]);
        ta.blk_sum[m] = reinterpret_cast<uint32_t*>(b + o_blk[m]);
        ta.out[m] = b + o_out[m];
        d[m].seq = ta.seq[m]; d[m].qual = ta.qual[m]; d[m].len = ta.len[m]; d[m].n = n; d[m].stride = stride;
        L.d_out[m] = ta.out[m]; L.d_rec_off[m] = ta.rec_off[m]; L.d_res[m] = dres[m];
    }
    ta.meta = L.d_meta; ta.n = n; ta.stride = stride; ta.mates = mates;
    make_id_filter(e->params, ta.idf);
    ta.fmt.strip = fmt->strip; ta.fmt.pe_info = fmt->pe_info; ta.fmt.fasta = fmt->fasta; ta.fmt.id_mode = fmt->id_mode;
    ta.fmt.qshift = e->params.out_quality_phred - e->params.quality_phred;
    const uint32_t seg_grid = nseg[0] > nseg[1] ? nseg[0] : nseg[1];
    const dim3 gseg(seg_grid, mates), grec(nblk, mates);
    CUDA_TRY(cudaEventRecord(L.ev[1], L.stream));
    text_meta_init_kernel<<<1, 1, 0, L.stream>>>(L.d_meta);
    newline_count_kernel<<<gseg, kSegThreads, 0, L.stream>>>(ta);
    segment_scan_kernel<<<mates, kScanThreads, 0, L.stream>>>(ta);
    line_offset_kernel<<<gseg, kSegThreads, 0, L.stream>>>(ta);
    {
        const uint64_t tasks = (uint64_t)n * 2u * (stride / 16u);
        const dim3 gpack((unsigned)((tasks + 255) / 256), mates);
        pack_rows_kernel<<<gpack, 256, 0, L.stream>>>(ta);
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(L.ev[2], L.stream));
    {
        std::lock_guard<std::mutex> g(e->mu);
        if (launch_filter(e, mates, &d[0], mates == 2 ? &d[1] : nullptr, dres[0], dres[1], first, L.stream, &L.d_meta->flags,
                          TEXT_STRIDE_OVERFLOW | TEXT_LINE_COUNT))
            return 1;
        e->launches += 9;      // the text kernels around it
    }
    CUDA_TRY(cudaEventRecord(L.ev[3], L.stream));
    out_len_kernel<<<grec, kRecThreads, 0, L.stream>>>(ta);
    block_scan_kernel<<<mates, kScanThreads, 0, L.stream>>>(ta);
    out_offset_kernel<<<grec, kRecThreads, 0, L.stream>>>(ta);
    format_kernel<<<dim3((n + 7) / 8, mates), 256, 0, L.stream>>>(ta);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(L.ev[4], L.stream));
    L.ev_submit = true;
    CUDA_TRY(cudaMemcpyAsync(L.h_meta, L.d_meta, sizeof(TextMeta), cudaMemcpyDeviceToHost, L.stream));
    L.text_mates = mates; L.text_n = n;
    L.pending = true;
    return 0;
}

} // namespace

extern "C" {

int snk_engine_create(const snk_params* p, int device, snk_engine** out)
{
    if (!p || !out) { snk::set_error("null argument"); return 1; }
    if (snk::params_check(*p)) return 1;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        snk::set_error("no CUDA device available: the filter engine has no CPU fallback");
        return 1;
    }
    if (device < 0 || device >= ndev) { snk::set_error("CUDA device index out of range"); return 1; }
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) { snk::set_error("this engine is built for sm_100a (Blackwell) only"); return 1; }
    snk_engine* e = new snk_engine();
    e->device = device;
    e->num_sms = prop.multiProcessorCount;
    if (const char* k = getenv("SNK_KERNEL")) e->kernel_choice = strcmp(k, "ws") == 0 ? 0 : 1;
    if (const char* w = getenv("SNK_WS_WPG")) e->ws_wpg = (uint32_t)atoi(w);
    if (const char* w = getenv("SNK_WS_MAP")) e->ws_interleave = strcmp(w, "block") == 0 ? 0u : 1u;
    e->params = *p;
    prepare_params(*p, e->dev);
    if (p->n_contams[0] > 0 || p->n_contams[1] > 0) {
        std::vector<ContamDev> host(2 * SNK_MAX_CONTAMS);
        prepare_contams(*p, host.data());
        if (cudaMalloc(&e->d_contams, host.size() * sizeof(ContamDev)) != cudaSuccess ||
            cudaMemcpy(e->d_contams, host.data(), host.size() * sizeof(ContamDev), cudaMemcpyHostToDevice) != cudaSuccess) {
            snk::set_error("cudaMalloc failed for the contaminant tables");
            delete e;
            return 1;
        }
        e->dev.contams = e->d_contams;
    }
    if (p->n_gcontams > 0) {
        std::vector<GContamDev> host(SNK_MAX_CONTAMS);
        prepare_gcontams(*p, host.data());
        if (cudaMalloc(&e->d_gcontams, host.size() * sizeof(GContamDev)) != cudaSuccess ||
            cudaMemcpy(e->d_gcontams, host.data(), host.size() * sizeof(GContamDev), cudaMemcpyHostToDevice) != cudaSuccess) {
            snk::set_error("cudaMalloc failed for the global contaminant table");
            delete e;
            return 1;
        }
        e->dev.gcontams = e->d_gcontams;
    }
    e->stats_words = (size_t)p->n_slots * SNK_SLOT_WORDS;
    if (cudaMalloc(&e->d_stats, e->stats_words * 8) != cudaSuccess ||
        cudaMalloc(&e->d_err, 16) != cudaSuccess || cudaMalloc(&e->d_err_index, 8) != cudaSuccess) {
        snk::set_error("cudaMalloc failed for the statistics tables");
        delete e;
        return 1;
    }
    for (int i = 0; i < kLanes; i++) {
        CUDA_TRY(cudaStreamCreateWithFlags(&e->lanes[i].stream, cudaStreamNonBlocking));
        for (cudaEvent_t& ev : e->lanes[i].ev) CUDA_TRY(cudaEventCreate(&ev));
    }
    *out = e;
    return snk_engine_stats_reset(e);
}

int snk_engine_destroy(snk_engine* e)
{
    if (!e) return 0;
    cudaSetDevice(e->device);
    for (int i = 0; i < kLanes; i++) {
        if (e->lanes[i].stream) { cudaStreamSynchronize(e->lanes[i].stream); cudaStreamDestroy(e->lanes[i].stream); }
        for (cudaEvent_t ev : e->lanes[i].ev) if (ev) cudaEventDestroy(ev);
        if (e->lanes[i].d_buf) cudaFree(e->lanes[i].d_buf);
        if (e->lanes[i].h_meta) cudaFreeHost(e->lanes[i].h_meta);
        if (e->lanes[i].d_meta) cudaFree(e->lanes[i].d_meta);
    }
    cudaFree(e->d_stats); cudaFree(e->d_err); cudaFree(e->d_err_index);
    if (e->d_contams) cudaFree(e->d_contams);
    if (e->d_gcontams) cudaFree(e->d_gcontams);
    delete e;
    return 0;
}

int snk_engine_lanes(snk_engine* e) { (void)e; return kLanes; }

int snk_filter_pe_async(snk_engine* e, int lane, const snk_batch* r1, const snk_batch* r2,
                        snk_read_result* out1, snk_read_result* out2, uint64_t first_index)
{
    return filter_host_async(e, lane, 2, r1, r2, out1, out2, first_index);
}
int snk_filter_se_async(snk_engine* e, int lane, const snk_batch* r1, snk_read_result* out1, uint64_t first_index)
{
    return filter_host_async(e, lane, 1, r1, nullptr, out1, nullptr, first_index);
}
int snk_engine_lane_sync(snk_engine* e, int lane)
{
    if (!e || lane < 0 || lane >= kLanes) { snk::set_error("bad lane"); return 1; }
    CUDA_TRY(cudaSetDevice(e->device));
    {
        NvtxRange r("snk:lane_sync");
        CUDA_TRY(cudaStreamSynchronize(e->lanes[lane].stream));
    }
    e->lanes[lane].pending = false;
    fold_stage_times(e, e->lanes[lane]);
    return 0;
}
int snk_filter_pe_host(snk_engine* e, const snk_batch* r1, const snk_batch* r2, snk_read_result* out1,
                       snk_read_result* out2, uint64_t first_index)
{
    if (!e) { snk::set_error("null engine"); return 1; }
    std::lock_guard<std::mutex> g(e->host_mu);        // callable from the reference's T worker threads at once
    if (filter_host_async(e, 0, 2, r1, r2, out1, out2, first_index)) return 1;
    return snk_engine_lane_sync(e, 0);
}
int snk_filter_se_host(snk_engine* e, const snk_batch* r1, snk_read_result* out1, uint64_t first_index)
{
    if (!e) { snk::set_error("null engine"); return 1; }
    std::lock_guard<std::mutex> g(e->host_mu);
    if (filter_host_async(e, 0, 1, r1, nullptr, out1, nullptr, first_index)) return 1;
    return snk_engine_lane_sync(e, 0);
}

int snk_filter_pe_text_async(snk_engine* e, int lane, const char* text1, size_t bytes1, const char* text2, size_t bytes2,
                             uint32_t n_records, uint32_t stride, const snk_text_format* fmt, uint64_t first_index)
{
    const char* const t[2] = {text1, text2};
    const size_t b[2] = {bytes1, bytes2};
    return filter_text_async(e, lane, 2, t, b, n_records, stride, fmt, first_index);
}
int snk_filter_se_text_async(snk_engine* e, int lane, const char* text1, size_t bytes1, uint32_t n_records, uint32_t stride,
                             const snk_text_format* fmt, uint64_t first_index)
{
    const char* const t[2] = {text1, nullptr};
    const size_t b[2] = {bytes1, 0};
    return filter_text_async(e, lane, 1, t, b, n_records, stride, fmt, first_index);
}
int snk_text_meta_sync(snk_engine* e, int lane, snk_text_meta* out)
{
    if (!e || lane < 0 || lane >= kLanes || !out) { snk::set_error("bad lane or null argument"); return 1; }
    Lane& L = e->lanes[lane];
    if (!L.h_meta || !L.text_n) { snk::set_error("no text batch was submitted on this lane"); return 1; }
    CUDA_TRY(cudaSetDevice(e->device));
    CUDA_TRY(cudaStreamSynchronize(L.stream));
    out->out_bytes[0] = L.h_meta->out_bytes[0]; out->out_bytes[1] = L.h_meta->out_bytes[1];
    out->kept = L.h_meta->kept; out->max_len = L.h_meta->max_len; out->flags = L.h_meta->flags; out->bad_record = L.h_meta->bad_record;
    return 0;
}
int snk_text_fetch_async(snk_engine* e, int lane, char* out1, char* out2, uint32_t* rec_off1, uint32_t* rec_off2,
                         snk_read_result* res1, snk_read_result* res2)
{
    if (!e || lane < 0 || lane >= kLanes) { snk::set_error("bad lane"); return 1; }
    Lane& L = e->lanes[lane];
    if (!L.h_meta || !L.text_n) { snk::set_error("no text batch was submitted on this lane"); return 1; }
    CUDA_TRY(cudaSetDevice(e->device));
    NvtxRange nvtx_fetch("snk:text_fetch");
    CUDA_TRY(cudaStreamSynchronize(L.stream));        // meta is final
    CUDA_TRY(cudaEventRecord(L.ev[5], L.stream));
    char* outs[2] = {out1, out2}; uint32_t* offs[2] = {rec_off1, rec_off2}; snk_read_result* ress[2] = {res1, res2};
    for (int m = 0; m < L.text_mates; m++) {
        if (outs[m] && L.h_meta->out_bytes[m])
            CUDA_TRY(cudaMemcpyAsync(outs[m], L.d_out[m], L.h_meta->out_bytes[m], cudaMemcpyDeviceToHost, L.stream));
        if (offs[m]) CUDA_TRY(cudaMemcpyAsync(offs[m], L.d_rec_off[m], ((size_t)L.text_n + 1) * 4, cudaMemcpyDeviceToHost, L.stream));
        if (ress[m]) CUDA_TRY(cudaMemcpyAsync(ress[m], L.d_res[m], (size_t)L.text_n * sizeof(snk_read_result), cudaMemcpyDeviceToHost, L.stream));
    }
    CUDA_TRY(cudaEventRecord(L.ev[6], L.stream));
    L.ev_fetch = true;
    return 0;
}

int snk_filter_pe_device(snk_engine* e, const snk_batch* d_r1, const snk_batch* d_r2, snk_read_result* d_out1,
                         snk_read_result* d_out2, uint64_t first_index, void* stream)
{
    if (!e) { snk::set_error("null engine"); return 1; }
    CUDA_TRY(cudaSetDevice(e->device));
    std::lock_guard<std::mutex> g(e->mu);
    return launch_filter(e, 2, d_r1, d_r2, d_out1, d_out2, first_index, (cudaStream_t)stream);
}
int snk_filter_se_device(snk_engine* e, const snk_batch* d_r1, snk_read_result* d_out1, uint64_t first_index, void* stream)
{
    if (!e) { snk::set_error("null engine"); return 1; }
    CUDA_TRY(cudaSetDevice(e->device));
    std::lock_guard<std::mutex> g(e->mu);
    return launch_filter(e, 1, d_r1, nullptr, d_out1, nullptr, first_index, (cudaStream_t)stream);
}

int snk_engine_stats_reset(snk_engine* e)
{
    if (!e) { snk::set_error("null engine"); return 1; }
    CUDA_TRY(cudaSetDevice(e->device));
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemset(e->d_stats, 0, e->stats_words * 8));
    CUDA_TRY(cudaMemset(e->d_err, 0, 16));
    CUDA_TRY(cudaMemset(e->d_err_index, 0xFF, 8));
    return 0;
}
int snk_engine_stats(snk_engine* e, uint64_t* dst)
{
    if (!e || !dst) { snk::set_error("null argument"); return 1; }
    CUDA_TRY(cudaSetDevice(e->device));
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(dst, e->d_stats, e->stats_words * 8, cudaMemcpyDeviceToHost));
    return 0;
}
int snk_engine_stats_device(snk_engine* e, uint64_t** d_ptr, size_t* words)
{
    if (!e || !d_ptr || !words) { snk::set_error("null argument"); return 1; }
    *d_ptr = reinterpret_cast<uint64_t*>(e->d_stats);
    *words = e->stats_words;
    return 0;
}
int snk_engine_stats_to_device(snk_engine* e, void* d_dst, void* stream)
{
    if (!e || !d_dst) { snk::set_error("null argument"); return 1; }
    CUDA_TRY(cudaSetDevice(e->device));
    CUDA_TRY(cudaMemcpyAsync(d_dst, e->d_stats, e->stats_words * 8, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}
int snk_engine_stats_from_device(snk_engine* e, const void* d_src, void* stream)
{
    if (!e || !d_src) { snk::set_error("null argument"); return 1; }
    CUDA_TRY(cudaSetDevice(e->device));
    CUDA_TRY(cudaMemcpyAsync(e->d_stats, d_src, e->stats_words * 8, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}
int snk_engine_error_flags(snk_engine* e, uint32_t* flags, uint64_t* first_bad_index)
{
    if (!e || !flags) { snk::set_error("null argument"); return 1; }
    CUDA_TRY(cudaSetDevice(e->device));
    CUDA_TRY(cudaDeviceSynchronize());
    unsigned int f = 0; unsigned long long idx = 0;
    CUDA_TRY(cudaMemcpy(&f, e->d_err, 4, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(&idx, e->d_err_index, 8, cudaMemcpyDeviceToHost));
    *flags = f;
    if (first_bad_index) *first_bad_index = idx;
    return 0;
}
uint64_t snk_engine_launch_count(snk_engine* e) { return e ? e->launches : 0; }
int snk_engine_stage_times(snk_engine* e, double* ms)
{
    if (!e || !ms) { snk::set_error("null argument"); return 1; }
    std::lock_guard<std::mutex> g(e->mu);
    for (int i = 0; i < SNK_STAGE_COUNT; i++) ms[i] = e->stage_ms[i];
    return 0;
}

int snk_host_alloc(void** p, size_t bytes)
{
    if (!p) { snk::set_error("null argument"); return 1; }
    CUDA_TRY(cudaHostAlloc(p, bytes, cudaHostAllocDefault));
    return 0;
}
int snk_host_free(void* p)
{
    CUDA_TRY(cudaFreeHost(p));
    return 0;
}

} // extern "C"
