// coretest.cpp — TEST-ONLY CPU replay of filter_kernel (soapnuke_b200/csrc/filter_kernel.cuh).
//
// The product runs the functions of filter_core.cuh inside a CUDA kernel. This file compiles the
// SAME header as plain C++ and walks the same tile -> phase A -> phase P -> phase B -> flush
// structure sequentially, so that the per-read / per-position device logic can be checked against
// the oracle in the CPU-only test tier (no GPU in the build container). It is never linked into the
// product and is not a fallback: it exists only under tests/.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include "../../soapnuke_b200/csrc/filter_kernel.cuh"
#include "../../soapnuke_b200/csrc/dev_params.h"
#include "../../soapnuke_b200/csrc/text_core.cuh"

using namespace snkcore;

namespace {

struct Ctx {
    DevParams P;
    uint64_t* stats;
    uint32_t err = 0;
    uint32_t stride, R, W, X;
    int mates, J;
    std::vector<QCounter> qhist;
    std::vector<BaseCnt<4>> bc;         // one per b-unit (per "thread")
    uint32_t threads = 0;               // CTA size the kernel would use (decides one or two units per item)
    uint64_t lastkey[8] = {0};
};

int file_of(int mates, int tab) { return mates == 2 ? tab : (tab == 0 ? SNK_RAW1 : SNK_CLEAN1); }

void flush(Ctx& c, int slot)
{
    uint64_t* S = c.stats + (size_t)slot * SNK_SLOT_WORDS;
    const int ntab = 2 * c.mates;
    const uint32_t J = (uint32_t)c.J;
    const uint32_t nraw = (uint32_t)c.mates * c.W;
    // raw cells count all records, delta cells "removed - added": raw += raw, clean += raw - delta
    for (uint32_t qj = 0; qj < (uint32_t)c.P.qb * J; qj++)
        for (uint32_t x = 0; x < nraw; x++) {
            const uint32_t j = qj % J, q = qj / J;
            const uint32_t e = qcell_index<4>(q, j, x, c.X), ed = qcell_index<4>(q, j, x + nraw, c.X);
            const uint32_t vr = c.qhist[e];
            const int vd = (int)(int16_t)c.qhist[ed];
            c.qhist[e] = 0; c.qhist[ed] = 0;
            const uint32_t tab = x / c.W, w = x % c.W;
            const uint64_t vc = (uint64_t)((int64_t)vr - (int64_t)vd);
            uint64_t* FR = S + SNK_SLOT_FILE_OFF(file_of(c.mates, tab));
            uint64_t* FC = S + SNK_SLOT_FILE_OFF(file_of(c.mates, tab + c.mates));
            const size_t cell = SNK_FILE_QS_OFF + (size_t)(J * w + j) * SNK_QBINS + q;
            FR[cell] += vr; FC[cell] += vc;
            if (q >= 20) { FR[SNK_FILE_GS_OFF + SNK_GS_Q20] += vr; FC[SNK_FILE_GS_OFF + SNK_GS_Q20] += vc; }
            if (q >= 30) { FR[SNK_FILE_GS_OFF + SNK_GS_Q30] += vr; FC[SNK_FILE_GS_OFF + SNK_GS_Q30] += vc; }
        }
    for (size_t u = 0; u < c.bc.size(); u++) {
        const uint32_t x = (uint32_t)(u % nraw);            // unit u counts for raw item x (two units per item when wide)
        const uint32_t tab = x / c.W, w = x % c.W;
        uint64_t* FR = S + SNK_SLOT_FILE_OFF(file_of(c.mates, tab));
        uint64_t* FC = S + SNK_SLOT_FILE_OFF(file_of(c.mates, tab + c.mates));
        for (int b = 0; b < 5; b++)
            for (int j = 0; j < c.J; j++) {
                const uint32_t vr = (c.bc[u].raw[b][j / 2] >> (16 * (j & 1))) & 0xFFFFu;
                const int vd = (int)((c.bc[u].del[b][j / 2] >> (16 * (j & 1))) & 0xFFFFu) - 0x8000;
                const uint64_t vc = (uint64_t)((int64_t)vr - (int64_t)vd);
                const size_t cell = SNK_FILE_BS_OFF + (size_t)(J * w + j) * 5 + b;
                FR[cell] += vr; FR[SNK_FILE_GS_OFF + SNK_GS_A + b] += vr; FR[SNK_FILE_GS_OFF + SNK_GS_BASES] += vr;
                FC[cell] += vc; FC[SNK_FILE_GS_OFF + SNK_GS_A + b] += vc; FC[SNK_FILE_GS_OFF + SNK_GS_BASES] += vc;
            }
        base_cnt_reset<4>(c.bc[u]);
    }
    for (int t = 0; t < ntab; t++) {
        uint64_t* G = S + SNK_SLOT_FILE_OFF(file_of(c.mates, t)) + SNK_FILE_GS_OFF;
        if (c.lastkey[t] > G[SNK_GS_LAST_KEY]) G[SNK_GS_LAST_KEY] = c.lastkey[t];
        G[SNK_GS_READS] += c.lastkey[4 + t];
        c.lastkey[t] = 0; c.lastkey[4 + t] = 0;
    }
}

template <int MAXC, int J>
void run(Ctx& c, const snk_batch* b[2], snk_read_result* out[2], uint64_t first, int grid, uint32_t flush_every)
{
    const int M = c.mates;
    const uint32_t n = b[0]->n;
    TileMap tm = make_tile_map(first, n, c.R, (uint64_t)c.P.slot_block);
    std::vector<uint8_t> rows[2][2];
    std::vector<ReadInfo> info[2];
    std::vector<uint8_t> keep(c.R);
    std::vector<DeltaEnt> dlist[2];
    const int nchunks = (int)(c.stride / 16);
    for (int m = 0; m < M; m++) { rows[m][0].assign((size_t)(c.R + 1) * c.stride + 16, 0xAB); rows[m][1].assign((size_t)(c.R + 1) * c.stride + 16, 0xAB); info[m].resize(c.R); }   // (+ one row: unit_q_raw loads one record ahead)
    c.qhist.assign((size_t)(c.P.qb + 1) * (size_t)J * c.X, 0);
    const uint32_t nraw = (uint32_t)M * c.W;
    const bool wide = c.threads >= 4u * nraw;               // same rule as the kernel
    c.bc.assign(wide ? 2u * nraw : nraw, BaseCnt<4>());
    for (auto& x : c.bc) base_cnt_reset<4>(x);
    for (int cta = 0; cta < grid; cta++) {
        const uint32_t t_begin = (uint32_t)((uint64_t)tm.ntiles * cta / grid), t_end = (uint32_t)((uint64_t)tm.ntiles * (cta + 1) / grid);
        int cur_slot = -1;
        uint32_t reads_in_hist = 0;
        for (uint32_t t = t_begin; t < t_end; t++) {
            uint32_t start, cnt;
            tile_range(tm, t, &start, &cnt);
            const uint64_t g0 = first + start;
            const int slot = slot_of(g0, (uint64_t)c.P.slot_block, c.P.n_slots);
            if (slot != cur_slot || reads_in_hist + cnt > flush_every) { if (cur_slot >= 0) flush(c, cur_slot); cur_slot = slot; reads_in_hist = 0; }
            reads_in_hist += cnt;
            uint64_t* S = c.stats + (size_t)slot * SNK_SLOT_WORDS;
            for (int m = 0; m < M; m++) {
                memcpy(rows[m][0].data(), b[m]->seq + (size_t)start * c.stride, (size_t)cnt * c.stride);
                memcpy(rows[m][1].data(), b[m]->qual + (size_t)start * c.stride, (size_t)cnt * c.stride);
            }
            // phase A
            bool tile_slow = false;
            for (int m = 0; m < M; m++)
                for (uint32_t r = 0; r < cnt; r++) {
                    const uint32_t len_word = b[m]->len[start + r];
                    int len = (int)(len_word & SNK_LEN_MASK);
                    if (len > (int)c.stride) len = (int)c.stride;
                    ReadInfo ri;
                    if (len <= 0) { memset(&ri, 0, sizeof ri); ri.head_hdcut = ri.head_lqcut = ri.tail_hdcut = ri.tail_lqcut = ri.adacut_pos = -1; ri.flags = RF_BAD_BASE | RF_QSLOW; }
                    else scan_read_serial<MAXC>(rows[m][0].data() + (size_t)r * c.stride, rows[m][1].data() + (size_t)r * c.stride, len, nchunks, m, c.P, ri);
                    ri.flags |= pre_flags(len_word);
                    info[m][r] = ri;
                    if (ri.flags & RF_QSLOW) tile_slow = true;
                }
            // phase P
            for (int m = 0; m < M; m++) dlist[m].clear();
            for (uint32_t r = 0; r < cnt; r++) {
                const uint64_t gi = g0 + r;
                int cat, mask = 0, fsb = -1;
                const ReadInfo& a = info[0][r];
                const ReadInfo& bb = info[M - 1][r];
                if (M == 2) {
                    cat = decide_pair(c.P, a, bb, &mask, &fsb);
                    if ((a.flags | bb.flags) & RF_BAD_BASE) c.err |= ERR_BAD_BASE;
                    if ((a.flags | bb.flags) & RF_BAD_QUAL) c.err |= ERR_BAD_QUAL;
                    if (cat == SNK_DROP_LOWQ && ((a.flags | bb.flags) & RF_LOWQ_GT1)) c.err |= ERR_LOWQ_RATIO;
                } else {
                    cat = c.P.srna ? decide_srna(c.P, a, &fsb) : decide_se(c.P, a, &fsb);
                    mask = cat ? 1 : 0;
                    if (a.flags & RF_BAD_BASE) c.err |= ERR_BAD_BASE;
                    if (a.flags & RF_BAD_QUAL) c.err |= ERR_BAD_QUAL;
                }
                keep[r] = cat == SNK_KEEP;
                for (int m = 0; m < M; m++) {
                    DeltaEnt de[2];
                    const int nde = delta_entries(info[m][r], cat == SNK_KEEP, r * c.stride, de);
                    for (int i = 0; i < nde; i++) dlist[m].push_back(de[i]);
                }
                if (fsb >= 0) {
                    S[fsb]++;
                    if (M == 2) { if (mask & 1) S[fsb + 1]++; if (mask & 2) S[fsb + 2]++; if (mask == 3) S[fsb + 3]++; }
                }
                for (int m = 0; m < M; m++) {
                    const ReadInfo& x = info[m][r];
                    snk_read_result res;
                    res.head_cut = (uint16_t)x.head_cut; res.clean_len = (uint16_t)x.clean_len;
                    res.category = (uint8_t)cat; res.mate_mask = (uint8_t)mask; res.adacut_pos = x.adacut_pos;
                    out[m][start + r] = res;
                    const int which = M == 2 ? m : 2;
                    int hf, tf;
                    if (c.P.cutback) {
                        trim_stat_indices(which, x.len, 0, x.head_hdcut, x.head_lqcut, x.tail_hdcut, x.tail_lqcut, x.adacut_pos, &hf, &tf);
                        uint64_t* T = S + SNK_SLOT_FILE_OFF(m == 0 ? SNK_RAW1 : SNK_RAW2) + SNK_FILE_TS_OFF;
                        if (hf >= 0) T[hf]++;
                        if (tf >= 0) T[tf]++;
                    }
                    if (cat == SNK_KEEP) {
                        trim_stat_indices(which, x.clean_len, x.len, x.head_hdcut, x.head_lqcut, x.tail_hdcut, x.tail_lqcut, x.adacut_pos, &hf, &tf);
                        uint64_t* T = S + SNK_SLOT_FILE_OFF(m == 0 ? SNK_CLEAN1 : SNK_CLEAN2) + SNK_FILE_TS_OFF;
                        if (hf >= 0) T[hf]++;
                        if (tf >= 0) T[tf]++;
                    }
                    const uint64_t kraw = ((gi + 1) << 16) | (uint64_t)(uint16_t)x.len;
                    if (kraw > c.lastkey[m]) c.lastkey[m] = kraw;
                    c.lastkey[4 + m]++;
                    if (cat == SNK_KEEP) {
                        const uint64_t kc = ((gi + 1) << 16) | (uint64_t)(uint16_t)x.clean_len;
                        if (kc > c.lastkey[M + m]) c.lastkey[M + m] = kc;
                        c.lastkey[4 + M + m]++;
                    }
                }
            }
            // phase B: the kernel's work units, one after the other (q-units then b-units)
            std::vector<uint32_t> rawdesc[2];
            for (int m = 0; m < M; m++) {
                rawdesc[m].resize(cnt);
                for (uint32_t r = 0; r < cnt; r++) rawdesc[m][r] = hist_desc(info[m][r].len, r * c.stride, info[m][r].flags & RF_QSLOW);
            }
            const uint32_t n_units = wide ? 2u * nraw : nraw;
            const int q_jstep = (int)c.X * 2 * (int)sizeof(QCounter), q_bstep = (J / 2) * q_jstep;      // rowstep, bstep (filter_core.cuh)
            for (uint32_t u = 0; u < n_units; u++) {
                const uint32_t x = wide ? u % nraw : u, half = wide ? u / nraw : 0u;
                const uint32_t w = x % c.W;
                const int m = (int)(x / c.W) % M;
                unsigned long long* f_raw = (unsigned long long*)(S + SNK_SLOT_FILE_OFF(file_of(M, m)));
                unsigned long long* f_clean = (unsigned long long*)(S + SNK_SLOT_FILE_OFF(file_of(M, m + M)));
                const int q_cell0 = (int)x * 2 * (int)sizeof(QCounter) - c.P.phred * q_bstep;
                const int q_cell0_del = q_cell0 + (int)nraw * 2 * (int)sizeof(QCounter);
                const uint8_t* rs = rows[m][0].data();
                const uint8_t* rq = rows[m][1].data();
                const DeltaEnt* dl = dlist[m].data();
                const uint32_t nd = (uint32_t)dlist[m].size();
                if (!tile_slow) {
                    if (wide) {
                        unit_q_raw<QCounter, J, J / 2>(rq, c.stride, cnt, (int)w, (int)half * (J / 2), (uint8_t*)c.qhist.data(), q_cell0, q_jstep, q_bstep);
                        unit_q_delta<QCounter, J, J / 2>(rq, dl, nd, (int)w, (int)half * (J / 2), (uint8_t*)c.qhist.data(), q_cell0_del, q_jstep, q_bstep);
                    } else {
                        unit_q_raw<QCounter, J, J>(rq, c.stride, cnt, (int)w, 0, (uint8_t*)c.qhist.data(), q_cell0, q_jstep, q_bstep);
                        unit_q_delta<QCounter, J, J>(rq, dl, nd, (int)w, 0, (uint8_t*)c.qhist.data(), q_cell0_del, q_jstep, q_bstep);
                    }
                    unit_b_raw<J>(rs, c.stride, cnt, (int)w, wide ? half : 0u, wide ? 2u : 1u, c.bc[u]);
                    unit_b_delta<J>(rs, dl, nd, (int)w, wide ? half : 0u, wide ? 2u : 1u, c.bc[u]);
                } else {
                    c.err |= unit_q_checked<QCounter, J>(rq, rawdesc[m].data(), cnt, dl, nd, (int)w, wide ? (int)half * (J / 2) : 0, wide ? J / 2 : J,
                                                         c.P.phred, c.P.qb, c.qhist.data() + 2u * x, nraw, (int)c.X, f_raw, f_clean);
                    unit_b_checked<J>(rs, rawdesc[m].data(), cnt, dl, nd, (int)w, wide ? half : 0u, wide ? 2u : 1u, c.bc[u]);
                }
            }
        }
        if (cur_slot >= 0) flush(c, cur_slot);
    }
}


// ---- replay of the warp-specialised kernel (ws_kernel.cuh): same tiles (R = the ws shape's tile), the same phase A /
// phase P per read, indicator planes stored the way the scan warps do, and the histogram ITEMS of ws_core.cuh (q-items
// on the owner-computes cells, b-items on vertical counters in a shared-memory-like array), flushed item by item.
template <int MAXC>
void run_ws(Ctx& c, const WsShape& S, const snk_batch* b[2], snk_read_result* out[2], uint64_t first, int grid, uint32_t flush_every)
{
    constexpr int J = kWsJ;
    const int M = c.mates;
    const uint32_t n = b[0]->n;
    TileMap tm = make_tile_map(first, n, S.R, (uint64_t)c.P.slot_block);
    std::vector<uint8_t> rows[2][2];
    std::vector<uint32_t> ind[2], desc((size_t)M * S.R);
    std::vector<DeltaEnt> dlist[2];
    std::vector<ReadInfo> info[2];
    const int nchunks = (int)(c.stride / 16);
    for (int m = 0; m < M; m++) {
        rows[m][0].assign((size_t)(S.R + 1) * c.stride + 16, 0xAB); rows[m][1].assign((size_t)(S.R + 1) * c.stride + 16, 0xAB);
        ind[m].assign(ind_words((int)S.nwd, S.rp), 0xDEADBEEFu);
        info[m].resize(S.R);
    }
    std::vector<QCounter> qhist((size_t)(c.P.qb + 1) * J * S.X, 0);
    std::vector<uint32_t> bstate(bstate_words(S.nb_pitch), 0);
    const uint32_t nraw = S.nq, magic = stride_magic(c.stride);
    const int q_jstep = (int)S.X * 2 * (int)sizeof(QCounter), q_bstep = (J / 2) * q_jstep;
    auto flush_ws = [&](int slot) {
        uint64_t* Sl = c.stats + (size_t)slot * SNK_SLOT_WORDS;
        for (uint32_t it = 0; it < S.nq + S.nb; it++) {
            if (it < S.nq) {
                const int mm = (int)(it / S.W), w = (int)(it % S.W);
                ws_flush_q_item<QCounter, J>(qhist.data(), it, nraw, S.X, w, c.P.qb, (unsigned long long*)(Sl + SNK_SLOT_FILE_OFF(file_of(M, mm))),
                                             (unsigned long long*)(Sl + SNK_SLOT_FILE_OFF(file_of(M, mm + M))));
            } else {
                const uint32_t bi = it - S.nq;
                const int mm = (int)(bi / (S.nwd * kSyms)), sk = (int)(bi % (S.nwd * kSyms));
                ws_flush_b_item(bstate.data() + bi, S.nb_pitch, sk / (int)S.nwd, sk % (int)S.nwd,
                                (unsigned long long*)(Sl + SNK_SLOT_FILE_OFF(file_of(M, mm))), (unsigned long long*)(Sl + SNK_SLOT_FILE_OFF(file_of(M, mm + M))));
            }
        }
        for (int t = 0; t < 2 * M; t++) {
            uint64_t* G = Sl + SNK_SLOT_FILE_OFF(file_of(M, t)) + SNK_FILE_GS_OFF;
            if (c.lastkey[t] > G[SNK_GS_LAST_KEY]) G[SNK_GS_LAST_KEY] = c.lastkey[t];
            G[SNK_GS_READS] += c.lastkey[4 + t];
            c.lastkey[t] = 0; c.lastkey[4 + t] = 0;
        }
    };
    for (int cta = 0; cta < grid; cta++) {
        const uint32_t t_begin = (uint32_t)((uint64_t)tm.ntiles * cta / grid), t_end = (uint32_t)((uint64_t)tm.ntiles * (cta + 1) / grid);
        int cur_slot = -1;
        uint32_t reads_in_hist = 0;
        for (uint32_t t = t_begin; t < t_end; t++) {
            uint32_t start, cnt;
            tile_range(tm, t, &start, &cnt);
            const uint64_t g0 = first + start;
            const int slot = slot_of(g0, (uint64_t)c.P.slot_block, c.P.n_slots);
            if (slot != cur_slot || reads_in_hist + cnt > flush_every) { if (cur_slot >= 0) flush_ws(cur_slot); cur_slot = slot; reads_in_hist = 0; }
            reads_in_hist += cnt;
            uint64_t* Sl = c.stats + (size_t)slot * SNK_SLOT_WORDS;
            for (int m = 0; m < M; m++) {
                memcpy(rows[m][0].data(), b[m]->seq + (size_t)start * c.stride, (size_t)cnt * c.stride);
                memcpy(rows[m][1].data(), b[m]->qual + (size_t)start * c.stride, (size_t)cnt * c.stride);
            }
            // scan warps: phase A per read + indicator planes (every tile slot r < R is written)
            bool tile_slow = false;
            for (int m = 0; m < M; m++)
                for (uint32_t r = 0; r < S.R; r++) {
                    if (r >= cnt) { for (int h = 0; h < kNT; h++) zero_indicators(h, ind[m].data(), r, (int)S.nwd, S.rp); continue; }
                    const uint32_t len_word = b[m]->len[start + r];
                    int len = (int)(len_word & SNK_LEN_MASK);
                    if (len > (int)c.stride || len > SNK_MAX_READ_LEN) { c.err |= ERR_BAD_LEN; len = 0; }
                    ReadInfo ri;
                    memset(&ri, 0, sizeof ri);
                    ri.head_hdcut = ri.head_lqcut = ri.tail_hdcut = ri.tail_lqcut = ri.adacut_pos = -1;
                    if (len <= 0) { ri.flags = RF_BAD_BASE | RF_QSLOW; for (int h = 0; h < kNT; h++) zero_indicators(h, ind[m].data(), r, (int)S.nwd, S.rp); }
                    else {
                        constexpr int NW = (MAXC + 1) / 2;
                        uint8_t* sq = rows[m][0].data() + (size_t)r * c.stride; uint8_t* ql = rows[m][1].data() + (size_t)r * c.stride;
                        scan_read_serial<MAXC>(sq, ql, len, nchunks, m, c.P, ri);
                        // the planes the scan warps hold after merge_scan (the scan normalised the row already: rescanning gives the same planes)
                        ScanPart<NW> Sp, S2;
                        scan_chunks<MAXC>(sq, ql, len, nchunks, c.P, 0, Sp);
                        for (int h = 1; h < kNT; h++) { scan_chunks<MAXC>(sq, ql, len, nchunks, c.P, h, S2); merge_scan(Sp, S2); }
                        for (int h = 0; h < kNT; h++) store_indicators<NW>(Sp, len, h, ind[m].data(), r, (int)S.nwd, S.rp);
                    }
                    ri.flags |= pre_flags(len_word);
                    info[m][r] = ri;
                    if (ri.flags & RF_QSLOW) tile_slow = true;
                }
            // phase P
            for (int m = 0; m < M; m++) dlist[m].clear();
            for (uint32_t r = 0; r < cnt; r++) {
                const uint64_t gi = g0 + r;
                int cat, mask = 0, fsb = -1;
                const ReadInfo& a = info[0][r];
                const ReadInfo& bb = info[M - 1][r];
                if (M == 2) {
                    cat = decide_pair(c.P, a, bb, &mask, &fsb);
                    if ((a.flags | bb.flags) & RF_BAD_BASE) c.err |= ERR_BAD_BASE;
                    if ((a.flags | bb.flags) & RF_BAD_QUAL) c.err |= ERR_BAD_QUAL;
                    if (cat == SNK_DROP_LOWQ && ((a.flags | bb.flags) & RF_LOWQ_GT1)) c.err |= ERR_LOWQ_RATIO;
                } else {
                    cat = c.P.srna ? decide_srna(c.P, a, &fsb) : decide_se(c.P, a, &fsb);
                    mask = cat ? 1 : 0;
                    if (a.flags & RF_BAD_BASE) c.err |= ERR_BAD_BASE;
                    if (a.flags & RF_BAD_QUAL) c.err |= ERR_BAD_QUAL;
                }
                if (fsb >= 0) {
                    Sl[fsb]++;
                    if (M == 2) { if (mask & 1) Sl[fsb + 1]++; if (mask & 2) Sl[fsb + 2]++; if (mask == 3) Sl[fsb + 3]++; }
                }
                for (int m = 0; m < M; m++) {
                    const ReadInfo& x = info[m][r];
                    desc[(size_t)m * S.R + r] = hist_desc(x.len, r * c.stride, x.flags & RF_QSLOW);
                    DeltaEnt de[2];
                    const int nde = delta_entries(x, cat == SNK_KEEP, r * c.stride, de);
                    for (int i = 0; i < nde; i++) dlist[m].push_back(de[i]);
                    snk_read_result res;
                    res.head_cut = (uint16_t)x.head_cut; res.clean_len = (uint16_t)x.clean_len;
                    res.category = (uint8_t)cat; res.mate_mask = (uint8_t)mask; res.adacut_pos = x.adacut_pos;
                    out[m][start + r] = res;
                    const int which = M == 2 ? m : 2;
                    int hf, tf;
                    if (c.P.cutback) {
                        trim_stat_indices(which, x.len, 0, x.head_hdcut, x.head_lqcut, x.tail_hdcut, x.tail_lqcut, x.adacut_pos, &hf, &tf);
                        uint64_t* T = Sl + SNK_SLOT_FILE_OFF(m == 0 ? SNK_RAW1 : SNK_RAW2) + SNK_FILE_TS_OFF;
                        if (hf >= 0) T[hf]++;
                        if (tf >= 0) T[tf]++;
                    }
                    if (cat == SNK_KEEP) {
                        trim_stat_indices(which, x.clean_len, x.len, x.head_hdcut, x.head_lqcut, x.tail_hdcut, x.tail_lqcut, x.adacut_pos, &hf, &tf);
                        uint64_t* T = Sl + SNK_SLOT_FILE_OFF(m == 0 ? SNK_CLEAN1 : SNK_CLEAN2) + SNK_FILE_TS_OFF;
                        if (hf >= 0) T[hf]++;
                        if (tf >= 0) T[tf]++;
                    }
                    const uint64_t kraw = ((gi + 1) << 16) | (uint64_t)(uint16_t)x.len;
                    if (kraw > c.lastkey[m]) c.lastkey[m] = kraw;
                    c.lastkey[4 + m]++;
                    if (cat == SNK_KEEP) {
                        const uint64_t kc = ((gi + 1) << 16) | (uint64_t)(uint16_t)x.clean_len;
                        if (kc > c.lastkey[M + m]) c.lastkey[M + m] = kc;
                        c.lastkey[4 + M + m]++;
                    }
                }
            }
            // histogram warps: every item, the way its owner thread runs it
            for (uint32_t it = 0; it < S.nq + S.nb; it++) {
                if (it < S.nq) {
                    const int mm = (int)(it / S.W), w = (int)(it % S.W);
                    const uint8_t* rq = rows[mm][1].data();
                    const DeltaEnt* dl = dlist[mm].data();
                    const uint32_t nd = (uint32_t)dlist[mm].size();
                    if (!tile_slow) {
                        const int cell0 = (int)it * 2 * (int)sizeof(QCounter) - c.P.phred * q_bstep;
                        unit_q_raw<QCounter, J, J>(rq, c.stride, cnt, w, 0, (uint8_t*)qhist.data(), cell0, q_jstep, q_bstep);
                        unit_q_delta<QCounter, J, J>(rq, dl, nd, w, 0, (uint8_t*)qhist.data(), cell0 + (int)nraw * 2 * (int)sizeof(QCounter), q_jstep, q_bstep);
                    } else {
                        c.err |= unit_q_checked<QCounter, J>(rq, desc.data() + (size_t)mm * S.R, cnt, dl, nd, w, 0, J, c.P.phred, c.P.qb, qhist.data() + 2u * it,
                                                             nraw, (int)S.X, (unsigned long long*)(Sl + SNK_SLOT_FILE_OFF(file_of(M, mm))),
                                                             (unsigned long long*)(Sl + SNK_SLOT_FILE_OFF(file_of(M, mm + M))));
                    }
                } else {
                    const uint32_t bi = it - S.nq;
                    const int mm = (int)(bi / (S.nwd * kSyms)), sk = (int)(bi % (S.nwd * kSyms));
                    const int sym = sk / (int)S.nwd, k = sk % (int)S.nwd;
                    ws_b_raw(ind[mm].data() + ind_index(sym, k, 0, (int)S.nwd, S.rp), cnt, bstate.data() + bi, S.nb_pitch);
                    ws_b_delta(ind[mm].data() + ind_index(sym, 0, 0, (int)S.nwd, S.rp), dlist[mm].data(), (uint32_t)dlist[mm].size(), k, c.stride, magic,
                               (int)S.nwd, S.rp, bstate.data() + (size_t)kVPlanes * S.nb_pitch + bi, S.nb_pitch);
                }
            }
        }
        if (cur_slot >= 0) flush_ws(cur_slot);
    }
}

} // namespace

extern "C" {

// The warp-specialised kernel's replay: same contract as coretest_filter; wpg = scan group size (0: the launcher's
// choice). Returns 1 when the shape is not served by that kernel (long rows).
int coretest_filter_ws(const snk_params* p, const snk_batch* r1, const snk_batch* r2, snk_read_result* out1, snk_read_result* out2,
                       uint64_t* stats, uint64_t first, uint32_t* err, int wpg, int grid, int qb_override)
{
    uint32_t flush_every = kQCounterMax;
    if (const char* fe = getenv("SNK_CORETEST_FLUSH_EVERY")) flush_every = (uint32_t)atoi(fe);
    Ctx c;
    prepare_params(*p, c.P);
    std::vector<ContamDev> contams(2 * SNK_MAX_CONTAMS);
    prepare_contams(*p, contams.data());
    c.P.contams = contams.data();
    std::vector<GContamDev> gcontams(SNK_MAX_CONTAMS);
    prepare_gcontams(*p, gcontams.data());
    c.P.gcontams = gcontams.data();
    if (qb_override >= 0) c.P.qb = qb_override;
    c.stats = stats;
    c.mates = p->is_pe ? 2 : 1;
    c.stride = r1->stride;
    WsShape S;
    if (!ws_make_shape(c.mates, c.stride, c.P.qb, ada_slots(c.P.n_adapters), 227u * 1024u, (uint32_t)wpg, S)) return 1;
    const snk_batch* b[2] = {r1, r2};
    snk_read_result* out[2] = {out1, out2};
    const uint32_t chunks = c.stride / 16;
    if (grid < 1) grid = 1;
    if (chunks <= 4) run_ws<4>(c, S, b, out, first, grid, flush_every);
    else if (chunks <= 7) run_ws<7>(c, S, b, out, first, grid, flush_every);
    else if (chunks <= 10) run_ws<10>(c, S, b, out, first, grid, flush_every);
    else run_ws<16>(c, S, b, out, first, grid, flush_every);
    *err |= c.err;
    return 0;
}

// Same contract as snk_filter_pe_host / snk_filter_se_host, on the CPU, accumulating into `stats`
// (n_slots * SNK_SLOT_WORDS). tile_r = 0 picks the kernel's default tile size; grid = CTAs to mimic.
int coretest_filter(const snk_params* p, const snk_batch* r1, const snk_batch* r2, snk_read_result* out1,
                    snk_read_result* out2, uint64_t* stats, uint64_t first, uint32_t* err, int tile_r, int grid, int qb_override)
{
    // flush_every: the kernel flushes its 16-bit cells before kQCounterMax records; tests shrink it via the environment
    uint32_t flush_every = kQCounterMax;
    if (const char* fe = getenv("SNK_CORETEST_FLUSH_EVERY")) flush_every = (uint32_t)atoi(fe);

    Ctx c;
    prepare_params(*p, c.P);
    std::vector<ContamDev> contams(2 * SNK_MAX_CONTAMS);
    prepare_contams(*p, contams.data());
    c.P.contams = contams.data();
    std::vector<GContamDev> gcontams(SNK_MAX_CONTAMS);
    prepare_gcontams(*p, gcontams.data());
    c.P.gcontams = gcontams.data();
    if (qb_override >= 0) c.P.qb = qb_override;
    c.stats = stats;
    c.mates = p->is_pe ? 2 : 1;
    c.stride = r1->stride;
    c.J = hist_j(c.stride);
    c.W = c.stride / c.J;
    c.X = align_up(hist_items(c.mates, c.stride), 32);
    c.threads = cta_threads(c.mates, c.stride);
    c.R = tile_r > 0 ? (uint32_t)tile_r : c.threads / (c.mates * kNT);
    const snk_batch* b[2] = {r1, r2};
    snk_read_result* out[2] = {out1, out2};
    const uint32_t chunks = c.stride / 16;
    if (grid < 1) grid = 1;
    if (chunks <= 4) run<4, 4>(c, b, out, first, grid, flush_every);
    else if (chunks <= 7) run<7, 4>(c, b, out, first, grid, flush_every);
    else if (chunks <= 10) run<10, 4>(c, b, out, first, grid, flush_every);
    else if (chunks <= 16) run<16, 4>(c, b, out, first, grid, flush_every);
    else if (chunks <= 32) run<32, 4>(c, b, out, first, grid, flush_every);
    else run<63, 4>(c, b, out, first, grid, flush_every);
    *err |= c.err;
    return 0;
}

// ---- text path replay (text_core.cuh driven the way text_kernels.cuh drives it, sequentially) ----
// text must be readable 32 bytes past `bytes`. line_off has 4n+1 entries. Returns the TextFlags.
uint32_t coretest_text_index_pack(const uint8_t* text, uint32_t bytes, uint32_t n, uint32_t strip, uint32_t stride,
                                  uint32_t* line_off, uint8_t* seq, uint8_t* qual, uint16_t* len, uint32_t* max_len)
{
    uint32_t flags = 0, k = 0;
    const uint32_t want = 4u * n;
    line_off[0] = 0;
    for (uint32_t off = 0; off < bytes; off += 16) {
        U4 v = load16(text + off);
        uint32_t mask = newline_mask16(v, (int)std::min<int64_t>((int64_t)bytes - off, 16));
        while (mask) {
            const int b = ctz32(mask);
            mask &= mask - 1u;
            k++;
            if (k <= want) line_off[k] = off + (uint32_t)b + 1u;
        }
    }
    if (k == want) {}
    else if (k + 1 == want && bytes > 0) line_off[want] = bytes;
    else return TEXT_LINE_COUNT;
    *max_len = 0;
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t* off = line_off + 4u * (size_t)i;
        const uint32_t sn = line_visible(off, 1, strip), qn = line_visible(off, 3, strip);
        if (sn != qn) flags |= TEXT_LEN_MISMATCH;
        if (sn > SNK_MAX_READ_LEN) flags |= TEXT_TOO_LONG;
        if (sn > stride) flags |= TEXT_STRIDE_OVERFLOW;
        if (sn > *max_len) *max_len = sn;
        len[i] = (uint16_t)std::min(sn, stride);
        for (uint32_t c = 0; c < stride / 16; c++) {
            const U4 a = pack_chunk(text, off[1], std::min(sn, stride), c);
            const U4 b = pack_chunk(text, off[3], std::min(std::min(qn, sn), stride), c);
            memcpy(seq + (size_t)i * stride + 16 * c, &a, 16);
            memcpy(qual + (size_t)i * stride + 16 * c, &b, 16);
        }
    }
    return flags;
}

// text path's id parse for the tile / fov removal lists (text_core.cuh id_prefilter)
uint32_t coretest_id_flags(const snk_params* p, const uint8_t* id, uint32_t n)
{
    IdFilter F;
    make_id_filter(*p, F);
    return id_prefilter(id, n, F);
}

// rec_off has n+1 entries; out must hold bytes + 2n + 64. `lanes` mimics the warp width. Returns the clean text size.
uint64_t coretest_text_format(const uint8_t* text, const uint32_t* line_off, const uint8_t* seq, const uint8_t* qual,
                              const snk_read_result* res, uint32_t n, uint32_t stride, int mate, int strip, int pe_info, int fasta,
                              int id_mode, int qshift, uint32_t lanes, uint8_t* out, uint32_t* rec_off)
{
    TextFormat F = {strip, pe_info, fasta, id_mode, qshift};
    uint64_t total = 0;
    for (uint32_t i = 0; i < n; i++) {
        rec_off[i] = (uint32_t)total;
        if (res[i].category != SNK_KEEP) continue;
        const uint32_t* off = line_off + 4u * (size_t)i;
        uint32_t idn = line_visible(off, 0, (uint32_t)strip);
        if (id_mode) idn = id_transform(text + off[0], idn, id_mode, nullptr);
        total += record_out_len(idn, res[i].clean_len, F);
    }
    rec_off[n] = (uint32_t)total;
    for (uint32_t i = 0; i < n; i++) {
        if (rec_off[i + 1] == rec_off[i]) continue;
        const uint32_t* off = line_off + 4u * (size_t)i;
        const uint8_t* id = text + off[0];
        const uint32_t idn = line_visible(off, 0, (uint32_t)strip);
        uint8_t* dst = out + rec_off[i];
        uint32_t id_out = idn;
        if (id_mode == 0) memcpy(dst, id, idn);
        else id_out = id_transform(id, idn, id_mode, dst);
        const size_t row = (size_t)i * stride + res[i].head_cut;
        for (uint32_t lane = 0; lane < lanes; lane++)
            format_tail(dst + id_out, seq + row, qual + row, res[i].clean_len, mate, F, lane, lanes);
        if (fasta) fasta_fix(dst, id_out + 2u * (uint32_t)pe_info);
    }
    return total;
}

}
