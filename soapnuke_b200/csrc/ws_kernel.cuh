// ws_kernel.cuh — the warp-specialised fused filter + statistics kernel (sm_100a), rows up to kWsMaxStride bytes.
//
// One persistent CTA per SM, kWsThreads threads, three roles that never meet at a CTA-wide barrier:
//
//   producer   1 thread        (last warp) walks the CTA's tiles in order; per tile 2*MATES bulk copies (TMA, cp.async.bulk) of
//                              the tile's contiguous seq / qual rows into the next free stage; completion on full[s]
//   scan       16 warps        groups of `wpg` warps take tiles round robin (tile lt -> group lt % ngroups). A warp
//                              owns 8 pairs (16 reads) of its group's tile: two adjacent lanes per read run phase A
//                              (scan_read_coop: counters, predicates, adapter search, trim), the pair's mates sit in
//                              the same warp, so the discard cascade, the 8-byte result records, the filter
//                              counters, the trim tables and the tile's delta list (phase P) need nothing but warp
//                              shuffles. The warp also stores each read's five indicator planes. -> scanned[s]
//   histogram  7 warps         every thread owns whole work items for the launch: a q-item (mate, 4 positions) counts
//                              the quality x position cells in shared memory (owner computes, no atomics), a b-item
//                              (mate, plane word, symbol) adds indicator words into vertical counters (ws_core.cuh);
//                              raw walk over the tile's rows, then the delta list; items are flushed by their owner
//                              with 64-bit global atomics when the CTA's statistics slot changes, before a 16-bit
//                              cell can wrap, and at the end. -> empty[s]
//
// Stage hand-over is by mbarriers only (full: TMA bytes landed; scanned: wpg warps arrived; empty: histogram warps
// arrived), so the TMA latency and the histogram walk of tile t overlap the scan of tiles t+1 .. t+ngroups.
#pragma once
#include <cuda_runtime.h>
#include "filter_kernel.cuh"
#include "ws_core.cuh"

namespace snkcore {

struct WsArgs {
    KernelArgs k;           // R / items_w / X hold the ws shape's values
    WsShape s;
    uint32_t magic;         // stride_magic(stride)
};

#ifdef __CUDACC__

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct WsStage {
    uint8_t* rows[2][2];
    uint32_t* ind[2];
    uint32_t* desc;
    DeltaEnt* dlist;
    uint32_t* ctl;          // ndelta[0], ndelta[1], tile_slow
};
__device__ __forceinline__ WsStage ws_stage(uint8_t* smem, const WsShape& S, uint32_t s)
{
    uint8_t* b = smem + S.off_stage0 + (size_t)s * S.stage_bytes;
    WsStage st;
    for (int m = 0; m < 2; m++) {
        st.rows[m][0] = b + S.so_rows[m][0]; st.rows[m][1] = b + S.so_rows[m][1];
        st.ind[m] = reinterpret_cast<uint32_t*>(b + S.so_ind[m]);
    }
    st.desc = reinterpret_cast<uint32_t*>(b + S.so_desc);
    st.dlist = reinterpret_cast<DeltaEnt*>(b + S.so_delta);
    st.ctl = reinterpret_cast<uint32_t*>(b + S.so_ctl);
    return st;
}

// ---------------------------------------------------------------------------------------------- scan role
template <int MAXC, int MATES>
__device__ __forceinline__ void ws_scan_role(const DevParams& P, const WsArgs& W, uint8_t* smem, unsigned long long* bar_full,
                                             unsigned long long* bar_scanned, uint32_t t_begin, uint32_t nloc, int sw, int lane)
{
    const KernelArgs& A = W.k;
    const WsShape& S = W.s;
    constexpr uint32_t RPW = 32u / (2u * MATES);          // pairs (reads) per warp
    const uint32_t full_groups = (uint32_t)kWsScanWarps / S.wpg;
    const uint32_t g = S.interleave ? (uint32_t)sw % full_groups : (uint32_t)sw / S.wpg;
    const uint32_t wg = S.interleave ? (uint32_t)sw / full_groups : (uint32_t)sw % S.wpg;
    if (g >= S.ngroups) return;                            // the pipeline has fewer stages than groups (long rows)
    const int h = lane & 1;
    const int m = (MATES == 2) ? ((lane >> 1) & 1) : 0;
    const uint32_t pj = (uint32_t)lane / (2u * MATES);
    const uint32_t r = wg * RPW + pj;                      // my record of the tile
    const unsigned pm = 3u << (lane & ~1);
    const bool actor = h == 0;                             // one lane per read acts in phase P
    const int nchunks = (int)(A.stride / 16);
    const AdaHot* ada = reinterpret_cast<const AdaHot*>(smem + S.off_ada) + ada_first_slot(P.n_adapters, m);
    // lanes of my mate's actors (append order of the delta list, leaders of the per-mate shared atomics)
    const unsigned mine = (MATES == 2) ? (0x11111111u << (2 * m)) : 0x55555555u;
    const int leader = (MATES == 2) ? 2 * m : 0;
    const unsigned lt_mask = ((1u << lane) - 1u) & mine;
    // per-lane share of the slot's "last record" keys and record counts (raw table, clean table of my mate)
    unsigned long long key_raw = 0, key_clean = 0;
    uint32_t n_raw = 0, n_clean = 0;
    int cur_slot = -1;
    auto flush_keys = [&]() {
        if (cur_slot < 0) return;
#pragma unroll
        for (int d = 2 * MATES; d < 32; d <<= 1) {           // lanes with equal lane % (2*MATES) hold the same mate's actors
            const unsigned long long kr = __shfl_xor_sync(0xFFFFFFFFu, key_raw, d), kc = __shfl_xor_sync(0xFFFFFFFFu, key_clean, d);
            key_raw = kr > key_raw ? kr : key_raw; key_clean = kc > key_clean ? kc : key_clean;
            n_raw += __shfl_xor_sync(0xFFFFFFFFu, n_raw, d); n_clean += __shfl_xor_sync(0xFFFFFFFFu, n_clean, d);
        }
        if (lane == leader) {
            unsigned long long* Sl = A.stats + (size_t)cur_slot * SNK_SLOT_WORDS;
            unsigned long long* Gr = Sl + SNK_SLOT_FILE_OFF(file_of_tab(MATES, m)) + SNK_FILE_GS_OFF;
            unsigned long long* Gc = Sl + SNK_SLOT_FILE_OFF(file_of_tab(MATES, m + MATES)) + SNK_FILE_GS_OFF;
            if (n_raw) { atomicMax(&Gr[SNK_GS_LAST_KEY], key_raw); atomicAdd(&Gr[SNK_GS_READS], (unsigned long long)n_raw); }
            if (n_clean) { atomicMax(&Gc[SNK_GS_LAST_KEY], key_clean); atomicAdd(&Gc[SNK_GS_READS], (unsigned long long)n_clean); }
        }
        key_raw = key_clean = 0; n_raw = n_clean = 0;
    };

    // software pipeline: the tile's len[] word is loaded one trip ahead (the rows come by TMA)
    uint32_t lt = g, start = 0, cnt = 0, len_word = 0;
    if (lt < nloc) {
        tile_range(A.tm, t_begin + lt, &start, &cnt);
        len_word = r < cnt ? (uint32_t)A.len[m][start + r] : 0u;
    }
    for (; lt < nloc; lt += S.ngroups) {
        const uint32_t my_start = start, my_cnt = cnt, my_len_word = len_word;
        if (lt + S.ngroups < nloc) {
            tile_range(A.tm, t_begin + lt + S.ngroups, &start, &cnt);
            len_word = r < cnt ? (uint32_t)A.len[m][start + r] : 0u;
        }
        const uint64_t g0 = A.tm.first + my_start;
        const int slot = slot_of(g0, (uint64_t)P.slot_block, P.n_slots);
        if (slot != cur_slot) { flush_keys(); cur_slot = slot; }
        const uint32_t s = lt % S.nstages;
        const WsStage st = ws_stage(smem, S, s);
        mbar_wait(&bar_full[s], (lt / S.nstages) & 1u);

        // ---- phase A
        const bool live = r < my_cnt;
        ReadInfo ri;
        ri.len = 0; ri.head_cut = 0; ri.clean_len = 0; ri.head_hdcut = ri.head_lqcut = ri.tail_hdcut = ri.tail_lqcut = ri.adacut_pos = -1;
        ri.flags = 0;
        int len = (int)(my_len_word & SNK_LEN_MASK);
        if (live && (len > (int)A.stride || len > SNK_MAX_READ_LEN)) {
            if (actor) report_error(A, ERR_BAD_LEN, g0 + r);
            len = 0;
        }
        if (live && len > 0) {
            scan_read_coop<MAXC, MATES>(st.rows[m][0] + (size_t)r * A.stride, st.rows[m][1] + (size_t)r * A.stride, len, nchunks, m, P, ada, h,
                                        pm, ri, st.ind[m], r, (int)S.nwd, S.rp);
        } else {
            zero_indicators(h, st.ind[m], r, (int)S.nwd, S.rp);
            if (live) ri.flags = RF_BAD_BASE | RF_QSLOW;      // "Error:empty sequence" (read_filter.cpp:250): row left as staged
        }
        if (live) {
            ri.flags |= pre_flags(my_len_word);
            if (actor && (ri.flags & RF_QSLOW)) st.ctl[2] = 1u;
        }
        __syncwarp();

        // ---- phase P: the pair's mates are two lanes apart
        int cat = SNK_DROP_EMPTY, mask = 0, fsb = -1;
        uint32_t err = 0;
        if (MATES == 2) {
            const uint32_t mine_w = (uint32_t)(uint16_t)ri.clean_len | ((uint32_t)ri.flags << 16);
            const uint32_t other_w = __shfl_xor_sync(0xFFFFFFFFu, mine_w, 2);
            ReadInfo o = ri;
            o.clean_len = (int16_t)(other_w & 0xFFFFu); o.flags = (uint16_t)(other_w >> 16);
            const ReadInfo& a = m ? o : ri;
            const ReadInfo& b = m ? ri : o;
            cat = decide_pair(P, a, b, &mask, &fsb);
            if ((a.flags | b.flags) & RF_BAD_BASE) err |= ERR_BAD_BASE;
            if ((a.flags | b.flags) & RF_BAD_QUAL) err |= ERR_BAD_QUAL;
            if (cat == SNK_DROP_LOWQ && ((a.flags | b.flags) & RF_LOWQ_GT1)) err |= ERR_LOWQ_RATIO;
        } else {
            cat = P.srna ? decide_srna(P, ri, &fsb) : decide_se(P, ri, &fsb);
            mask = cat ? 1 : 0;
            if (ri.flags & RF_BAD_BASE) err |= ERR_BAD_BASE;
            if (ri.flags & RF_BAD_QUAL) err |= ERR_BAD_QUAL;
        }
        const bool act = live && actor;
        const uint64_t gi = g0 + r;
        DeltaEnt de[2];
        int nde = 0;
        if (act) {
            unsigned long long* Sl = A.stats + (size_t)slot * SNK_SLOT_WORDS;
            if (m == 0) {                                   // once per pair
                if (err) report_error(A, err, gi);
                if (fsb >= 0) {
                    atomicAdd(&Sl[fsb], 1ull);
                    if (MATES == 2) {
                        if (mask & 1) atomicAdd(&Sl[fsb + 1], 1ull);
                        if (mask & 2) atomicAdd(&Sl[fsb + 2], 1ull);
                        if (mask == 3) atomicAdd(&Sl[fsb + 3], 1ull);
                    }
                }
            }
            const uint32_t row0 = r * A.stride;
            st.desc[(size_t)m * S.R + r] = hist_desc(ri.len, row0, ri.flags & RF_QSLOW);
            nde = delta_entries(ri, cat == SNK_KEEP, row0, de);
            const unsigned long long packed = (unsigned long long)(uint16_t)ri.head_cut |
                ((unsigned long long)(uint16_t)ri.clean_len << 16) | ((unsigned long long)(uint8_t)cat << 32) |
                ((unsigned long long)(uint8_t)mask << 40) | ((unsigned long long)(uint16_t)ri.adacut_pos << 48);
            reinterpret_cast<unsigned long long*>(A.out[m])[my_start + r] = packed;
            const int which = (MATES == 2) ? m : 2;
            int hf, tf;
            if (P.cutback) {
                trim_stat_indices(which, ri.len, 0, ri.head_hdcut, ri.head_lqcut, ri.tail_hdcut, ri.tail_lqcut, ri.adacut_pos, &hf, &tf);
                unsigned long long* T = Sl + SNK_SLOT_FILE_OFF(m == 0 ? SNK_RAW1 : SNK_RAW2) + SNK_FILE_TS_OFF;
                if (hf >= 0) atomicAdd(&T[hf], 1ull);
                if (tf >= 0) atomicAdd(&T[tf], 1ull);
            }
            if (cat == SNK_KEEP) {
                trim_stat_indices(which, ri.clean_len, ri.len, ri.head_hdcut, ri.head_lqcut, ri.tail_hdcut, ri.tail_lqcut, ri.adacut_pos, &hf, &tf);
                unsigned long long* T = Sl + SNK_SLOT_FILE_OFF(m == 0 ? SNK_CLEAN1 : SNK_CLEAN2) + SNK_FILE_TS_OFF;
                if (hf >= 0) atomicAdd(&T[hf], 1ull);
                if (tf >= 0) atomicAdd(&T[tf], 1ull);
            }
            const unsigned long long kr = ((gi + 1) << 16) | (unsigned long long)(uint16_t)ri.len;
            key_raw = kr > key_raw ? kr : key_raw; n_raw++;
            if (cat == SNK_KEEP) {
                const unsigned long long kc = ((gi + 1) << 16) | (unsigned long long)(uint16_t)ri.clean_len;
                key_clean = kc > key_clean ? kc : key_clean; n_clean++;
            }
        }
        // append the delta entries to the stage's per-mate lists: one shared atomic per warp and mate
        {
            const unsigned b1 = __ballot_sync(0xFFFFFFFFu, nde >= 1) & mine, b2 = __ballot_sync(0xFFFFFFFFu, nde >= 2) & mine;
            uint32_t base = 0;
            if (lane == leader && b1) base = atomicAdd(&st.ctl[m], (uint32_t)(__popc(b1) + __popc(b2)));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (nde >= 1) {
                DeltaEnt* dl = st.dlist + (size_t)m * 2u * S.R + base + __popc(b1 & lt_mask) + __popc(b2 & lt_mask);
                dl[0] = de[0];
                if (nde >= 2) dl[1] = de[1];
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_scanned[s]);
    }
    flush_keys();
}

// ---------------------------------------------------------------------------------------------- histogram role
template <int MATES>
__device__ __forceinline__ void ws_hist_role(const DevParams& P, const WsArgs& W, uint8_t* smem, unsigned long long* bar_scanned,
                                             unsigned long long* bar_empty, uint32_t t_begin, uint32_t nloc, int ht, int lane)
{
    constexpr int J = kWsJ;
    const KernelArgs& A = W.k;
    const WsShape& S = W.s;
    QCounter* qhist = reinterpret_cast<QCounter*>(smem + S.off_qhist);
    uint32_t* bstate = reinterpret_cast<uint32_t*>(smem + S.off_bstate);
    const uint32_t nraw = S.nq;
    const uint32_t nitems = S.nq + S.nb;
    const int q_jstep = (int)S.X * 2 * (int)sizeof(QCounter), q_bstep = (J / 2) * q_jstep;
    int cur_slot = -1;
    uint32_t reads_in_hist = 0;

    auto flush = [&](int slot) {
        unsigned long long* Sl = A.stats + (size_t)slot * SNK_SLOT_WORDS;
        for (uint32_t it = (uint32_t)ht; it < nitems; it += kWsHistThreads) {
            if (it < S.nq) {
                const int mm = (int)(it / S.W), w = (int)(it % S.W);
                ws_flush_q_item<QCounter, J>(qhist, it, nraw, S.X, w, P.qb, Sl + SNK_SLOT_FILE_OFF(file_of_tab(MATES, mm)),
                                             Sl + SNK_SLOT_FILE_OFF(file_of_tab(MATES, mm + MATES)));
            } else {
                const uint32_t bi = it - S.nq;
                const int mm = (int)(bi / (S.nwd * kSyms)), sk = (int)(bi % (S.nwd * kSyms));
                ws_flush_b_item(bstate + bi, S.nb_pitch, sk / (int)S.nwd, sk % (int)S.nwd, Sl + SNK_SLOT_FILE_OFF(file_of_tab(MATES, mm)),
                                Sl + SNK_SLOT_FILE_OFF(file_of_tab(MATES, mm + MATES)));
            }
        }
    };

    for (uint32_t lt = 0; lt < nloc; lt++) {
        uint32_t start, cnt;
        tile_range(A.tm, t_begin + lt, &start, &cnt);
        const uint64_t g0 = A.tm.first + start;
        const int slot = slot_of(g0, (uint64_t)P.slot_block, P.n_slots);
        if (slot != cur_slot || reads_in_hist + cnt > kQCounterMax) {
            if (cur_slot >= 0) flush(cur_slot);
            cur_slot = slot;
            reads_in_hist = 0;
        }
        reads_in_hist += cnt;
        const uint32_t s = lt % S.nstages;
        const WsStage st = ws_stage(smem, S, s);
        mbar_wait(&bar_scanned[s], (lt / S.nstages) & 1u);
        const bool slow = st.ctl[2] != 0u;
        for (uint32_t it = (uint32_t)ht; it < nitems; it += kWsHistThreads) {
            if (it < S.nq) {
                const int mm = (int)(it / S.W), w = (int)(it % S.W);
                const uint8_t* rows_q = st.rows[mm][1];
                const DeltaEnt* dl = st.dlist + (size_t)mm * 2u * S.R;
                const uint32_t nd = st.ctl[mm];
                if (!slow) {
                    const int cell0 = (int)it * 2 * (int)sizeof(QCounter) - P.phred * q_bstep;
                    unit_q_raw<QCounter, J, J>(rows_q, A.stride, cnt, w, 0, reinterpret_cast<uint8_t*>(qhist), cell0, q_jstep, q_bstep);
                    unit_q_delta<QCounter, J, J>(rows_q, dl, nd, w, 0, reinterpret_cast<uint8_t*>(qhist),
                                                 cell0 + (int)nraw * 2 * (int)sizeof(QCounter), q_jstep, q_bstep);
                } else {
                    unsigned long long* Sl = A.stats + (size_t)slot * SNK_SLOT_WORDS;
                    const uint32_t err = unit_q_checked<QCounter, J>(rows_q, st.desc + (size_t)mm * S.R, cnt, dl, nd, w, 0, J, P.phred, P.qb,
                                                                     qhist + 2u * it, nraw, (int)S.X,
                                                                     Sl + SNK_SLOT_FILE_OFF(file_of_tab(MATES, mm)),
                                                                     Sl + SNK_SLOT_FILE_OFF(file_of_tab(MATES, mm + MATES)));
                    if (err) report_error(A, err, g0);
                }
            } else {
                const uint32_t bi = it - S.nq;
                const int mm = (int)(bi / (S.nwd * kSyms)), sk = (int)(bi % (S.nwd * kSyms));
                const int sym = sk / (int)S.nwd, k = sk % (int)S.nwd;
                const uint32_t* ind = st.ind[mm];
                ws_b_raw(ind + ind_index(sym, k, 0, (int)S.nwd, S.rp), cnt, bstate + bi, S.nb_pitch);
                ws_b_delta(ind + ind_index(sym, 0, 0, (int)S.nwd, S.rp), st.dlist + (size_t)mm * 2u * S.R, st.ctl[mm], k, A.stride, W.magic,
                           (int)S.nwd, S.rp, bstate + (size_t)kVPlanes * S.nb_pitch + bi, S.nb_pitch);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_empty[s]);
    }
    if (cur_slot >= 0) flush(cur_slot);
}

// ---------------------------------------------------------------------------------------------- kernel
template <int MAXC, int MATES>
__global__ void __maxnreg__(kWsMaxRegs) filter_ws_kernel(const __grid_constant__ DevParams P, const __grid_constant__ WsArgs W)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const KernelArgs& A = W.k;
    const WsShape& S = W.s;
    if (A.skip_word && (*A.skip_word & A.skip_mask)) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned long long* bar_full = reinterpret_cast<unsigned long long*>(smem + S.off_bars);
    unsigned long long* bar_scanned = bar_full + kWsMaxStages;
    unsigned long long* bar_empty = bar_full + 2 * kWsMaxStages;

    for (uint32_t e = tid; e < (S.off_ada - S.off_qhist) / 4; e += blockDim.x) reinterpret_cast<uint32_t*>(smem + S.off_qhist)[e] = 0;   // cells + vertical counters
    for (uint32_t e = tid; e < ada_slots(P.n_adapters); e += blockDim.x) {
        const int m = e >= ada_first_slot(P.n_adapters, 1) ? 1 : 0;
        make_ada_hot(P.ada[m][e - ada_first_slot(P.n_adapters, m)], reinterpret_cast<AdaHot*>(smem + S.off_ada)[e]);
    }
    if (tid == 0) {
        for (uint32_t s = 0; s < S.nstages; s++) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_scanned[s], S.wpg);
            mbar_init(&bar_empty[s], kWsHistWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint32_t nt = A.tm.ntiles;
    const uint32_t t_begin = (uint32_t)((uint64_t)nt * blockIdx.x / gridDim.x);
    const uint32_t t_end = (uint32_t)((uint64_t)nt * (blockIdx.x + 1) / gridDim.x);
    const uint32_t nloc = t_end - t_begin;

    if (warp < kWsScanWarps) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kWsScanRegs));
        ws_scan_role<MAXC, MATES>(P, W, smem, bar_full, bar_scanned, t_begin, nloc, warp, lane);
        return;
    }
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kWsHistRegs));
    if (warp == kWsScanWarps + kWsHistWarps) {
        // ---- producer: one thread keeps the TMA engine fed, one stage per tile
        if (lane != 0) return;
        for (uint32_t lt = 0; lt < nloc; lt++) {
            const uint32_t s = lt % S.nstages, k = lt / S.nstages;
            if (k > 0) mbar_wait(&bar_empty[s], (k - 1) & 1u);
            uint32_t start, cnt;
            tile_range(A.tm, t_begin + lt, &start, &cnt);
            const WsStage st = ws_stage(smem, S, s);
            st.ctl[0] = 0; st.ctl[1] = 0; st.ctl[2] = 0;
            // the stage's previous rows were written (padding normalised) and read through the generic proxy: order those
            // accesses before the async-proxy writes of the bulk copies
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            const uint32_t row_bytes = cnt * A.stride;
            mbar_expect_tx(&bar_full[s], 2u * MATES * row_bytes);
#pragma unroll
            for (int m = 0; m < MATES; m++) {
                bulk_g2s(st.rows[m][0], A.seq[m] + (size_t)start * A.stride, row_bytes, &bar_full[s]);
                bulk_g2s(st.rows[m][1], A.qual[m] + (size_t)start * A.stride, row_bytes, &bar_full[s]);
            }
        }
    } else {
        ws_hist_role<MATES>(P, W, smem, bar_scanned, bar_empty, t_begin, nloc, tid - 32 * kWsScanWarps, lane);
    }
}

#endif // __CUDACC__

} // namespace snkcore
