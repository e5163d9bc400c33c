// ws_core.cuh — histogram work units of the warp-specialised filter kernel (ws_kernel.cuh).
//
// Like filter_core.cuh everything here is `__host__ __device__`, so that tests/coretest can replay the
// kernel's work units on a CPU. What is new relative to filter_core.cuh:
//
//  * base x position counts (position_acgt_content of stat_pe_fqs / stat_se_fqs, peprocess.cpp:1144-1203,
//    seprocess.cpp:683-739) are no longer counted byte by byte. The scan warps already hold every read as bit
//    planes (32 bases per word); they store five INDICATOR planes per read (is-A, is-C, is-G, is-T, is-N) and a
//    b-item (mate, plane word k, symbol) adds those words into VERTICAL bit-sliced counters: plane b of the
//    counter holds bit b of the 32 per-position counts. Eight records go through a carry-save adder tree
//    (Harley-Seal) for 14 logic operations, i.e. < 2 instructions per record for 32 positions, against ~13
//    instructions per record for 4 positions with packed byte arithmetic.
//  * the clean tables stay "raw - delta": delta entries (dropped records, trimmed ends, 5'-cut records re-added
//    at their shifted positions) are applied to a second, signed vertical counter by ripple add / subtract.
//  * quality x position cells are the owner-computes 16-bit cells of filter_core.cuh (unit_q_raw / unit_q_delta /
//    unit_q_checked), one thread per item of J positions, flushed by their owner without any CTA-wide barrier.
#pragma once
#include "filter_core.cuh"

namespace snkcore {

// ------------------------------------------------------------------ indicator planes
// One block per mate and stage: ind[(sym * nwd + k) * rp + r], sym 0..4 = A,C,G,T,N (the order of the
// position_acgt_content columns), k = plane word (bases 32k..32k+31), r = record of the tile, rp = row pitch
// (tile capacity + 4 words: keeps 16-byte loads of 4 consecutive records aligned and bank-conflict free).
constexpr int kSyms = 5;
SNK_HD uint32_t ind_index(int sym, int k, uint32_t r, int nwd, uint32_t rp) { return (uint32_t)(sym * nwd + k) * rp + r; }
SNK_HD uint32_t ind_words(int nwd, uint32_t rp) { return (uint32_t)(kSyms * nwd) * rp; }

// thread h of the read's group stores the plane words k with k % kNT == h (merged planes: both threads hold all words)
template <int NW>
SNK_HD void store_indicators(const ScanPart<NW>& S, int len, int h, uint32_t* ind, uint32_t r, int nwd, uint32_t rp)
{
#pragma unroll
    for (int k = 0; k < NW; k++) {
        if ((k % kNT) != h || k >= nwd) continue;
        const uint32_t pv = plane_valid(len, k);
        const uint32_t p0 = S.p0[k], p1 = S.p1[k], pn = S.pn[k];
        ind[ind_index(0, k, r, nwd, rp)] = pv & ~(p0 | p1);       // A: no code bit (bases behind the read carry none either)
        ind[ind_index(1, k, r, nwd, rp)] = p0 & ~p1;              // C
        ind[ind_index(2, k, r, nwd, rp)] = p0 & p1 & ~pn;         // G
        ind[ind_index(3, k, r, nwd, rp)] = ~p0 & p1;              // T
        ind[ind_index(4, k, r, nwd, rp)] = pn;                    // N
    }
}
// records that are not scanned (tile slots behind the last record, empty rows): no base anywhere
SNK_HD void zero_indicators(int h, uint32_t* ind, uint32_t r, int nwd, uint32_t rp)
{
    for (int k = h; k < nwd; k += kNT)
        for (int s = 0; s < kSyms; s++) ind[ind_index(s, k, r, nwd, rp)] = 0u;
}

// ------------------------------------------------------------------ vertical counters
// State of one b-item: kVPlanes words for the raw count (unsigned, < 2^16 records per flush interval) and
// kVPlanes words for "removed - added" (two's complement modulo 2^16). Word `plane` of item `it` lives at
// bstate[(set * kVPlanes + plane) * pitch + it]: consecutive items in consecutive banks.
constexpr int kVPlanes = 16;
SNK_HD uint32_t bstate_words(uint32_t pitch) { return 2u * kVPlanes * pitch; }

// full adder on 32 independent bit columns
SNK_HD void csa(uint32_t& hi, uint32_t& lo, uint32_t a, uint32_t b, uint32_t c)
{
    const uint32_t u = a ^ b;
    hi = (a & b) | (u & c);
    lo = u ^ c;
}
// add the one-bit-per-column word `carry` at weight 2^first (ripple; stops as soon as nothing is carried on)
SNK_HD void vripple_add(uint32_t* v, uint32_t pitch, int first, uint32_t carry)
{
    for (int b = first; b < kVPlanes && carry; b++) {
        const uint32_t x = v[(size_t)b * pitch];
        v[(size_t)b * pitch] = x ^ carry;
        carry &= x;
    }
}
SNK_HD void vripple_sub(uint32_t* v, uint32_t pitch, uint32_t borrow)
{
    for (int b = 0; b < kVPlanes && borrow; b++) {
        const uint32_t x = v[(size_t)b * pitch];
        v[(size_t)b * pitch] = x ^ borrow;
        borrow &= ~x;
    }
}

// raw walk of one b-item: the item's indicator words of records 0 .. cnt-1 (rows are zero behind cnt up to the next
// multiple of 8) are added into its raw counter. ind_item = &ind[ind_index(sym, k, 0, ..)]; v = &bstate[item].
SNK_HD void ws_b_raw(const uint32_t* ind_item, uint32_t cnt, uint32_t* v, uint32_t pitch)
{
    uint32_t ones = v[0], twos = v[pitch], fours = v[2 * (size_t)pitch];
    for (uint32_t r = 0; r < cnt; r += 8) {
        const U4 a = load16(reinterpret_cast<const uint8_t*>(ind_item + r));
        const U4 b = load16(reinterpret_cast<const uint8_t*>(ind_item + r + 4));
        uint32_t ta, tb, fa, fb, e;
        csa(ta, ones, ones, a.x, a.y);
        csa(tb, ones, ones, a.z, a.w);
        csa(fa, twos, twos, ta, tb);
        csa(ta, ones, ones, b.x, b.y);
        csa(tb, ones, ones, b.z, b.w);
        csa(fb, twos, twos, ta, tb);
        csa(e, fours, fours, fa, fb);
        vripple_add(v, pitch, 3, e);
    }
    v[0] = ones; v[pitch] = twos; v[2 * (size_t)pitch] = fours;
}

// 32 bases of the record view that starts `shift` bases into the record, beginning at view position 32k
SNK_HD uint32_t ind_view_word(const uint32_t* ind_sym /* &ind[ind_index(sym, 0, r, ..)] */, int k, uint32_t shift, int nwd, uint32_t rp)
{
    const int kk = k + (int)(shift >> 5);
    const uint32_t lo = kk < nwd ? ind_sym[(size_t)kk * rp] : 0u;
    const uint32_t hi = kk + 1 < nwd ? ind_sym[(size_t)(kk + 1) * rp] : 0u;
    return funnel_r(lo, hi, shift & 31u);
}
// fast division of a tile byte address (< 2^21) by the row stride (16..1008): magic = floor(2^32 / stride) + 1
SNK_HD uint32_t stride_magic(uint32_t stride) { return (uint32_t)(0x100000000ull / stride) + 1u; }
SNK_HD uint32_t div_stride(uint32_t addr, uint32_t magic) { return (uint32_t)(((uint64_t)addr * magic) >> 32); }

// delta entries of one b-item (mate's list `dl`, nd entries): positions [start, end) of the record view at byte
// address addr = r * stride + shift are removed from (add flag: added to) the clean set.
// ind_mate_sym = &ind[ind_index(sym, 0, 0, ..)] of the item's mate; vd = &bstate[kVPlanes * pitch + item].
SNK_HD void ws_b_delta(const uint32_t* ind_mate_sym, const DeltaEnt* dl, uint32_t nd, int k, uint32_t stride, uint32_t magic, int nwd,
                       uint32_t rp, uint32_t* vd, uint32_t pitch)
{
    const int first = 32 * k;
    for (uint32_t e = 0; e < nd; e++) {
        const DeltaEnt d = dl[e];
        int hi = (int)(d.d0 & 0x3FFu) - first, lo = (int)(d.d1 & 0x3FFu) - first;
        if (hi <= 0 || lo >= 32) continue;
        if (lo < 0) lo = 0;
        if (hi > 32) hi = 32;
        const uint32_t mask = (hi >= 32 ? 0xFFFFFFFFu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);      // lo < 32
        const uint32_t addr = (d.d0 >> 10) & 0x1FFFFFu;
        const uint32_t r = div_stride(addr, magic), shift = addr - r * stride;
        const uint32_t x = ind_view_word(ind_mate_sym + r, k, shift, nwd, rp) & mask;
        if (!x) continue;
        if (d.d1 & kDeltaAdd) vripple_sub(vd, pitch, x);
        else vripple_add(vd, pitch, 0, x);
    }
}

// ------------------------------------------------------------------ flush of thread-owned items
// (64-bit wrap-around arithmetic: every sum of a table is exact and non-negative, a single owner's share of a
// clean count may be "negative")
SNK_HD void stat_add(unsigned long long* p, unsigned long long v)
{
#ifdef __CUDA_ARCH__
    atomicAdd(p, v);
#else
    *p += v;
#endif
}

// count of column i of a vertical counter
SNK_HD uint32_t vcolumn(const uint32_t* planes /* kVPlanes registers */, int i)
{
    uint32_t c = 0;
#pragma unroll
    for (int b = 0; b < kVPlanes; b++) c |= ((planes[b] >> i) & 1u) << b;
    return c;
}

// b-item (sym, k): raw table += raw, clean table += raw - delta; clears the item's state
// f_raw / f_clean = the mate's raw / clean file blocks of the slot
SNK_HD void ws_flush_b_item(uint32_t* v, uint32_t pitch, int sym, int k, unsigned long long* f_raw, unsigned long long* f_clean)
{
    uint32_t pr[kVPlanes], pd[kVPlanes];
    uint32_t any = 0;
#pragma unroll
    for (int b = 0; b < kVPlanes; b++) {
        pr[b] = v[(size_t)b * pitch]; pd[b] = v[(size_t)(kVPlanes + b) * pitch];
        any |= pr[b] | pd[b];
        v[(size_t)b * pitch] = 0u; v[(size_t)(kVPlanes + b) * pitch] = 0u;
    }
    if (!any) return;
    unsigned long long sum_r = 0, sum_c = 0;
    for (int i = 0; i < 32; i++) {
        const int pos = 32 * k + i;
        if (pos >= SNK_MAX_READ_LEN) break;
        const uint32_t vr = vcolumn(pr, i);
        const long long vdl = (long long)(int16_t)(uint16_t)vcolumn(pd, i);
        if (!vr && !vdl) continue;
        const unsigned long long vc = (unsigned long long)((long long)vr - vdl);
        const size_t cell = SNK_FILE_BS_OFF + (size_t)pos * 5 + (size_t)sym;
        if (vr) stat_add(&f_raw[cell], vr);
        if (vc) stat_add(&f_clean[cell], vc);
        sum_r += vr; sum_c += vc;
    }
    if (sum_r) { stat_add(&f_raw[SNK_FILE_GS_OFF + SNK_GS_A + sym], sum_r); stat_add(&f_raw[SNK_FILE_GS_OFF + SNK_GS_BASES], sum_r); }
    if (sum_c) { stat_add(&f_clean[SNK_FILE_GS_OFF + SNK_GS_A + sym], sum_c); stat_add(&f_clean[SNK_FILE_GS_OFF + SNK_GS_BASES], sum_c); }
}

// q-item x = (mate, item w of J positions): raw cells and delta cells (item x + nraw) of the owner-computes quality table
template <typename CounterT, int J>
SNK_HD void ws_flush_q_item(CounterT* qhist, uint32_t x, uint32_t nraw, uint32_t X, int w, int qb, unsigned long long* f_raw,
                            unsigned long long* f_clean)
{
    unsigned long long r20 = 0, r30 = 0, c20 = 0, c30 = 0;
    for (int q = 0; q < qb; q++) {
#pragma unroll
        for (int j = 0; j < J; j++) {
            const uint32_t e = qcell_index<J>((uint32_t)q, (uint32_t)j, x, X), ed = qcell_index<J>((uint32_t)q, (uint32_t)j, x + nraw, X);
            const uint32_t vr = qhist[e];
            const long long vdl = (long long)(int16_t)qhist[ed];
            if (!vr && !vdl) continue;
            qhist[e] = 0; qhist[ed] = 0;
            const unsigned long long vc = (unsigned long long)((long long)vr - vdl);
            const int pos = J * w + j;
            if (pos >= SNK_MAX_READ_LEN) continue;
            const size_t cell = SNK_FILE_QS_OFF + (size_t)pos * SNK_QBINS + (size_t)q;
            if (vr) stat_add(&f_raw[cell], vr);
            if (vc) stat_add(&f_clean[cell], vc);
            if (q >= 20) { r20 += vr; c20 += vc; }
            if (q >= 30) { r30 += vr; c30 += vc; }
        }
    }
    // the dump row (padding bytes of the rows) is only cleared
#pragma unroll
    for (int j = 0; j < J; j++) qhist[qcell_index<J>((uint32_t)qb, (uint32_t)j, x, X)] = 0;
    if (r20) stat_add(&f_raw[SNK_FILE_GS_OFF + SNK_GS_Q20], r20);
    if (r30) stat_add(&f_raw[SNK_FILE_GS_OFF + SNK_GS_Q30], r30);
    if (c20) stat_add(&f_clean[SNK_FILE_GS_OFF + SNK_GS_Q20], c20);
    if (c30) stat_add(&f_clean[SNK_FILE_GS_OFF + SNK_GS_Q30], c30);
}

// ------------------------------------------------------------------ shapes shared by the kernel, its launcher and the replay
constexpr int kWsScanWarps = 16;            // scan warps per CTA: groups of wpg warps, one tile per group at a time
constexpr int kWsHistWarps = 7;             // histogram warps per CTA
constexpr int kWsHistThreads = 32 * kWsHistWarps;
constexpr int kWsThreads = 32 * (kWsScanWarps + kWsHistWarps + 1);   // 24 warps = 6 warpgroups; the last warp is the producer
#ifndef SNK_WS_J
#define SNK_WS_J 4
#endif
constexpr int kWsJ = SNK_WS_J;              // positions per quality item
// J = 2 (twice as many, half as long quality items) passes the CPU replay but its tuning build aborted on the GPU in round 2
// and was not pursued: only 4 is supported.
static_assert(kWsJ == 4, "the warp-specialised kernel is validated for 4 positions per quality item only");
constexpr uint32_t kWsMaxStages = 12;
// Registers: a scheduler (SM sub-partition) owns 16 384 registers and gets every fourth warp, i.e. 4 scan warps and 2
// histogram / producer warps. The kernel is launched with 80 registers per thread (6 warps x 32 x 80 = 15 360); the two
// histogram / producer warpgroups then lower their share to kWsHistRegs and the four scan warpgroups raise theirs to
// kWsScanRegs with setmaxnreg. Only registers the CTA was launched with can be handed around (the scheduler's 1 024
// unallocated ones are not in the CTA's pool: asking for more blocks forever), so 4 x 88 + 2 x 64 = 6 x 80.
constexpr int kWsMaxRegs = 80;
constexpr int kWsScanRegs = 88;
constexpr int kWsHistRegs = 64;
static_assert(4 * kWsScanRegs + 2 * kWsHistRegs <= 6 * kWsMaxRegs, "setmaxnreg can only redistribute the launch allocation");
constexpr uint32_t kWsMaxStride = 256;      // longer rows stay on filter_kernel

struct WsShape {
    uint32_t wpg;          // scan warps per group
    uint32_t ngroups;      // kWsScanWarps / wpg
    uint32_t interleave;   // 1: group g = warps {g, g + ngroups, ...} (a scheduler's scan warps work on the same tile), 0: consecutive warps
    uint32_t R;            // tile capacity in reads (SE) or pairs (PE) = wpg * 32 / (2 * mates)
    uint32_t nstages;
    uint32_t nwd;          // plane words per read
    uint32_t rp;           // indicator row pitch
    uint32_t W;            // quality items per table
    uint32_t X;            // quality table pitch (items of all tables, rounded up to 32)
    uint32_t nq, nb;       // q-items (mates * W), b-items (mates * nwd * 5)
    uint32_t nb_pitch;
    int qb;
    // shared memory byte offsets: CTA-wide part, then nstages stage blocks of stage_bytes
    uint32_t off_qhist, off_bstate, off_ada, off_bars, off_stage0, stage_bytes, total;
    // inside a stage block
    uint32_t so_rows[2][2], so_ind[2], so_desc, so_delta, so_ctl;
};
// ctl words of a stage: ndelta[2], tile_slow, pad
constexpr uint32_t kWsCtlBytes = 16;

// wpg = 0: pick the largest group size whose pipeline (one stage per scanning group + one in the histogram
// warps + one being loaded, at least) fits the shared memory budget; returns false when nothing fits
inline bool ws_make_shape(int mates, uint32_t stride, int qb, uint32_t nada, uint32_t smem_limit, uint32_t wpg_want, WsShape& s)
{
    if (stride % 16 != 0 || stride == 0 || stride > kWsMaxStride) return false;
    const uint32_t rpw = 32u / (2u * (uint32_t)mates);            // reads (pairs) per scan warp
    s.nwd = (stride + 31) / 32;
    s.W = stride / kWsJ;
    s.nq = (uint32_t)mates * s.W;
    s.nb = (uint32_t)mates * s.nwd * kSyms;
    s.X = (2u * s.nq + 31) / 32 * 32;
    s.nb_pitch = (s.nb + 31) / 32 * 32;
    s.qb = qb;
    if (s.nq + s.nb > 2u * kWsHistThreads) return false;
    uint32_t o = 0;
    s.off_qhist = o; o += ((uint32_t)(qb + 1) * kWsJ * s.X * 2u + 15) / 16 * 16;
    s.off_bstate = o; o += bstate_words(s.nb_pitch) * 4u;
    s.off_ada = o; o += (nada * (uint32_t)sizeof(AdaHot) + 15) / 16 * 16;
    s.off_bars = o; o += 3u * kWsMaxStages * 8u;                  // full / scanned / empty, up to kWsMaxStages stages
    s.off_stage0 = (o + 127) / 128 * 128;
    if (s.off_stage0 >= smem_limit) return false;
    // candidates: group sizes 8, 4, 2, 1; a pipeline needs one stage per scanning group plus one for the histogram
    // warps / the load in flight. When the full set of groups does not fit, fewer groups scan (the other scan warps
    // exit): take the candidate that keeps most scan warps busy, larger tiles first.
    uint32_t best_active = 0;
    WsShape best = s;
    for (uint32_t wpg = wpg_want ? wpg_want : 8u; wpg >= 1; wpg /= 2) {
        if (kWsScanWarps % wpg) continue;
        WsShape c = s;
        c.wpg = wpg; c.R = wpg * rpw; c.rp = c.R + 4;
        uint32_t q = 0;
        for (int m = 0; m < 2; m++)
            for (int a = 0; a < 2; a++) { c.so_rows[m][a] = q; if (m < mates) q += c.R * stride; }
        q += 16;                                                  // unit_q_raw loads one row word past the last record
        for (int m = 0; m < 2; m++) { c.so_ind[m] = q; if (m < mates) q += ind_words((int)c.nwd, c.rp) * 4u; }
        c.so_desc = q; q += (uint32_t)mates * c.R * 4u;
        c.so_delta = (q + 7) / 8 * 8; q = c.so_delta + (uint32_t)mates * 2u * c.R * (uint32_t)sizeof(DeltaEnt);
        c.so_ctl = q; q += kWsCtlBytes;
        c.stage_bytes = (q + 127) / 128 * 128;
        uint32_t ns = (smem_limit - c.off_stage0) / c.stage_bytes;
        const uint32_t full_groups = kWsScanWarps / wpg;
        if (ns > kWsMaxStages) ns = kWsMaxStages;
        if (ns > full_groups + 2) ns = full_groups + 2;
        if (ns >= 2) {
            c.ngroups = ns - 1 < full_groups ? ns - 1 : full_groups;
            c.nstages = ns;
            c.total = c.off_stage0 + ns * c.stage_bytes;
            const uint32_t active = c.ngroups * wpg;
            if (active > best_active) { best_active = active; best = c; }
        }
        if (wpg_want) break;
    }
    if (!best_active) return false;
    s = best;
    return true;
}

} // namespace snkcore
