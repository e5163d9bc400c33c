// filter_kernel.cuh — the fused filter + statistics kernel (sm_100a).
//
// One persistent CTA per SM walks a contiguous range of tiles. A tile is up to R reads (SE) or R
// pairs (PE) of the fixed-stride SoA batch, staged into shared memory, then:
//   phase A  one thread per read: counters, predicates, adapter search, trim   (scan_read)
//   phase P  one thread per pair: discard cascade, result record, counters     (decide_pair/se)
//   phase B  one thread per (table, 4 positions): per-position base x quality histograms for the
//            raw and the clean records of the tile, owner-computes (no atomics) in shared memory
// Histograms stay in shared memory across tiles and are added to the slot's global tables with
// 64-bit atomics only when the CTA's slot changes and at kernel end.
#pragma once
#include <cuda_runtime.h>
#include "filter_core.cuh"

namespace snkcore {

typedef uint16_t QCounter;                 // shared-memory quality counters; flushed before they can wrap
constexpr uint32_t kQCounterMax = 65535;

// CTA shape: one thread per histogram item (J positions of one table); a tile holds
// threads / (mates * kNT) reads or pairs, so that phase A (kNT threads per read) and phase B (one
// thread per item) both keep every thread busy. PE150: 320 threads, 80-pair tiles, 2 CTAs per SM.
__host__ __device__ inline uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }
__host__ __device__ inline int hist_j(uint32_t) { return 4; }
// histogram items (J = 4 positions each) of all tables = threads busy in phase B; phase A runs kNT
// threads per read, so the CTA has kNT x that many threads and a tile of threads/(mates*kNT) reads
__host__ __device__ inline uint32_t hist_items(int mates, uint32_t stride) { return 2u * mates * (stride / 4); }
__host__ __device__ inline uint32_t cta_threads(int mates, uint32_t stride)
{
    uint32_t t = align_up(kNT * hist_items(mates, stride), 32);
    if (t > 1024) t = 1024;
    return t < 64 ? 64 : t;
}
template <int MAXC, int MATES, int J> struct KernelShape {
    static constexpr int kItems = kNT * 2 * MATES * (16 * MAXC) / J;
    static constexpr int kT = ((kItems + 31) / 32 * 32) < 64 ? 64 : ((kItems + 31) / 32 * 32);
    static constexpr int kMaxThreads = kT > 1024 ? 1024 : kT;
    // register budget: keep 640 threads (20 warps) resident per SM whenever the CTA is small enough
    static constexpr int kMinBlocks = kMaxThreads <= 320 ? 640 / kMaxThreads : 1;
};

struct KernelArgs {
    const uint8_t* seq[2];
    const uint8_t* qual[2];
    const uint16_t* len[2];
    snk_read_result* out[2];
    unsigned long long* stats;      // n_slots * SNK_SLOT_WORDS
    unsigned int* err_flags;        // sticky error bits
    unsigned long long* err_index;  // smallest global read index that raised an error
    const unsigned int* skip_word;  // optional: when (*skip_word & skip_mask) != 0 the launch does nothing
    unsigned int skip_mask;         //   (text path: a read did not fit the row stride, the batch is resubmitted)
    uint32_t stride;                // bytes per row
    uint32_t R;                     // tile capacity (reads or pairs)
    uint32_t items_w;               // histogram items per table = stride / J
    uint32_t X;                     // histogram row pitch = items rounded up to 32
    TileMap tm;
};

// shared memory layout (dynamic): [tile seq/qual rows][len][ReadInfo][desc][qhist][misc]
struct SmemPlan {
    uint32_t off_rows[2][2];   // [mate][0 seq, 1 qual]
    uint32_t off_len[2];
    uint32_t off_info[2];
    uint32_t off_desc;         // hist_desc words: [raw m0][raw m1][clean m0][clean m1], R each
    uint32_t off_qhist;
    uint32_t off_misc;
    uint32_t total;
};
constexpr uint32_t kMiscBytes = 8 * 8 + 4 * 8 * 8 + 16; // lastkey[8] + gsum[4][8] + the staging mbarrier
__host__ __device__ inline SmemPlan plan_smem(int mates, uint32_t R, uint32_t stride, uint32_t X, int qb)
{
    SmemPlan p;
    uint32_t o = 0;
    for (int m = 0; m < 2; m++)
        for (int a = 0; a < 2; a++) {
            p.off_rows[m][a] = o;
            if (m < mates) o += R * stride + 16;      // +16: hist_load may read one word past the last row
        }
    for (int m = 0; m < 2; m++) { p.off_len[m] = o; if (m < mates) o += align_up(R * 2, 16); }
    for (int m = 0; m < 2; m++) { p.off_info[m] = o; if (m < mates) o += align_up(R * (uint32_t)sizeof(ReadInfo), 16); }
    p.off_desc = o; o += align_up(4u * R * 4u, 16);
    p.off_qhist = o; o += align_up((uint32_t)qb * (uint32_t)hist_j(stride) * X * (uint32_t)sizeof(QCounter), 16);
    p.off_misc = o; o += kMiscBytes;
    p.total = o;
    return p;
}

#ifdef __CUDACC__

// ---- TMA (bulk async copy) staging of a tile: global -> shared, completion on an mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SNK_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SNK_DONE_%=;\n"
        "bra SNK_WAIT_%=;\n"
        "SNK_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ void report_error(const KernelArgs& A, uint32_t bits, uint64_t gi)
{
    atomicOr(A.err_flags, bits);
    atomicMin(A.err_index, (unsigned long long)gi);
}

__device__ __forceinline__ int file_of_tab(int mates, int tab) { return mates == 2 ? tab : (tab == 0 ? SNK_RAW1 : SNK_CLEAN1); }

// lane exchange inside a phase-A thread group (kNT == 2: partner = lane ^ 1)
template <int NW>
__device__ __forceinline__ void shfl_scan(ScanPart<NW>& d, const ScanPart<NW>& s, unsigned pm)
{
#pragma unroll
    for (int k = 0; k < NW; k++) {
        d.p0[k] = __shfl_xor_sync(pm, s.p0[k], 1); d.p1[k] = __shfl_xor_sync(pm, s.p1[k], 1);
        d.pn[k] = __shfl_xor_sync(pm, s.pn[k], 1); d.pl[k] = __shfl_xor_sync(pm, s.pl[k], 1);
    }
    d.accA = __shfl_xor_sync(pm, s.accA, 1); d.accN = __shfl_xor_sync(pm, s.accN, 1);
    d.accLow = __shfl_xor_sync(pm, s.accLow, 1); d.qsum = __shfl_xor_sync(pm, s.qsum, 1);
    d.viol = __shfl_xor_sync(pm, s.viol, 1); d.qviol = __shfl_xor_sync(pm, s.qviol, 1); d.qover = __shfl_xor_sync(pm, s.qover, 1);
}

// phase A for one read, executed by the kNT adjacent lanes of its group (h = lane's index in the group)
template <int MAXC>
__device__ __forceinline__ void scan_read_coop(const uint8_t* seq, const uint8_t* qual, int len, int mate, const DevParams& P,
                                               int h, unsigned pm, ReadInfo& R)
{
    static_assert(kNT == 2, "lane exchange below is written for pairs");
    constexpr int NW = (MAXC + 1) / 2;
    const bool want_planes = P.n_adapters[mate] > 0 || P.polyX_num != -1;
    ScanPart<NW> S, O;
    scan_chunks<MAXC>(seq, qual, len, P, h, want_planes, S);
    shfl_scan<NW>(O, S, pm);
    merge_scan(S, O);
    const bool polyx = P.polyX_num != -1 && polyx_hit(S, len, P.polyX_num);
    int ada_pos = -1;
    if (P.n_adapters[mate] > 0) {
        uint32_t p0[NW + 2], p1[NW + 2], pb[NW + 2];
#pragma unroll
        for (int k = 0; k < NW; k++) { p0[k] = S.p0[k]; p1[k] = S.p1[k]; pb[k] = S.pn[k] | S.pl[k] | ~plane_valid(len, k); }
        p0[NW] = p0[NW + 1] = 0; p1[NW] = p1[NW + 1] = 0; pb[NW] = pb[NW + 1] = 0xFFFFFFFFu;
        for (int i = 0; i < P.n_adapters[mate]; i++) {
            const AdapterDev& a = P.ada[mate][i];
            if (a.len == 0) continue;
            if (a.fast && len >= a.len - 1) {
                AdaPart ap, op;
                adapter_part<NW>(len, p0, p1, pb, a, h, ap);
                op.hit1 = __shfl_xor_sync(pm, ap.hit1, 1); op.pos2 = __shfl_xor_sync(pm, ap.pos2, 1); op.pos3 = __shfl_xor_sync(pm, ap.pos3, 1);
                merge_ada(ap, op);
                ada_pos = ada_result(ap);
            } else {
                int pos = (h == 0) ? adapter_pos_bytes(seq, len, a) : -1;
                const int other = __shfl_xor_sync(pm, pos, 1);
                ada_pos = (h == 0) ? pos : other;
            }
            if (ada_pos >= 0) break;
        }
    }
    TrimPart T, OT;
    trim_part(seq, qual, len, P, h, T);
    OT.hix = __shfl_xor_sync(pm, T.hix, 1); OT.tix = __shfl_xor_sync(pm, T.tix, 1); OT.ng = __shfl_xor_sync(pm, T.ng, 1);
    merge_trim(T, OT);
    finish_read<NW>(S, polyx, ada_pos, T, len, mate, P, R);
}

// add this CTA's histograms (shared-memory quality counters, per-thread base counters) to the
// slot's global tables and clear them
template <int J>
__device__ void flush_hist(const KernelArgs& A, const DevParams& P, int mates, QCounter* qhist, BaseCnt<J>& bc, uint32_t base_item,
                           unsigned long long* lastkey, unsigned long long* gsum, int slot)
{
    unsigned long long* S = A.stats + (size_t)slot * SNK_SLOT_WORDS;
    const uint32_t X = A.X, W = A.items_w;
    const int ntab = 2 * mates;                        // raw1,[raw2],clean1,[clean2] -> item groups
    // quality cells: entry e = (q*J + j)*X + x
    const uint32_t qent = (uint32_t)P.qb * (uint32_t)J * X;
    for (uint32_t e = threadIdx.x; e < qent; e += blockDim.x) {
        const uint32_t v = qhist[e];
        if (!v) continue;
        qhist[e] = 0;
        const uint32_t x = e % X, j = (e / X) % (uint32_t)J, q = e / ((uint32_t)J * X);
        const uint32_t tab = x / W, w = x % W;
        if ((int)tab >= ntab) continue;
        unsigned long long* F = S + SNK_SLOT_FILE_OFF(file_of_tab(mates, (int)tab));
        atomicAdd(&F[SNK_FILE_QS_OFF + (size_t)(J * w + j) * SNK_QBINS + q], (unsigned long long)v);
        if (q >= 20) atomicAdd(&gsum[tab * 8 + 6], (unsigned long long)v);
        if (q >= 30) atomicAdd(&gsum[tab * 8 + 7], (unsigned long long)v);
    }
    // base cells: the item whose base counters this thread holds (base_item = ~0u: none)
    {
        const uint32_t x = base_item;
        const uint32_t tab = x / W, w = x % W;
        if ((int)tab < ntab) {
            unsigned long long* F = S + SNK_SLOT_FILE_OFF(file_of_tab(mates, (int)tab));
            unsigned long long bases = 0;
#pragma unroll
            for (int b = 0; b < 5; b++) {
                unsigned long long sym = 0;
#pragma unroll
                for (int j = 0; j < J; j++) {
                    const uint32_t v = bc.v[b][j];
                    if (v) atomicAdd(&F[SNK_FILE_BS_OFF + (size_t)(J * w + j) * 5 + b], (unsigned long long)v);
                    sym += v;
                    bc.v[b][j] = 0;
                }
                if (sym) atomicAdd(&gsum[tab * 8 + b], sym);
                bases += sym;
            }
            if (bases) atomicAdd(&gsum[tab * 8 + 5], bases);
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int t = threadIdx.x >> 3, k = threadIdx.x & 7;      // 4 tables x 8 sums
        if (t < ntab) {
            unsigned long long* G = S + SNK_SLOT_FILE_OFF(file_of_tab(mates, t)) + SNK_FILE_GS_OFF;
            // gsum order: A,C,G,T,N,bases,q20,q30
            const int gs_index[8] = {SNK_GS_A, SNK_GS_C, SNK_GS_G, SNK_GS_T, SNK_GS_N, SNK_GS_BASES, SNK_GS_Q20, SNK_GS_Q30};
            const unsigned long long v = gsum[t * 8 + k];
            if (v) atomicAdd(&G[gs_index[k]], v);
            gsum[t * 8 + k] = 0;
            if (k == 0) {
                if (lastkey[t]) atomicMax(&G[SNK_GS_LAST_KEY], lastkey[t]);
                if (lastkey[4 + t]) atomicAdd(&G[SNK_GS_READS], lastkey[4 + t]);
                lastkey[t] = 0; lastkey[4 + t] = 0;
            }
        }
    }
    __syncthreads();
}

template <int MAXC, int MATES, int J>
__global__ void __launch_bounds__((KernelShape<MAXC, MATES, J>::kMaxThreads), (KernelShape<MAXC, MATES, J>::kMinBlocks))
filter_kernel(const __grid_constant__ DevParams P, const __grid_constant__ KernelArgs A)
{
    extern __shared__ __align__(16) uint8_t smem[];
    if (A.skip_word && (*A.skip_word & A.skip_mask)) return;
    const SmemPlan sp = plan_smem(MATES, A.R, A.stride, A.X, P.qb);
    QCounter* qhist = reinterpret_cast<QCounter*>(smem + sp.off_qhist);
    uint32_t* desc = reinterpret_cast<uint32_t*>(smem + sp.off_desc);
    // misc: lastkey[0..3] = max key per table, lastkey[4..7] = record counts per table, gsum[4][8]
    unsigned long long* lastkey = reinterpret_cast<unsigned long long*>(smem + sp.off_misc);
    unsigned long long* gsum = lastkey + 8;
    unsigned long long* stage_bar = gsum + 32;
    uint32_t stage_phase = 0;
    const int tid = threadIdx.x;
    const int lane = tid & 31;

    for (uint32_t e = tid; e < (sp.total - sp.off_qhist) / 4; e += blockDim.x) reinterpret_cast<uint32_t*>(qhist)[e] = 0;   // qhist + misc
    BaseCnt<J> bc;
#pragma unroll
    for (int b = 0; b < 5; b++)
#pragma unroll
        for (int j = 0; j < J; j++) bc.v[b][j] = 0;
    __syncthreads();
    if (tid == 0) {
        mbar_init(stage_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint32_t nt = A.tm.ntiles;
    const uint32_t t_begin = (uint32_t)((uint64_t)nt * blockIdx.x / gridDim.x);
    const uint32_t t_end = (uint32_t)((uint64_t)nt * (blockIdx.x + 1) / gridDim.x);
    int cur_slot = -1;
    uint32_t reads_in_hist = 0;                     // records counted since the last flush (u16 cells)
    const uint32_t W = A.items_w;
    const uint32_t nitems = 2u * MATES * W;         // <= blockDim.x / kNT
    // Phase B roles: the first `nitems` threads own the quality cells of one item each; when the CTA is
    // large enough a second set of threads (starting at X, warp aligned) owns the base counters of the
    // same items, so that both halves of the CTA work during phase B. Otherwise one thread does both.
    const bool split = blockDim.x >= A.X + nitems;
    const bool q_role = (uint32_t)tid < nitems;
    const uint32_t item = (split && (uint32_t)tid >= A.X) ? (uint32_t)tid - A.X : (uint32_t)tid;
    const bool b_role = split ? ((uint32_t)tid >= A.X && item < nitems) : q_role;
    const uint32_t my_tab = item / W, my_w = item % W;
    const bool my_item = q_role || b_role;
    const int my_m = (int)(my_tab % MATES);
    const bool my_clean = my_tab >= (uint32_t)MATES;
    // byte offsets into the quality table for hist_item_fast: cell(b, j) = cell0 + j*jstep + b*bstep
    const int q_jstep = (int)A.X * (int)sizeof(QCounter), q_bstep = J * q_jstep;
    const int q_cell0 = (int)item * (int)sizeof(QCounter) - P.phred * q_bstep;
    const uint32_t* my_desc = desc + (size_t)((my_clean ? 2 : 0) + my_m) * A.R;

    for (uint32_t t = t_begin; t < t_end; t++) {
        uint32_t start, cnt;
        tile_range(A.tm, t, &start, &cnt);
        const uint64_t g0 = A.tm.first + start;
        const int slot = slot_of(g0, (uint64_t)P.slot_block, P.n_slots);
        if (slot != cur_slot || reads_in_hist + cnt > kQCounterMax) {
            if (cur_slot >= 0) flush_hist<J>(A, P, MATES, qhist, bc, b_role ? item : 0xFFFFFFFFu, lastkey, gsum, cur_slot);
            cur_slot = slot;
            reads_in_hist = 0;
        }
        reads_in_hist += cnt;
        // ---- stage the tile: the rows of a tile are one contiguous block per array, so one elected thread
        // hands 2*MATES bulk copies to the TMA engine and everybody waits on the mbarrier they complete on
        const uint32_t row_bytes = cnt * A.stride;
        if (tid == 0) {
            mbar_expect_tx(stage_bar, 2u * MATES * row_bytes);
#pragma unroll
            for (int m = 0; m < MATES; m++) {
                bulk_g2s(smem + sp.off_rows[m][0], A.seq[m] + (size_t)start * A.stride, row_bytes, stage_bar);
                bulk_g2s(smem + sp.off_rows[m][1], A.qual[m] + (size_t)start * A.stride, row_bytes, stage_bar);
            }
        }
#pragma unroll
        for (int m = 0; m < MATES; m++) {
            uint16_t* sl = reinterpret_cast<uint16_t*>(smem + sp.off_len[m]);
            for (uint32_t i = tid; i < cnt; i += blockDim.x) sl[i] = A.len[m][start + i];
        }
        mbar_wait(stage_bar, stage_phase);
        stage_phase ^= 1u;
        __syncthreads();

        // ---- phase A: kNT adjacent lanes per read
        for (uint32_t i = tid; i < cnt * MATES * kNT; i += blockDim.x) {
            const uint32_t ir = i / kNT;
            const int h = (int)(i % kNT);
            const int m = (MATES == 2) ? (int)(ir / cnt) : 0;
            const uint32_t r = (MATES == 2) ? ir % cnt : ir;
            const unsigned pm = 3u << (lane & ~1);                  // my group's lanes
            const uint16_t* sl = reinterpret_cast<const uint16_t*>(smem + sp.off_len[m]);
            ReadInfo* info = reinterpret_cast<ReadInfo*>(smem + sp.off_info[m]);
            int len = sl[r];
            if (len > (int)A.stride) len = (int)A.stride;
            ReadInfo ri;
            if (len <= 0) {             // "Error:empty sequence" (read_filter.cpp:250)
                ri.len = 0; ri.head_cut = 0; ri.clean_len = 0; ri.head_hdcut = ri.head_lqcut = ri.tail_hdcut = ri.tail_lqcut = ri.adacut_pos = -1;
                ri.flags = RF_BAD_BASE;
            } else {
                scan_read_coop<MAXC>(smem + sp.off_rows[m][0] + (size_t)r * A.stride, smem + sp.off_rows[m][1] + (size_t)r * A.stride,
                                     len, m, P, h, pm, ri);
            }
            if (h == 0) info[r] = ri;
        }
        __syncthreads();

        // ---- phase P: one thread per pair / read
        for (uint32_t r = tid; r < align_up(cnt, 32); r += blockDim.x) {
            const bool live = r < cnt;
            int cat = SNK_DROP_EMPTY, mask = 0, fsb = -1;
            ReadInfo a, b;
            const uint64_t gi = g0 + r;
            if (live) {
                a = reinterpret_cast<const ReadInfo*>(smem + sp.off_info[0])[r];
                if (MATES == 2) {
                    b = reinterpret_cast<const ReadInfo*>(smem + sp.off_info[1])[r];
                    cat = decide_pair(P, a, b, &mask, &fsb);
                    uint32_t err = 0;
                    if ((a.flags | b.flags) & RF_BAD_BASE) err |= ERR_BAD_BASE;
                    if ((a.flags | b.flags) & RF_BAD_QUAL) err |= ERR_BAD_QUAL;
                    if (cat == SNK_DROP_LOWQ && ((a.flags | b.flags) & RF_LOWQ_GT1)) err |= ERR_LOWQ_RATIO;
                    if (err) report_error(A, err, gi);
                } else {
                    cat = decide_se(P, a, &fsb);
                    mask = cat ? 1 : 0;
                    uint32_t err = 0;
                    if (a.flags & RF_BAD_BASE) err |= ERR_BAD_BASE;
                    if (a.flags & RF_BAD_QUAL) err |= ERR_BAD_QUAL;
                    if (err) report_error(A, err, gi);
                }
                const uint32_t row0 = r * A.stride;
                desc[0 * A.R + r] = hist_desc(a.len, row0, a.flags & RF_QSLOW);
                desc[2 * A.R + r] = cat == SNK_KEEP ? hist_desc(a.clean_len, row0 + (uint32_t)a.head_cut, a.flags & RF_QSLOW) : 0u;
                if (MATES == 2) {
                    desc[1 * A.R + r] = hist_desc(b.len, row0, b.flags & RF_QSLOW);
                    desc[3 * A.R + r] = cat == SNK_KEEP ? hist_desc(b.clean_len, row0 + (uint32_t)b.head_cut, b.flags & RF_QSLOW) : 0u;
                }
                unsigned long long* S = A.stats + (size_t)slot * SNK_SLOT_WORDS;
                if (fsb >= 0) {
                    atomicAdd(&S[fsb], 1ull);
                    if (MATES == 2) {
                        if (mask & 1) atomicAdd(&S[fsb + 1], 1ull);
                        if (mask & 2) atomicAdd(&S[fsb + 2], 1ull);
                        if (mask == 3) atomicAdd(&S[fsb + 3], 1ull);
                    }
                }
#pragma unroll
                for (int m = 0; m < MATES; m++) {
                    const ReadInfo& x = m ? b : a;
                    // snk_read_result packed into one 8-byte store (little endian field order)
                    const unsigned long long packed = (unsigned long long)(uint16_t)x.head_cut |
                        ((unsigned long long)(uint16_t)x.clean_len << 16) | ((unsigned long long)(uint8_t)cat << 32) |
                        ((unsigned long long)(uint8_t)mask << 40) | ((unsigned long long)(uint16_t)x.adacut_pos << 48);
                    reinterpret_cast<unsigned long long*>(A.out[m])[start + r] = packed;
                    // trimming-position tables: raw record (cut ints only if copied back, raw_length 0) and clean copy
                    const int which = (MATES == 2) ? m : 2;
                    int hf, tf;
                    if (P.cutback) {
                        trim_stat_indices(which, x.len, 0, x.head_hdcut, x.head_lqcut, x.tail_hdcut, x.tail_lqcut, x.adacut_pos, &hf, &tf);
                        unsigned long long* T = S + SNK_SLOT_FILE_OFF(m == 0 ? SNK_RAW1 : SNK_RAW2) + SNK_FILE_TS_OFF;
                        if (hf >= 0) atomicAdd(&T[hf], 1ull);
                        if (tf >= 0) atomicAdd(&T[tf], 1ull);
                    }
                    if (cat == SNK_KEEP) {
                        trim_stat_indices(which, x.clean_len, x.len, x.head_hdcut, x.head_lqcut, x.tail_hdcut, x.tail_lqcut, x.adacut_pos, &hf, &tf);
                        unsigned long long* T = S + SNK_SLOT_FILE_OFF(m == 0 ? SNK_CLEAN1 : SNK_CLEAN2) + SNK_FILE_TS_OFF;
                        if (hf >= 0) atomicAdd(&T[hf], 1ull);
                        if (tf >= 0) atomicAdd(&T[tf], 1ull);
                    }
                }
            }
            // last record keys + record counts, one shared atomic per warp and table
            const unsigned kept = __ballot_sync(0xFFFFFFFFu, live && cat == SNK_KEEP);
            const unsigned lanes = __ballot_sync(0xFFFFFFFFu, live);
            if (lanes && lane == 31 - __clz(lanes)) {      // last live read of this warp's group
                atomicMax(&lastkey[0], ((gi + 1) << 16) | (unsigned long long)(uint16_t)a.len);
                atomicAdd(&lastkey[4 + 0], (unsigned long long)__popc(lanes));
                if (MATES == 2) {
                    atomicMax(&lastkey[1], ((gi + 1) << 16) | (unsigned long long)(uint16_t)b.len);
                    atomicAdd(&lastkey[4 + 1], (unsigned long long)__popc(lanes));
                }
            }
            if (kept && lane == 31 - __clz(kept)) {
                atomicMax(&lastkey[MATES], ((gi + 1) << 16) | (unsigned long long)(uint16_t)a.clean_len);
                atomicAdd(&lastkey[4 + MATES], (unsigned long long)__popc(kept));
                if (MATES == 2) {
                    atomicMax(&lastkey[MATES + 1], ((gi + 1) << 16) | (unsigned long long)(uint16_t)b.clean_len);
                    atomicAdd(&lastkey[4 + MATES + 1], (unsigned long long)__popc(kept));
                }
            }
        }
        __syncthreads();

        // ---- phase B: per-position histograms, owner computes. Each role has its own tight loop; the
        // descriptor carries the record's byte address, so a trip is: load descriptor, load word, update.
        if (my_item) {
            const uint8_t* rows_s = smem + sp.off_rows[my_m][0];
            const uint8_t* rows_q = smem + sp.off_rows[my_m][1];
            const int first_pos = J * (int)my_w;
            BaseAcc acc = {0, 0, 0, 0, 0};
            uint32_t err = 0;
            bool any_slow = false;
            if (q_role) {
#pragma unroll 2
                for (uint32_t r = 0; r < cnt; r++) {
                    const uint32_t d = my_desc[r];
                    const int nvalid = (int)(d & 0x3FFu) - first_pos;
                    if (nvalid <= 0) continue;
                    if (d & 0x80000000u) { any_slow = true; continue; }
                    qual_update_fast<QCounter, J>(hist_load_word<J>(rows_q, (int)((d >> 10) & 0x1FFFFFu), (int)my_w), nvalid,
                                                  reinterpret_cast<uint8_t*>(qhist), q_cell0, q_jstep, q_bstep);
                }
            }
            if (b_role) {
#pragma unroll 2
                for (uint32_t r = 0; r < cnt; r++) {
                    const uint32_t d = my_desc[r];
                    const int nvalid = (int)(d & 0x3FFu) - first_pos;
                    if (nvalid <= 0) continue;
                    base_update<J>(hist_load_word<J>(rows_s, (int)((d >> 10) & 0x1FFFFFu), (int)my_w), nvalid, acc);
                    if (cnt > 255 && (r & 127) == 127) base_acc_spill<J>(acc, bc);     // packed 8-bit lanes must not wrap
                }
            }
            if (any_slow) {
                // checked path for reads with qualities outside the shared-memory bins (rare): qualities only,
                // the bases were already counted above
                BaseAcc dummy = {0, 0, 0, 0, 0};
                unsigned long long* file_base = A.stats + (size_t)slot * SNK_SLOT_WORDS + SNK_SLOT_FILE_OFF(file_of_tab(MATES, (int)my_tab));
                for (uint32_t r = 0; r < cnt; r++) {
                    const uint32_t d = my_desc[r];
                    if (!(d & 0x80000000u)) continue;
                    const int addr = (int)((d >> 10) & 0x1FFFFFu);
                    err |= hist_item<QCounter, J>(rows_s, rows_q, addr, (int)(d & 0x3FFu), (int)my_w, P.phred, P.qb, dummy, qhist + item, (int)A.X,
                                                  file_base, false, true);
                }
            }
            base_acc_spill<J>(acc, bc);
            if (err) report_error(A, err, g0);
        }
        __syncthreads();
    }
    if (cur_slot >= 0) flush_hist<J>(A, P, MATES, qhist, bc, b_role ? item : 0xFFFFFFFFu, lastkey, gsum, cur_slot);
}

#endif // __CUDACC__

} // namespace snkcore
