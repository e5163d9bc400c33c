// filter_kernel.cuh — the fused filter + statistics kernel (sm_100a).
//
// One persistent CTA per SM walks a contiguous range of tiles. A tile is up to R reads (SE) or R
// pairs (PE) of the fixed-stride SoA batch, staged into shared memory, then:
//   phase A  one thread per read: counters, predicates, adapter search, trim   (scan_read)
//   phase P  one thread per read: discard cascade, result record, counters, delta list  (decide_pair/se)
//   phase B  work units per (table, 4 positions): per-position base x quality histograms, owner-computes
//            (no atomics) in shared memory: raw records counted directly, clean = raw - delta entries
// Histograms stay in shared memory across tiles and are added to the slot's global tables with
// 64-bit atomics only when the CTA's slot changes and at kernel end.
#pragma once
#include <cuda_runtime.h>
#include "filter_core.cuh"
#include "ws_core.cuh"

namespace snkcore {

typedef uint16_t QCounter;                 // shared-memory quality counters; flushed before they can wrap
constexpr uint32_t kQCounterMax = 32767;   // delta cells are read back as signed 16-bit values

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

// shared memory layout (dynamic): [tile seq/qual rows][len][ReadInfo][desc][delta][qhist][adapters][misc]
struct SmemPlan {
    uint32_t off_rows[2][2];   // [mate][0 seq, 1 qual]
    uint32_t off_len[2];
    uint32_t off_info[2];
    uint32_t off_desc;         // hist_desc words of the raw records: [m0][m1], R each (checked path only)
    uint32_t off_delta;        // DeltaEnt lists: [m0][m1], 2R entries each
    uint32_t off_qhist;        // (qb + 1) * J rows of X cells; row group qb is the padding dump bin
    uint32_t off_ada;          // AdaHot[ada_slots]: the sweep constants of every adapter, mate 0's first, built at kernel start
    uint32_t off_misc;
    uint32_t total;
};
// misc: lastkey[8] + gsum[4][8] + the staging mbarrier (8) + ndelta[2] + tile_slow + pad
constexpr uint32_t kMiscBytes = 8 * 8 + 4 * 8 * 8 + 8 + 2 * 4 + 4 + 12;
// nada = adapters staged in shared memory (see ada_slots)
__host__ __device__ inline SmemPlan plan_smem(int mates, uint32_t R, uint32_t stride, uint32_t X, int qb, uint32_t nada)
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
    p.off_desc = o; o += align_up((uint32_t)mates * R * 4u, 16);
    p.off_delta = o; o += align_up((uint32_t)mates * 2u * R * (uint32_t)sizeof(DeltaEnt), 16);
    p.off_qhist = o; o += align_up((uint32_t)(qb + 1) * (uint32_t)hist_j(stride) * X * (uint32_t)sizeof(QCounter), 16);
    p.off_ada = o; o += align_up(nada * (uint32_t)sizeof(AdaHot), 16);
    p.off_misc = o; o += kMiscBytes;
    p.total = o;
    return p;
}

// adapters of mate 0 occupy AdaHot slots [0, max(1,n0)), those of mate 1 follow
__host__ __device__ inline uint32_t ada_first_slot(const int32_t* n_adapters, int mate) { return mate ? (uint32_t)(n_adapters[0] > 1 ? n_adapters[0] : 1) : 0u; }
__host__ __device__ inline uint32_t ada_slots(const int32_t* n_adapters) { return ada_first_slot(n_adapters, 1) + (uint32_t)(n_adapters[1] > 1 ? n_adapters[1] : 1); }

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
    d.low128 = __shfl_xor_sync(pm, s.low128, 1); d.qsum = __shfl_xor_sync(pm, s.qsum, 1);
    d.viol = __shfl_xor_sync(pm, s.viol, 1); d.qbad = __shfl_xor_sync(pm, s.qbad, 1);
}

// phase A for one read, executed by the kNT adjacent lanes of its group (h = lane's index in the group)
// ada0 = shared-memory AdaHot array of the mate's adapters (the compiler re-derives the parameter-space
// address of a dynamically indexed adapter at every use; a shared copy costs one pointer register)
// ind != nullptr (warp-specialised kernel): the read's indicator planes are stored for the base-count items
// (record r of the tile; ws_core.cuh)
template <int MAXC, int MATES>
__device__ __forceinline__ void scan_read_coop(uint8_t* seq, uint8_t* qual, int len, int nchunks, int mate, const DevParams& P,
                                               const AdaHot* ada0, int h, unsigned pm, ReadInfo& R, uint32_t* ind = nullptr,
                                               uint32_t r = 0, int nwd = 0, uint32_t rp = 0)
{
    static_assert(kNT == 2, "lane exchange below is written for pairs");
    constexpr int NW = (MAXC + 1) / 2;
    ScanPart<NW> S, O;
    scan_chunks<MAXC>(seq, qual, len, nchunks, P, h, S);
    shfl_scan<NW>(O, S, pm);
    merge_scan(S, O);
    if (ind) store_indicators<NW>(S, len, h, ind, r, nwd, rp);
    const bool polyx = P.polyX_num != -1 && polyx_hit(S, len, P.polyX_num);
    int ada_pos = -1;
    bool has5 = false;
    if (MATES == 1 && P.srna) {
        // filtersRNA (single-end only: the paired kernels carry none of its code): the group's first lane aligns the
        // 3' adapter, the second the 5' adapter
        uint32_t q0[NW + 2], q1[NW + 2], qn[NW + 2], ql[NW + 2];
#pragma unroll
        for (int k = 0; k < NW + 2; k++) { q0[k] = k < NW ? S.p0[k] : 0u; q1[k] = k < NW ? S.p1[k] : 0u; qn[k] = k < NW ? S.pn[k] : 0u; ql[k] = k < NW ? S.pl[k] : 0u; }
        const int v = srna_find<NW>(seq, q0, q1, qn, ql, len, P, h == 0 ? 1 : 0);
        const int o = __shfl_xor_sync(pm, v, 1);
        ada_pos = (h == 0) ? v : o;
        has5 = ((h == 0) ? o : v) != 0;
    } else if (P.n_adapters[mate] > 0) {
        uint32_t p0[NW + 2], p1[NW + 2], pb[NW + 2];
#pragma unroll
        for (int k = 0; k < NW; k++) { p0[k] = S.p0[k]; p1[k] = S.p1[k]; pb[k] = S.pn[k] | S.pl[k] | ~plane_valid(len, k); }
        p0[NW] = p0[NW + 1] = 0; p1[NW] = p1[NW + 1] = 0; pb[NW] = pb[NW + 1] = 0xFFFFFFFFu;
        for (int i = 0; i < P.n_adapters[mate]; i++) {
            const AdaHot& a = ada0[i];
            if (a.len == 0) continue;
            if (a.fast && len >= a.len - 1) {
                AdaPart ap, op;
                adapter_part<NW>(len, p0, p1, pb, a, h, ap);
                op.hit1 = __shfl_xor_sync(pm, ap.hit1, 1); op.pos2 = __shfl_xor_sync(pm, ap.pos2, 1); op.pos3 = __shfl_xor_sync(pm, ap.pos3, 1);
                merge_ada(ap, op);
                ada_pos = ada_result(ap);
            } else {
                int pos = (h == 0) ? adapter_pos_bytes(seq, len, P.ada[mate][i]) : -1;
                const int other_pos = __shfl_xor_sync(pm, pos, 1);
                ada_pos = (h == 0) ? pos : other_pos;
            }
            if (ada_pos >= 0) break;
        }
    }
    const int cur = srna_cut_len(P, ada_pos, len);
    TrimPart T, OT;
    trim_part(seq, qual, len, cur, P, h, T);
    OT.hix = __shfl_xor_sync(pm, T.hix, 1); OT.tix = __shfl_xor_sync(pm, T.tix, 1); OT.ng = __shfl_xor_sync(pm, T.ng, 1);
    merge_trim(T, OT);
    uint16_t contam = 0;
    if (!P.srna && (P.n_contams[mate] > 0 || P.n_gcontams > 0)) {      // uncommon options: the group's first lane runs the byte-wise searches
        const int v = (h == 0) ? (int)contam_flags(seq, len, mate, P) : 0;
        contam = (uint16_t)(v | __shfl_xor_sync(pm, v, 1));
    }
    finish_read<NW>(S, S.qbad && qual_violation(qual, len, P.phred), polyx, ada_pos, has5, contam, cur, T, len, mate, P, R);
}

// add this CTA's histograms (shared-memory quality counters, per-thread base counters) to the
// slot's global tables and clear them. Raw cells count all records, delta cells "removed - added":
// raw table += raw, clean table += raw - delta (64-bit wrap-around arithmetic, the sums are exact).
template <int J>
__device__ void flush_hist(const KernelArgs& A, const DevParams& P, int mates, QCounter* qhist, BaseCnt<J>& bc, uint32_t base_item,
                           unsigned long long* lastkey, unsigned long long* gsum, int slot)
{
    unsigned long long* S = A.stats + (size_t)slot * SNK_SLOT_WORDS;
    const uint32_t X = A.X, W = A.items_w;
    const uint32_t nraw = (uint32_t)mates * W;         // raw items; the delta item of raw item x is x + nraw
    // quality cells (qcell_index). A thread keeps one raw item x (hence one table) and strides
    // over the (q, j) rows, so the q20/q30 totals stay in registers until one 32-bit shared atomic each
    // (all per-interval sums fit 32 bits: at most kQCounterMax records x 1008 positions).
    const uint32_t qrows = (uint32_t)P.qb * (uint32_t)J;
    const uint32_t rows_par = blockDim.x / nraw;            // >= 2: the CTA has at least 2*nraw threads
    if (threadIdx.x < rows_par * nraw) {
        const uint32_t x = threadIdx.x % nraw, tab = x / W, w = x % W;
        unsigned long long* FR = S + SNK_SLOT_FILE_OFF(file_of_tab(mates, (int)tab)) + SNK_FILE_QS_OFF + (size_t)(J * w) * SNK_QBINS;
        unsigned long long* FC = S + SNK_SLOT_FILE_OFF(file_of_tab(mates, (int)tab + mates)) + SNK_FILE_QS_OFF + (size_t)(J * w) * SNK_QBINS;
        uint32_t r20 = 0, r30 = 0, c20 = 0, c30 = 0;
        for (uint32_t qj = threadIdx.x / nraw; qj < qrows; qj += rows_par) {
            const uint32_t j = qj % (uint32_t)J, q = qj / (uint32_t)J;
            const uint32_t e = qcell_index<J>(q, j, x, X), ed = qcell_index<J>(q, j, x + nraw, X);
            const uint32_t vr = qhist[e];
            const int vd = (int)(int16_t)qhist[ed];
            if (!vr && !vd) continue;
            qhist[e] = 0; qhist[ed] = 0;
            const uint32_t vc = (uint32_t)((int)vr - vd);      // records of the clean set in this cell: never negative
            if (vr) atomicAdd(&FR[(size_t)j * SNK_QBINS + q], (unsigned long long)vr);
            if (vc) atomicAdd(&FC[(size_t)j * SNK_QBINS + q], (unsigned long long)vc);
            if (q >= 20) { r20 += vr; c20 += vc; }
            if (q >= 30) { r30 += vr; c30 += vc; }
        }
        uint32_t* g32 = reinterpret_cast<uint32_t*>(gsum);  // low words (little endian); high words stay 0
        if (r20) atomicAdd(&g32[2 * (tab * 8 + 6)], r20);
        if (r30) atomicAdd(&g32[2 * (tab * 8 + 7)], r30);
        if (c20) atomicAdd(&g32[2 * ((tab + mates) * 8 + 6)], c20);
        if (c30) atomicAdd(&g32[2 * ((tab + mates) * 8 + 7)], c30);
    }
    // base cells: the raw item whose base counters this thread holds (base_item = ~0u: none)
    {
        const uint32_t x = base_item;
        const uint32_t tab = x / W, w = x % W;
        if ((int)tab < mates) {
            unsigned long long* FR = S + SNK_SLOT_FILE_OFF(file_of_tab(mates, (int)tab));
            unsigned long long* FC = S + SNK_SLOT_FILE_OFF(file_of_tab(mates, (int)tab + mates));
            uint32_t* g32 = reinterpret_cast<uint32_t*>(gsum);
            uint32_t bases_r = 0, bases_c = 0;
#pragma unroll
            for (int b = 0; b < 5; b++) {
                uint32_t sym_r = 0, sym_c = 0;
#pragma unroll
                for (int j = 0; j < J; j++) {
                    const uint32_t vr = (bc.raw[b][j / 2] >> (16 * (j & 1))) & 0xFFFFu;
                    const int vd = (int)((bc.del[b][j / 2] >> (16 * (j & 1))) & 0xFFFFu) - 0x8000;
                    // this thread's share of the clean count may be negative when two b-units split an item's
                    // records (one saw the record, the other its delta entry): 64-bit wrap-around add
                    const long long vc = (long long)vr - (long long)vd;
                    const size_t cell = SNK_FILE_BS_OFF + (size_t)(J * w + j) * 5 + b;
                    if (vr) atomicAdd(&FR[cell], (unsigned long long)vr);
                    if (vc) atomicAdd(&FC[cell], (unsigned long long)vc);
                    sym_r += vr; sym_c += (uint32_t)vc;
                }
                if (sym_r) atomicAdd(&g32[2 * (tab * 8 + b)], sym_r);
                if (sym_c) atomicAdd(&g32[2 * ((tab + mates) * 8 + b)], sym_c);      // mod 2^32: the table's total is non-negative
                bases_r += sym_r; bases_c += sym_c;
            }
            if (bases_r) atomicAdd(&g32[2 * (tab * 8 + 5)], bases_r);
            if (bases_c) atomicAdd(&g32[2 * ((tab + mates) * 8 + 5)], bases_c);
            base_cnt_reset<J>(bc);
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int t = threadIdx.x >> 3, k = threadIdx.x & 7;      // 4 tables x 8 sums
        if (t < 2 * mates) {
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
    const SmemPlan sp = plan_smem(MATES, A.R, A.stride, A.X, P.qb, ada_slots(P.n_adapters));
    QCounter* qhist = reinterpret_cast<QCounter*>(smem + sp.off_qhist);
    uint32_t* desc = reinterpret_cast<uint32_t*>(smem + sp.off_desc);
    DeltaEnt* dlist = reinterpret_cast<DeltaEnt*>(smem + sp.off_delta);
    // misc: lastkey[0..3] = max key per table, lastkey[4..7] = record counts per table, gsum[4][8]
    unsigned long long* lastkey = reinterpret_cast<unsigned long long*>(smem + sp.off_misc);
    unsigned long long* gsum = lastkey + 8;
    unsigned long long* stage_bar = gsum + 32;
    uint32_t* ndelta = reinterpret_cast<uint32_t*>(stage_bar + 1);     // [2] entries in each mate's delta list
    uint32_t* tile_slow = ndelta + 2;                                    // some record of the tile needs the checked path
    uint32_t stage_phase = 0;
    const int tid = threadIdx.x;
    const int lane = tid & 31;

    for (uint32_t e = tid; e < (sp.total - sp.off_qhist) / 4; e += blockDim.x) reinterpret_cast<uint32_t*>(qhist)[e] = 0;   // qhist + misc
    BaseCnt<J> bc;
    base_cnt_reset<J>(bc);
    __syncthreads();
    for (uint32_t e = tid; e < ada_slots(P.n_adapters); e += blockDim.x) {
        const int m = e >= ada_first_slot(P.n_adapters, 1) ? 1 : 0;
        make_ada_hot(P.ada[m][e - ada_first_slot(P.n_adapters, m)], reinterpret_cast<AdaHot*>(smem + sp.off_ada)[e]);
    }
    if (tid == 0) {
        mbar_init(stage_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint32_t nt = A.tm.ntiles;
    const uint32_t t_begin = (uint32_t)((uint64_t)nt * blockIdx.x / gridDim.x);
    const uint32_t t_end = (uint32_t)((uint64_t)nt * (blockIdx.x + 1) / gridDim.x);
    int cur_slot = -1;
    uint32_t reads_in_hist = 0;                     // records counted since the last flush (16-bit cells)
    const uint32_t W = A.items_w;
    const int nchunks = (int)(A.stride / 16);
    // Phase B roles (work units of filter_core.cuh). nraw raw items = MATES tables x W items; the delta
    // item of raw item x is x + nraw. "wide" CTAs (4*nraw threads, the size phase A wants) run two
    // q-units per item (half of the J sub-positions each) on threads [0, 2*nraw) and two b-units per item
    // (every other record each) on threads [2*nraw, 4*nraw); CTAs capped at 1024 threads run one unit
    // per item.
    const uint32_t nraw = (uint32_t)MATES * W;
    const bool wide = blockDim.x >= 4u * nraw;
    const uint32_t n_units = wide ? 2u * nraw : nraw;
    const uint32_t b_first = wide ? 2u * nraw : ((blockDim.x >= align_up(nraw, 32) + nraw) ? align_up(nraw, 32) : 0u);
    const bool q_role = (uint32_t)tid < n_units;
    const bool b_role = (uint32_t)tid >= b_first && (uint32_t)tid < b_first + n_units;
    const uint32_t unit = b_role ? (uint32_t)tid - b_first : (uint32_t)tid;     // a thread with both roles has b_first == 0
    const uint32_t item = wide ? unit % nraw : unit;                             // raw item
    const uint32_t half = wide ? unit / nraw : 0u;                               // q: which J/2 sub-positions; b: record parity
    const uint32_t my_w = item % W;
    const int my_m = (int)(item / W) % MATES;
    const bool my_item = q_role || b_role;
    // byte offsets into the quality table: cell(b, j) = cell0 + qj_off(j, rowstep) + b*bstep (filter_core.cuh)
    const int q_jstep = (int)A.X * 2 * (int)sizeof(QCounter), q_bstep = (J / 2) * q_jstep;      // q_jstep = rowstep
    const int q_cell0 = (int)item * 2 * (int)sizeof(QCounter) - P.phred * q_bstep;
    const int q_cell0_del = q_cell0 + (int)nraw * 2 * (int)sizeof(QCounter);
    const uint32_t* my_desc = desc + (size_t)my_m * A.R;
    const DeltaEnt* my_dlist = dlist + (size_t)my_m * 2u * A.R;
    const uint32_t flush_item = b_role ? item : 0xFFFFFFFFu;

    for (uint32_t t = t_begin; t <= t_end; t++) {
        // (the trip t == t_end only flushes: one inlined copy of flush_hist instead of two)
        uint32_t start = 0, cnt = 0;
        if (t < t_end) tile_range(A.tm, t, &start, &cnt);
        const uint64_t g0 = A.tm.first + start;
        const int slot = t < t_end ? slot_of(g0, (uint64_t)P.slot_block, P.n_slots) : -1;
        if (slot != cur_slot || reads_in_hist + cnt > kQCounterMax) {
            if (cur_slot >= 0) flush_hist<J>(A, P, MATES, qhist, bc, flush_item, lastkey, gsum, cur_slot);
            cur_slot = slot;
            reads_in_hist = 0;
        }
        if (t == t_end) break;
        reads_in_hist += cnt;
        // ---- stage the tile: the rows of a tile are one contiguous block per array, so one elected thread
        // hands 2*MATES bulk copies to the TMA engine and everybody waits on the mbarrier they complete on
        const uint32_t row_bytes = cnt * A.stride;
        if (tid == 0) {
            // the previous tile's rows were written (padding normalised) and read through the generic proxy: order those
            // accesses before the async-proxy writes of the bulk copies into the same bytes
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(stage_bar, 2u * MATES * row_bytes);
#pragma unroll
            for (int m = 0; m < MATES; m++) {
                bulk_g2s(smem + sp.off_rows[m][0], A.seq[m] + (size_t)start * A.stride, row_bytes, stage_bar);
                bulk_g2s(smem + sp.off_rows[m][1], A.qual[m] + (size_t)start * A.stride, row_bytes, stage_bar);
            }
            ndelta[0] = 0; ndelta[1] = 0; *tile_slow = 0;
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
            const uint32_t len_word = sl[r];
            int len = (int)(len_word & SNK_LEN_MASK);
            if (len > (int)A.stride || len > SNK_MAX_READ_LEN) {      // not a row of this batch / no table row behind READ_MAX_LEN
                if (h == 0) report_error(A, ERR_BAD_LEN, g0 + r);
                len = 0;                                               // handled like an empty row (checked path, nothing counted)
            }
            ReadInfo ri;
            if (len <= 0) {             // "Error:empty sequence" (read_filter.cpp:250)
                ri.len = 0; ri.head_cut = 0; ri.clean_len = 0; ri.head_hdcut = ri.head_lqcut = ri.tail_hdcut = ri.tail_lqcut = ri.adacut_pos = -1;
                ri.flags = RF_BAD_BASE | RF_QSLOW;      // row left as staged: not safe for the unchecked histogram walk
            } else {
                scan_read_coop<MAXC, MATES>(smem + sp.off_rows[m][0] + (size_t)r * A.stride, smem + sp.off_rows[m][1] + (size_t)r * A.stride,
                                     len, nchunks, m, P, reinterpret_cast<const AdaHot*>(smem + sp.off_ada) + ada_first_slot(P.n_adapters, m), h, pm, ri);
            }
            if (h == 0) {
                ri.flags |= pre_flags(len_word);
                info[r] = ri;
                if (ri.flags & RF_QSLOW) *tile_slow = 1u;
            }
        }
        __syncthreads();

        // ---- phase P: one thread per read (pair r, mate m on adjacent lanes): discard cascade (both lanes of a
        // pair evaluate it), result record, counters, trim-position tables, raw descriptor, delta entries.
        // Wide CTAs run it on the b-unit warps only: the q-unit warps go straight on to the raw histogram
        // walk, which does not depend on anything phase P produces.
        const uint32_t p_first = wide ? (b_first & ~31u) : 0u;       // whole warps (the ballots below use full masks)
        if ((uint32_t)tid >= p_first) {
            const unsigned par_mask = MATES == 2 ? 0x55555555u : 0xFFFFFFFFu;         // lanes of mate 0
            for (uint32_t i = (uint32_t)tid - p_first; i < align_up(cnt * MATES, 32); i += blockDim.x - p_first) {
                const uint32_t r = i / MATES;
                const int m = (int)(i % MATES);
                const bool live = r < cnt;
                int cat = SNK_DROP_EMPTY, mask = 0, fsb = -1;
                ReadInfo x;
                x.len = 0; x.clean_len = 0;
                DeltaEnt de[2];
                int nde = 0;
                const uint64_t gi = g0 + r;
                if (live) {
                    const ReadInfo a = reinterpret_cast<const ReadInfo*>(smem + sp.off_info[0])[r];
                    uint32_t err = 0;
                    if (MATES == 2) {
                        const ReadInfo b = reinterpret_cast<const ReadInfo*>(smem + sp.off_info[1])[r];
                        cat = decide_pair(P, a, b, &mask, &fsb);
                        if ((a.flags | b.flags) & RF_BAD_BASE) err |= ERR_BAD_BASE;
                        if ((a.flags | b.flags) & RF_BAD_QUAL) err |= ERR_BAD_QUAL;
                        if (cat == SNK_DROP_LOWQ && ((a.flags | b.flags) & RF_LOWQ_GT1)) err |= ERR_LOWQ_RATIO;
                        x = m ? b : a;
                    } else {
                        cat = P.srna ? decide_srna(P, a, &fsb) : decide_se(P, a, &fsb);
                        mask = cat ? 1 : 0;
                        if (a.flags & RF_BAD_BASE) err |= ERR_BAD_BASE;
                        if (a.flags & RF_BAD_QUAL) err |= ERR_BAD_QUAL;
                        x = a;
                    }
                    unsigned long long* S = A.stats + (size_t)slot * SNK_SLOT_WORDS;
                    if (m == 0) {                                   // once per pair
                        if (err) report_error(A, err, gi);
                        if (fsb >= 0) {
                            atomicAdd(&S[fsb], 1ull);
                            if (MATES == 2) {
                                if (mask & 1) atomicAdd(&S[fsb + 1], 1ull);
                                if (mask & 2) atomicAdd(&S[fsb + 2], 1ull);
                                if (mask == 3) atomicAdd(&S[fsb + 3], 1ull);
                            }
                        }
                    }
                    const uint32_t row0 = r * A.stride;
                    desc[(size_t)m * A.R + r] = hist_desc(x.len, row0, x.flags & RF_QSLOW);
                    nde = delta_entries(x, cat == SNK_KEEP, row0, de);
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
                // per-mate warp aggregation: the lanes of mate m are those with lane % MATES == m
                const unsigned mine = par_mask << m;
                const unsigned lt = ((1u << lane) - 1u) & mine;
                // append the delta entries to the tile's lists: one shared atomic per warp and mate
                const unsigned b1 = __ballot_sync(0xFFFFFFFFu, nde >= 1) & mine, b2 = __ballot_sync(0xFFFFFFFFu, nde >= 2) & mine;
                uint32_t base = 0;
                if (lane == m && b1) base = atomicAdd(&ndelta[m], (uint32_t)(__popc(b1) + __popc(b2)));
                base = __shfl_sync(0xFFFFFFFFu, base, m);
                if (nde >= 1) {
                    DeltaEnt* dl = dlist + (size_t)m * 2u * A.R + base + __popc(b1 & lt) + __popc(b2 & lt);
                    dl[0] = de[0];
                    if (nde >= 2) dl[1] = de[1];
                }
                // last record keys + record counts, one shared atomic per warp and table
                const unsigned lanes_live = __ballot_sync(0xFFFFFFFFu, live) & mine;
                const unsigned kept = __ballot_sync(0xFFFFFFFFu, live && cat == SNK_KEEP) & mine;
                if (lanes_live && lane == 31 - __clz(lanes_live)) {      // last live read of this mate in the warp
                    atomicMax(&lastkey[m], ((gi + 1) << 16) | (unsigned long long)(uint16_t)x.len);
                    atomicAdd(&lastkey[4 + m], (unsigned long long)__popc(lanes_live));
                }
                if (kept && lane == 31 - __clz(kept)) {
                    atomicMax(&lastkey[MATES + m], ((gi + 1) << 16) | (unsigned long long)(uint16_t)x.clean_len);
                    atomicAdd(&lastkey[4 + MATES + m], (unsigned long long)__popc(kept));
                }
            }
        }

        // ---- phase B: per-position histograms, owner computes. Raw items walk every (padded) row of the
        // tile (no dependency on phase P); after a barrier the delta items walk the tile's short list of
        // removed / added record parts.
        const uint8_t* rows_s = smem + sp.off_rows[my_m][0];
        const uint8_t* rows_q = smem + sp.off_rows[my_m][1];
        const bool slow_tile = *tile_slow != 0u;           // written in phase A, CTA-uniform here
        if (!slow_tile) {
            if (q_role) {
                if (wide) unit_q_raw<QCounter, J, J / 2>(rows_q, A.stride, cnt, (int)my_w, (int)half * (J / 2), reinterpret_cast<uint8_t*>(qhist), q_cell0, q_jstep, q_bstep);
                else unit_q_raw<QCounter, J, J>(rows_q, A.stride, cnt, (int)my_w, 0, reinterpret_cast<uint8_t*>(qhist), q_cell0, q_jstep, q_bstep);
            }
            if (b_role) unit_b_raw<J>(rows_s, A.stride, cnt, (int)my_w, wide ? half : 0u, wide ? 2u : 1u, bc);
        }
        __syncthreads();                                    // phase P done: descriptors and delta lists are complete
        if (my_item) {
            const uint32_t nd = ndelta[my_m];
            if (!slow_tile) {
                if (q_role) {
                    if (wide) unit_q_delta<QCounter, J, J / 2>(rows_q, my_dlist, nd, (int)my_w, (int)half * (J / 2), reinterpret_cast<uint8_t*>(qhist), q_cell0_del, q_jstep, q_bstep);
                    else unit_q_delta<QCounter, J, J>(rows_q, my_dlist, nd, (int)my_w, 0, reinterpret_cast<uint8_t*>(qhist), q_cell0_del, q_jstep, q_bstep);
                }
                if (b_role) unit_b_delta<J>(rows_s, my_dlist, nd, (int)my_w, wide ? half : 0u, wide ? 2u : 1u, bc);
            } else {
                unsigned long long* slot_base = A.stats + (size_t)slot * SNK_SLOT_WORDS;
                unsigned long long* f_raw = slot_base + SNK_SLOT_FILE_OFF(file_of_tab(MATES, my_m));
                unsigned long long* f_clean = slot_base + SNK_SLOT_FILE_OFF(file_of_tab(MATES, my_m + MATES));
                uint32_t err = 0;
                if (q_role)
                    err = unit_q_checked<QCounter, J>(rows_q, my_desc, cnt, my_dlist, nd, (int)my_w, wide ? (int)half * (J / 2) : 0, wide ? J / 2 : J,
                                                      P.phred, P.qb, qhist + 2u * item, nraw, (int)A.X, f_raw, f_clean);
                if (b_role) unit_b_checked<J>(rows_s, my_desc, cnt, my_dlist, nd, (int)my_w, wide ? half : 0u, wide ? 2u : 1u, bc);
                if (err) report_error(A, err, g0);
            }
        }
        __syncthreads();
    }
}

#endif // __CUDACC__

} // namespace snkcore
