// filter_core.cuh — per-read / per-position device functions of the filter kernel.
//
// Everything here is `__host__ __device__`: the CUDA kernel (filter_kernel.cu) is the only product
// user, but the same functions also compile as plain C++ so that tests can replay the kernel's
// thread/tile structure on a CPU without a GPU (tests/coretest, TEST-ONLY harness).
//
// What the functions restate (reference file:line):
//   scan_read        stat_read counters + adapter search   read_filter.cpp:80-313, 707-790
//                    followed by fastq_trim                 read_filter.cpp:338-482
//   decide_pair/se   pe_discard / se_discard                sequence.cpp:198-387, 76-178
//   qual_update_* /  stat_pe_fqs / stat_se_fqs tables       peprocess.cpp:1144-1203, seprocess.cpp:683-739
//   base_acc_*
//   trim_stat_*      trimming-position tables               peprocess.cpp:1107-1143, 1325-1360; seprocess.cpp:647-682
#pragma once
#include <stdint.h>
#include <string.h>
#include "../../include/snk_engine.h"

#if defined(__CUDACC__)
#define SNK_HD __host__ __device__ __forceinline__
#define SNK_HD_NOINLINE __host__ __device__ __noinline__
#define SNK_HD_MEMBER static __host__ __device__ __forceinline__
#define SNK_ALIGN16 __align__(16)
#else
#define SNK_HD static inline
#define SNK_HD_NOINLINE static
#define SNK_HD_MEMBER static inline
#define SNK_ALIGN16 alignas(16)
#endif

namespace snkcore {

// ------------------------------------------------------------------ portable intrinsics
SNK_HD uint32_t popc32(uint32_t x)
{
#ifdef __CUDA_ARCH__
    return (uint32_t)__popc(x);
#else
    return (uint32_t)__builtin_popcount(x);
#endif
}
// low 32 bits of (hi:lo) >> sh, sh in [0,31]
SNK_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh)
{
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
#endif
}
// sum of the four bytes of w
SNK_HD uint32_t bytesum(uint32_t w)
{
#ifdef __CUDA_ARCH__
    return __dp4a(w, 0x01010101u, 0u);
#else
    return (w & 0xFF) + ((w >> 8) & 0xFF) + ((w >> 16) & 0xFF) + (w >> 24);
#endif
}
// byte j (0..3) of w, zero extended
SNK_HD uint32_t byte_of(uint32_t w, int j)
{
#ifdef __CUDA_ARCH__
    return __byte_perm(w, 0u, 0x4440u | (uint32_t)j);
#else
    return (w >> (8 * j)) & 0xFFu;
#endif
}
// general byte permute (PRMT, default mode): result byte k = byte (selector nibble k & 7) of the 8-byte
// pool {a (bytes 0-3), b (bytes 4-7)}; selector nibbles must have bit 3 clear
SNK_HD uint32_t perm8(uint32_t a, uint32_t b, uint32_t sel)
{
#ifdef __CUDA_ARCH__
    return __byte_perm(a, b, sel);
#else
    const uint64_t pool = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int k = 0; k < 4; k++) r |= (uint32_t)((pool >> (8 * ((sel >> (4 * k)) & 7u))) & 0xFFu) << (8 * k);
    return r;
#endif
}
// acc + sum over the four byte lanes of a[k] * b[k] (unsigned bytes)
SNK_HD uint32_t dot4(uint32_t a, uint32_t b, uint32_t acc)
{
#ifdef __CUDA_ARCH__
    return __dp4a(a, b, acc);
#else
    for (int k = 0; k < 4; k++) acc += ((a >> (8 * k)) & 0xFFu) * ((b >> (8 * k)) & 0xFFu);
    return acc;
#endif
}
SNK_HD int ctz32(uint32_t x)   // x != 0
{
#ifdef __CUDA_ARCH__
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
SNK_HD int clz32(uint32_t x)   // x != 0
{
#ifdef __CUDA_ARCH__
    return __clz((int)x);
#else
    return __builtin_clz(x);
#endif
}

struct SNK_ALIGN16 U4 { uint32_t x, y, z, w; };
SNK_HD U4 load16(const uint8_t* p)
{
#ifdef __CUDA_ARCH__
    return *reinterpret_cast<const U4*>(p);
#else
    U4 v; memcpy(&v, p, 16); return v;
#endif
}
SNK_HD uint32_t load4(const uint8_t* p)   // p 4-byte aligned
{
#ifdef __CUDA_ARCH__
    return *reinterpret_cast<const uint32_t*>(p);
#else
    uint32_t v; memcpy(&v, p, 4); return v;
#endif
}
SNK_HD void store16(uint8_t* p, const U4& v)   // p 16-byte aligned
{
#ifdef __CUDA_ARCH__
    *reinterpret_cast<U4*>(p) = v;
#else
    memcpy(p, &v, 16);
#endif
}

// ------------------------------------------------------------------ device parameters
// One adapter, preprocessed on the host (engine.cu: prepare_adapter) so that no float arithmetic
// of adapter_pos (read_filter.cpp:714-724,769) runs on the device: the per-offset mismatch budgets
// are evaluated once with the reference's own expression types.
struct AdapterDev {
    int32_t  len;        // adptLen (0 => no match possible)
    int32_t  fast;       // 1: only uppercase A/C/G/T and adptLen <= 64: the bit-plane matcher applies
    int32_t  seg_thr;    // segMatchThr = (int)ceil(adptLen * adaMR)
    int32_t  budget2;    // phase 2 budget = adaMis
    int32_t  edge;       // adaEdge
    int32_t  n3;         // number of phase-3 offsets = max(0, adptLen - adaEdge)
    int32_t  pre_k;      // prefilter window: min(32, adptLen, seg_thr)
    uint32_t pre_mask;   // low pre_k bits
    uint32_t a0_lo, a0_hi, a1_lo, a1_hi;   // bit planes of the adapter's 2-bit codes (A=0,C=1,T=2,G=3), bit c = base c
    int32_t  budget1[5]; // phase 1 budgets for r1 = 1..5
    int32_t  pad_[3];
    int32_t  budget3[SNK_MAX_ADAPTER_LEN];   // phase 3 budget per r1
    uint8_t  seq[SNK_MAX_ADAPTER_LEN];
};

// One contaminant sequence (config keys contam1 / contam2), preprocessed on the host like an adapter: the
// per-offset mismatch budgets and run thresholds of hasContam (read_filter.cpp:596-706) are evaluated once
// with the reference's expression types. The records live in device memory (DevParams::contams).
struct ContamDev {
    int32_t len, seg_thr, budget2, edge;
    int32_t n13;                              // offsets of phases 1 and 3 = max(0, len - edge)
    int32_t pad_[3];
    int32_t mis_t[SNK_MAX_ADAPTER_LEN];       // (int)(r1 / misGrad)
    int32_t seg1_t[SNK_MAX_ADAPTER_LEN];      // phase 1: segGrad != 0 ? 7 + r1/segGrad : 7
    int32_t seg3_t[SNK_MAX_ADAPTER_LEN];      // phase 3: 7 + r1/segGrad (no zero test, read_filter.cpp:686)
    uint8_t seq[SNK_MAX_ADAPTER_LEN];
};

// One global contaminant (config key global_contams): both strands, thresholds of global_contam_pos
struct GContamDev {
    int32_t len, min_match, mismatch, pad_;
    uint8_t fwd[SNK_MAX_ADAPTER_LEN], rev[SNK_MAX_ADAPTER_LEN];
};

// The fields of a `fast` adapter the bit-plane sweep reads, compact enough to keep one per mate in shared
// memory. Budgets are clamped to [-1, 64]: any negative budget behaves like -1 and, windows being at most
// 64 bases, any budget above 64 like 64.
struct AdaHot {
    int32_t  len, seg_thr, budget2, edge;
    uint32_t pre_mask, a0_lo, a0_hi, a1_lo, a1_hi;
    int32_t  budget1[5];
    int32_t  fast, pad_;
    int8_t   budget3[64];
};
SNK_HD int8_t clamp_budget(int32_t b) { return (int8_t)(b < 0 ? -1 : (b > 64 ? 64 : b)); }
SNK_HD void make_ada_hot(const AdapterDev& a, AdaHot& h)
{
    h.len = a.len; h.seg_thr = a.seg_thr; h.budget2 = a.budget2; h.edge = a.edge;
    h.pre_mask = a.pre_mask; h.a0_lo = a.a0_lo; h.a0_hi = a.a0_hi; h.a1_lo = a.a1_lo; h.a1_hi = a.a1_hi;
    for (int i = 0; i < 5; i++) h.budget1[i] = a.budget1[i];
    h.fast = a.fast; h.pad_ = 0;
    for (int i = 0; i < 64; i++) h.budget3[i] = clamp_budget(a.budget3[i]);
}

struct DevParams {
    int32_t is_pe;
    int32_t phred;
    int32_t low_qual;
    float   low_qual_ratio;
    int32_t mean_quality;
    float   n_ratio, highA_ratio, polyG_tail;
    int32_t polyX_num;
    int32_t min_len, max_len;
    int32_t ada_trim;
    int32_t trimming;      // read_filter.cpp:354: any trim option or polyG
    int32_t cutback;       // peprocess.cpp:1441: cut ints copied back to the raw records
    int32_t has_hard, hard_head[2], hard_tail[2];
    int32_t has_lq;        // trimBadHead or trimBadTail given
    int32_t bad_head_thr, bad_head_max, bad_tail_thr, bad_tail_max;
    int32_t n_adapters[2];
    int32_t n_slots;
    int32_t srna;          // filtersRNA: ada[0][0] = 5' adapter, ada[1][0] = 3' adapter (bytes only)
    int32_t ada_rctg, ada_rma, ada_rmm;
    float   ada_rar, ada_rer;
    int32_t qb;            // quality bins kept in shared memory (max_base_quality+1, <= SNK_QBINS)
    int32_t contam_discard, n_contams[2];
    const ContamDev* contams;      // [2][SNK_MAX_CONTAMS], device memory (null when no contaminant is configured)
    int32_t n_gcontams, pad_gc;
    const GContamDev* gcontams;    // [SNK_MAX_CONTAMS], device memory
    int64_t slot_block;
    AdapterDev ada[2][SNK_MAX_ADAPTERS];
};

// per-mate result of scan_read; lives in shared memory between the kernel's phases
struct ReadInfo {
    int16_t len;
    int16_t head_cut, clean_len;
    int16_t head_hdcut, head_lqcut, tail_hdcut, tail_lqcut, adacut_pos;
    uint16_t flags;
};
enum : uint16_t {
    RF_N = 1, RF_HIGHA = 2, RF_POLYX = 4, RF_LOWQ = 8, RF_MEANQ = 16, RF_ADAPTER = 32,
    RF_LOWQ_GT1 = 64, RF_BAD_BASE = 128, RF_BAD_QUAL = 256,
    RF_QSLOW = 512,       // some quality falls outside the shared-memory bins: histogram takes the checked path
    RF_NO3 = 1024, RF_INSNULL = 2048,    // filtersRNA: no 3' adapter / 3' adapter within the first three bases
    RF_TILE = 4096, RF_FOV = 8192,       // the id selected the read for removal (SNK_PRE_TILE / SNK_PRE_FOV of len[])
    RF_CONTAM = 16384,                   // a contaminant sequence of the mate's list was found
    RF_GCONTAM = 32768                   // a global contaminant (either strand) was found
};
SNK_HD uint16_t pre_flags(uint32_t len_word) { return (uint16_t)(((len_word & SNK_PRE_TILE) ? RF_TILE : 0) | ((len_word & SNK_PRE_FOV) ? RF_FOV : 0)); }
enum : uint32_t { ERR_BAD_BASE = 1, ERR_BAD_QUAL = 2, ERR_LOWQ_RATIO = 4, ERR_BAD_LEN = 8 };

// ------------------------------------------------------------------ adapter matching
// Exact restatement of one window of adapter_pos: compare adapter[aoff+c] with read[roff+c] for
// c < winlen, in order; a running match counter that resets on mismatch accepts at seg_thr, the
// (budget+1)-th mismatch rejects, finishing within budget accepts. Positions outside [0,len) of
// the read are mismatches (the reference reads out of bounds there).
SNK_HD bool window_exact(const uint8_t* read, int len, int roff, const uint8_t* ada, int aoff, int winlen,
                         int budget, int seg_thr)
{
    int mis = 0, seg = 0;
    for (int c = 0; c < winlen; c++) {
        int ri = roff + c;
        bool same = (ri >= 0 && ri < len) && read[ri] == ada[aoff + c];
        if (same) { if (++seg >= seg_thr) return true; }
        else { mis++; seg = 0; if (mis > budget) return false; }
    }
    return mis <= budget;
}

SNK_HD int popc64(uint64_t x) { return (int)(popc32((uint32_t)x) + popc32((uint32_t)(x >> 32))); }
SNK_HD int ctz64(uint64_t x)      // x != 0
{
    const uint32_t lo = (uint32_t)x;
    return lo ? ctz32(lo) : 32 + ctz32((uint32_t)(x >> 32));
}

// Decision of one adapter_pos window from its mismatch bit mask (bit c = position c mismatches),
// without walking the window: equivalent to window_exact. The (max(budget,0)+1)-th mismatch is the
// abort position; the first run of seg_thr matches that ends before it accepts; otherwise the
// window accepts iff it finishes with mis <= budget. Loop trip counts depend only on the
// (warp-uniform) parameters, so a warp stays converged. Out of line: it runs only for the few
// offsets that pass both prefilters, and inlining it at every call site of the sweeps multiplies
// the kernel's code size (instruction-cache pressure) for nothing.
SNK_HD_NOINLINE bool window_decide(uint64_t M, int winlen, int budget, int seg_thr)
{
    const uint64_t valid = winlen >= 64 ? ~0ull : ((1ull << winlen) - 1ull);
    M &= valid;
    const int nmis = popc64(M);
    const int nb = budget < 0 ? 0 : budget;
    uint64_t t = M;
    for (int i = 0; i < nb; i++) t &= t - 1;            // drop the nb lowest mismatches
    const int p_abort = t ? ctz64(t) : 64;
    const int T = seg_thr < 1 ? 1 : seg_thr;
    uint64_t run = ~M & valid;                          // bit i survives iff positions i..i+T-1 all match
    if (T > 64) run = 0;
    else for (int k = 1; k < T;) { const int sft = (k < T - k) ? k : T - k; run &= run >> sft; k += sft; }
    const bool run_ok = run != 0 && (ctz64(run) + T - 1) < p_abort;
    return run_ok || nmis <= budget;
}

// ------------------------------------------------------------------ phase A: kNT threads per read
// Phase A splits one read over kNT cooperating threads (adjacent lanes). Each stage below returns
// the calling thread's PART of the result (h = thread's index within the group); the parts are
// merged with merge_*() - on the GPU after a lane shuffle, in the CPU replay by a plain call.
constexpr int kNT = 2;

// 0x80 in every byte lane of w that is < k (bytes are ASCII < 128, k in 1..128)
SNK_HD uint32_t bytes_lt(uint32_t w, uint32_t k)
{
    return ~((w | 0x80808080u) - k * 0x01010101u) & 0x80808080u;
}

// does the low 32 bits of e contain a run of at least T set bits? (T >= 1, warp-uniform)
SNK_HD bool has_run32(uint32_t e, int T)
{
    if (T > 32) return false;
    for (int k = 1; k < T;) { const int sft = (k < T - k) ? k : T - k; e &= e >> sft; k += sft; }
    return e != 0;
}

// Bit planes of a read, 32 bases per word: p0/p1 = the two code bits (A=0,C=1,T=2,G=3 from ASCII bits
// 1,2), pn = N, pl = lowercase. For valid bases (p0,p1,pn,pl) identifies the byte exactly.
template <int NW>
struct ScanPart {
    uint32_t p0[NW], p1[NW], pn[NW], pl[NW];
    uint32_t low128, qsum;                  // 128 x (number of low-quality bases), plain quality byte sum
    uint32_t viol, qbad;                    // OR flags: unrecognized base; quality byte outside [phred, phred+qb)
};
template <int NW>
SNK_HD void merge_scan(ScanPart<NW>& a, const ScanPart<NW>& b)
{
#pragma unroll
    for (int k = 0; k < NW; k++) { a.p0[k] |= b.p0[k]; a.p1[k] |= b.p1[k]; a.pn[k] |= b.pn[k]; a.pl[k] |= b.pl[k]; }
    a.low128 += b.low128; a.qsum += b.qsum;
    a.viol |= b.viol; a.qbad |= b.qbad;
}

// byte lanes 0 .. nv-1 of a word (nv may be <= 0 or >= 4)
SNK_HD uint32_t low_lanes(int nv) { return nv >= 4 ? 0xFFFFFFFFu : (nv <= 0 ? 0u : ((1u << (8 * nv)) - 1u)); }

struct ScanConst { uint32_t phred4, over4, low4, qpad; };
SNK_HD ScanConst scan_const(const DevParams& P)
{
    ScanConst c;
    int low_k = P.low_qual + P.phred + 1;                          // q <= lowQual  <=>  byte < low_k
    low_k = low_k < 0 ? 0 : (low_k > 128 ? 128 : low_k);           // 0: never low, 128: always (bytes are < 128)
    const uint32_t over_k = (uint32_t)(P.qb + P.phred);            // q >= qb  <=>  byte >= over_k  (<= 128)
    c.phred4 = (uint32_t)P.phred * 0x01010101u;
    c.over4 = over_k * 0x01010101u;
    c.low4 = (uint32_t)low_k * 0x01010101u;
    c.qpad = (over_k & 0xFFu) * 0x01010101u;                        // padding quality: the dump bin
    return c;
}

// One 16-byte chunk (4 words of bases sw, 4 of qualities qw), every byte inside the read (a chunk the read ends in
// is padded by the caller with 'A' / phred bytes and corrected afterwards, see scan_chunks). Per word:
//   bases      fold case, look the expected byte up by bits 3..1 (A 000, C 001, T 010, G 011, N 111;
//              100/101/110 -> 0) with one PRMT and compare: exact membership in {A,C,G,T,N,a,c,g,t,n}
//   qualities  three packed subtractions on (q | 0x80) give "q >= phred", "q >= phred+qb", "q >= low_k"
//              per byte lane in bit 7 (no borrow between lanes); DP4A sums the flags and the bytes
// and per pair of words one DP4A per plane gathers bit b of 8 consecutive bytes into 8 adjacent bits.
SNK_HD void scan_chunk16(const uint32_t* sw, const uint32_t* qw, const ScanConst& K, uint32_t& c0, uint32_t& c1, uint32_t& cn,
                         uint32_t& cl, uint32_t& low128, uint32_t& qsum, uint32_t& viol, uint32_t& qbad)
{
    uint32_t f[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        f[k] = sw[k] & 0xDFDFDFDFu;                                  // fold case
        const uint32_t t = (f[k] >> 1) & 0x07070707u;
        const uint32_t u = t | (t >> 4);
        const uint32_t e = perm8(0x47544341u, 0x4E000000u, perm8(u, 0u, 0x4420u));
        viol |= e ^ f[k];
        const uint32_t q = qw[k];
        const uint32_t g = q | 0x80808080u;
        qbad |= (~(g - K.phred4) | (g - K.over4) | q) & 0x80808080u;
        low128 = dot4(~(g - K.low4) & 0x80808080u, 0x01010101u, low128);
        qsum = dot4(q, 0x01010101u, qsum);
    }
#pragma unroll
    for (int pr = 0; pr < 2; pr++) {
        // bits 3..1 of word 2pr in the low nibbles, of word 2pr+1 in the high nibbles; weights 1,2,4,8 turn
        // bit b of the 8 bytes into 8 adjacent bits starting at bit b
        const uint32_t x = (f[2 * pr] & 0x0E0E0E0Eu) | ((f[2 * pr + 1] & 0x0E0E0E0Eu) << 4);
        const uint32_t y = ((sw[2 * pr] >> 4) & 0x02020202u) | (sw[2 * pr + 1] & 0x20202020u);     // bit 5: lowercase
        c0 |= (dot4(x & 0x22222222u, 0x08040201u, 0u) >> 1) << (8 * pr);
        c1 |= (dot4(x & 0x44444444u, 0x08040201u, 0u) >> 2) << (8 * pr);
        cn |= (dot4(x & 0x88888888u, 0x08040201u, 0u) >> 3) << (8 * pr);
        cl |= (dot4(y, 0x08040201u, 0u) >> 1) << (8 * pr);
    }
}

// stage 1: packed scan of this thread's 16-byte chunks (c % kNT == h) of the bases and qualities.
// The scan also normalises the row's padding in place (the tile copy in shared memory, never the
// caller's batch): base bytes at positions >= len become 0 (counted by nobody) and quality bytes
// become phred+qb (the "dump" bin one past the last real bin), so that phase B can walk whole rows of
// the raw tables without per-read length tests. nchunks = stride / 16.
template <int MAXC>
SNK_HD void scan_chunks(uint8_t* seq, uint8_t* qual, int len, int nchunks, const DevParams& P, int h, ScanPart<(MAXC + 1) / 2>& S)
{
    constexpr int NW = (MAXC + 1) / 2;
#pragma unroll
    for (int k = 0; k < NW; k++) { S.p0[k] = 0; S.p1[k] = 0; S.pn[k] = 0; S.pl[k] = 0; }
    S.low128 = S.qsum = 0;
    S.viol = S.qbad = 0;
    const ScanConst K = scan_const(P);
    // Both lanes of a group run the same instruction stream on different data: step cc handles chunk
    // c = kNT*cc + h, i.e. (kNT == 2) lane h fills half h of plane word cc.
    static_assert(kNT == 2, "plane half-word placement below assumes two lanes per read");
#pragma unroll(MAXC <= 16 ? NW : 1)
    for (int cc = 0; cc < NW; cc++) {
        const int c = kNT * cc + h;
        if (c >= nchunks) continue;
        if (16 * c >= len) {                                   // chunk entirely behind the read: padding only
            const U4 z = {0u, 0u, 0u, 0u}, qp = {K.qpad, K.qpad, K.qpad, K.qpad};
            store16(seq + 16 * c, z); store16(qual + 16 * c, qp);
            continue;
        }
        uint32_t c0 = 0, c1 = 0, cn = 0, cl = 0;
        const U4 sv = load16(seq + 16 * c);
        const U4 qv = load16(qual + 16 * c);
        const uint32_t sw[4] = {sv.x, sv.y, sv.z, sv.w};
        const uint32_t qw[4] = {qv.x, qv.y, qv.z, qv.w};
        const int nv = len - 16 * c;
        if (nv >= 16) scan_chunk16(sw, qw, K, c0, c1, cn, cl, S.low128, S.qsum, S.viol, S.qbad);
        else {                                                 // the read ends inside this chunk
            // the tile copy gets its final padding (bases 0, qualities in the dump bin); the scan itself sees 'A'
            // bases (code 00: no plane bit, valid) and phred qualities (in range) behind the read, and the two
            // sums those bytes disturb are corrected afterwards
            uint32_t ss[4], qq[4], sa[4], qa[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t mk = low_lanes(nv - 4 * k);
                ss[k] = sw[k] & mk; qq[k] = (qw[k] & mk) | (K.qpad & ~mk);
                sa[k] = ss[k] | (0x41414141u & ~mk); qa[k] = (qw[k] & mk) | (K.phred4 & ~mk);
            }
            const U4 s4 = {ss[0], ss[1], ss[2], ss[3]}, q4 = {qq[0], qq[1], qq[2], qq[3]};
            store16(seq + 16 * c, s4); store16(qual + 16 * c, q4);
            scan_chunk16(sa, qa, K, c0, c1, cn, cl, S.low128, S.qsum, S.viol, S.qbad);
            const uint32_t pad = (uint32_t)(16 - nv), phred = K.phred4 & 0xFFu;
            S.qsum -= pad * phred;
            if (phred < (K.low4 & 0xFFu)) S.low128 -= 128u * pad;      // a phred byte counts as low quality when low_k > phred
        }
        const int sh = 16 * h;
        S.p0[cc] = c0 << sh; S.p1[cc] = c1 << sh; S.pn[cc] = cn << sh; S.pl[cc] = cl << sh;
    }
}

// mask of the bases of plane word k that lie inside the read
SNK_HD uint32_t plane_valid(int len, int k)
{
    const int nb = len - 32 * k;
    return nb >= 32 ? 0xFFFFFFFFu : (nb <= 0 ? 0u : ((1u << nb) - 1u));
}

// stage 2 (on merged planes): is there a run of >= polyX_num identical consecutive bases?
// (read_filter.cpp:255-268: contig_base >= polyX_num; the first base is compared with 'Q')
template <int NW>
SNK_HD bool polyx_hit(const ScanPart<NW>& S, int len, int polyX_num)
{
    const int runT = polyX_num - 1;         // run of "same as previous base" bits needed
    if (runT <= 0) return true;
    int cur = 0;
    bool hit = false;
    uint32_t c0 = 0, c1 = 0, cn = 0, cl = 0;        // bit 31 of the previous word
#pragma unroll(NW <= 8 ? NW : 1)
    for (int k = 0; k < NW; k++) {
        const uint32_t vm = plane_valid(len, k);
        if (vm == 0) break;
        uint32_t diff = (S.p0[k] ^ ((S.p0[k] << 1) | c0)) | (S.p1[k] ^ ((S.p1[k] << 1) | c1)) |
                        (S.pn[k] ^ ((S.pn[k] << 1) | cn)) | (S.pl[k] ^ ((S.pl[k] << 1) | cl));
        if (k == 0) diff |= 1u;                         // first base never equals the initial 'Q'
        const uint32_t e = ~diff & vm;
        c0 = S.p0[k] >> 31; c1 = S.p1[k] >> 31; cn = S.pn[k] >> 31; cl = S.pl[k] >> 31;
        if (e == vm && vm == 0xFFFFFFFFu) { cur += 32; hit = hit || cur >= runT; }
        else {
            const int nb = 32 - clz32(vm);              // valid bases in this word (vm is a low mask, != 0)
            const int lead = ctz32(~e);
            hit = hit || (cur + lead >= runT) || has_run32(e, runT);
            cur = (e == vm) ? cur + nb : clz32(~(e << (32 - nb)));
        }
    }
    return hit;
}

// stage 3: adapter_pos for one adapter (read_filter.cpp:707-790), this thread's share of the windows.
// p0/p1 = code planes, pb = "can never equal an uppercase A/C/G/T" (N, lowercase, or beyond the read
// end); NW+2 words each (pad words: codes 0, pb all ones).
// Phases 2 and 3 are one forward sweep over window offsets: offsets 0..len-A are phase-2 windows
// (full adapter, budget adaMis, smallest accepted offset wins), offsets len-A+1..len-adaEdge are the
// phase-3 windows (adapter prefix of length len-off hanging over the 3' end; the reference scans them
// shortest-overlap first, so the LARGEST accepted offset wins), and phase 2 beats phase 3.
struct AdaPart { int hit1, pos2, pos3; };
SNK_HD void merge_ada(AdaPart& a, const AdaPart& b)
{
    a.hit1 |= b.hit1;
    if (b.pos2 >= 0 && (a.pos2 < 0 || b.pos2 < a.pos2)) a.pos2 = b.pos2;
    if (b.pos3 > a.pos3) a.pos3 = b.pos3;
}
SNK_HD int ada_result(const AdaPart& a) { return a.hit1 ? 0 : (a.pos2 >= 0 ? a.pos2 : a.pos3); }

template <int NW, class ADA /* AdapterDev or AdaHot */>
SNK_HD void adapter_part(int len, const uint32_t* p0, const uint32_t* p1, const uint32_t* pb, const ADA& a, int h, AdaPart& out)
{
    // The lanes of a group take the window offsets with offset % kNT == h (phase 1: r1 % kNT), so
    // they run the same instruction stream.
    const int A = a.len;
    const uint64_t a0 = ((uint64_t)a.a0_hi << 32) | a.a0_lo, a1 = ((uint64_t)a.a1_hi << 32) | a.a1_lo;
    out.hit1 = 0; out.pos2 = -1; out.pos3 = -1;
    // phase 1: adapter starts r1 = 1..5 bases before the read; read window is bases [0, A-r1)
    {
        const uint64_t r0 = ((uint64_t)p0[1] << 32) | p0[0], r1w = ((uint64_t)p1[1] << 32) | p1[0], rb = ((uint64_t)pb[1] << 32) | pb[0];
        for (int r1 = 1 + h; r1 <= 5; r1 += kNT) {
            const uint64_t M = (r0 ^ (a0 >> r1)) | (r1w ^ (a1 >> r1)) | rb;
            const int budget = a.budget1[r1 - 1];
            const int nb = budget < 0 ? 0 : budget;
            const int wl = A - r1;
            uint32_t pm = a.pre_mask;
            if (wl < 32) pm &= wl <= 0 ? 0u : ((1u << wl) - 1u);
            if ((int)popc32((uint32_t)M & pm) > nb) continue;
            if (window_decide(M, wl, budget, a.seg_thr)) out.hit1 = 1;
        }
    }
    const int last2 = len - A, last3 = len - a.edge;
    // adapter fields used in the sweeps, in registers (the adapter is indexed dynamically in the
    // parameter block, the compiler would otherwise reload them per offset)
    const int budget2 = a.budget2, seg_thr = a.seg_thr, edge = a.edge;
    const int nb2 = budget2 < 0 ? 0 : budget2;
    const uint32_t pm2 = a.pre_mask, a0lo = a.a0_lo, a1lo = a.a1_lo, a0hi = a.a0_hi, a1hi = a.a1_hi;
    // levels 2+3 for one candidate offset of word kw (rare): all planes over the prefilter window, then
    // the exact decision over the whole window
#define SNK_VERIFY(kw_, sft_, x0_, pm_, nb_, winlen_, budget_)                                                                   \
    ([&]() -> bool {                                                                                                             \
        const uint32_t x_ = (x0_) | (funnel_r(p1[kw_], p1[(kw_) + 1], sft_) ^ a1lo) | funnel_r(pb[kw_], pb[(kw_) + 1], sft_);    \
        if ((int)popc32(x_ & (pm_)) > (nb_)) return false;                                                                       \
        const uint32_t xh_ = (funnel_r(p0[(kw_) + 1], p0[(kw_) + 2], sft_) ^ a0hi) | (funnel_r(p1[(kw_) + 1], p1[(kw_) + 2], sft_) ^ a1hi) | \
                             funnel_r(pb[(kw_) + 1], pb[(kw_) + 2], sft_);                                                        \
        return window_decide(((uint64_t)xh_ << 32) | x_, winlen_, budget_, seg_thr);                                             \
    }())
    // ---- phase 2: offsets 0..last2, first accepted offset wins. Level 1 = plane 0 alone (a plane-0
    // difference is a base mismatch); four offsets are tested per loop trip with a single branch.
#pragma unroll(NW <= 8 ? NW : 1)
    for (int kw = 0; kw < NW; kw++) {
        if (32 * kw <= last2 && out.pos2 < 0) {
            const uint32_t l0 = p0[kw], m0 = p0[kw + 1];
            const int send = (last2 - 32 * kw) >= 31 ? 32 : (last2 - 32 * kw + 1);
            int sft = h;
            for (; sft + 3 * kNT < send; sft += 4 * kNT) {
                const uint32_t xa = funnel_r(l0, m0, sft) ^ a0lo, xb = funnel_r(l0, m0, sft + kNT) ^ a0lo;
                const uint32_t xc = funnel_r(l0, m0, sft + 2 * kNT) ^ a0lo, xd = funnel_r(l0, m0, sft + 3 * kNT) ^ a0lo;
                const int ca = (int)popc32(xa & pm2), cb = (int)popc32(xb & pm2), cc = (int)popc32(xc & pm2), cd = (int)popc32(xd & pm2);
                int cmin = ca < cb ? ca : cb;
                const int cmin2 = cc < cd ? cc : cd;
                cmin = cmin < cmin2 ? cmin : cmin2;
                if (cmin > nb2) continue;
                // rare: some of the four offsets passed level 1. One verification site for all of them (in offset
                // order) keeps the sweep's code small
                uint32_t cand = (ca <= nb2 ? 1u : 0u) | (cb <= nb2 ? 2u : 0u) | (cc <= nb2 ? 4u : 0u) | (cd <= nb2 ? 8u : 0u);
                while (cand) {
                    const int off = sft + ctz32(cand) * kNT;
                    cand &= cand - 1u;
                    const uint32_t x0 = funnel_r(l0, m0, off) ^ a0lo;
                    if (SNK_VERIFY(kw, off, x0, pm2, nb2, A, budget2)) { out.pos2 = 32 * kw + off; break; }
                }
                if (out.pos2 >= 0) break;
            }
            if (out.pos2 < 0)
                for (; sft < send; sft += kNT) {
                    const uint32_t x0 = funnel_r(l0, m0, sft) ^ a0lo;
                    if ((int)popc32(x0 & pm2) > nb2) continue;
                    if (SNK_VERIFY(kw, sft, x0, pm2, nb2, A, budget2)) { out.pos2 = 32 * kw + sft; break; }
                }
        }
    }
    // ---- phase 3: offsets last2+1..last3 (window = adapter prefix of length len-off), last accepted wins
    const int first3 = last2 + 1 < 0 ? 0 : last2 + 1;
#pragma unroll(NW <= 8 ? NW : 1)
    for (int kw = 0; kw < NW; kw++) {
        if (32 * kw + 31 >= first3 && 32 * kw <= last3) {
            const uint32_t l0 = p0[kw], m0 = p0[kw + 1];
            int s0 = first3 > 32 * kw ? first3 - 32 * kw : 0;
            s0 += ((s0 % kNT) != h) ? 1 : 0;                       // first offset of my parity (kNT == 2)
            const int send = (last3 - 32 * kw) >= 31 ? 32 : (last3 - 32 * kw + 1);
            // window length and its low-bit mask are carried along (the window shrinks by kNT per step; it is
            // shorter than the adapter, so at most 63 bases)
            int winlen = len - (32 * kw + s0);
            uint64_t wmask = winlen > 0 ? ((1ull << winlen) - 1ull) : 0ull;
            for (int sft = s0; sft < send; sft += kNT, winlen -= kNT, wmask >>= kNT) {
                const int budget = a.budget3[winlen - edge];
                const int nb = budget < 0 ? 0 : budget;
                const uint32_t pm = pm2 & (uint32_t)wmask;
                const uint32_t x0 = funnel_r(l0, m0, sft) ^ a0lo;
                if ((int)popc32(x0 & pm) > nb) continue;
                if (SNK_VERIFY(kw, sft, x0, pm, nb, winlen, budget)) out.pos3 = 32 * kw + sft;
            }
        }
    }
#undef SNK_VERIFY
}

// byte-wise adapter_pos: adapters with N / lowercase / length > 64, and reads shorter than the
// adapter (windows that start before the read). Same phase order as the reference. One thread does it.
SNK_HD_NOINLINE int adapter_pos_bytes(const uint8_t* seq, int len, const AdapterDev& a)
{
    const int A = a.len;
    for (int r1 = 1; r1 <= 5; r1++)
        if (window_exact(seq, len, 0, a.seq, r1, A - r1, a.budget1[r1 - 1], a.seg_thr)) return 0;
    for (int r1 = 0; r1 <= len - A; r1++)
        if (window_exact(seq, len, r1, a.seq, 0, A, a.budget2, a.seg_thr)) return r1;
    for (int r1 = 0; r1 < a.n3; r1++) {
        const int base = len - r1 - a.edge;
        if (window_exact(seq, len, base, a.seq, 0, r1 + a.edge, a.budget3[r1], a.seg_thr)) return base;
    }
    return -1;
}

// ---- contaminant search, hasContam (read_filter.cpp:596-706), byte-wise (an uncommon option). Like
// adapter_pos, with three differences: an uppercase 'N' of the read is neither match nor mismatch (and does
// not break a run), phases 1 and 3 accept runs of seg1_t / seg3_t matches, and their budgets grow with the
// overlap. Read positions outside the read are mismatches.
SNK_HD bool contam_window(const uint8_t* read, int len, int roff, const uint8_t* ct, int coff, int winlen, int budget, int seg_thr)
{
    int mis = 0, seg = 0;
    for (int c = 0; c < winlen; c++) {
        const int ri = roff + c;
        const int rc = (ri >= 0 && ri < len) ? (int)read[ri] : -1;
        if (rc >= 0 && (int)ct[coff + c] == rc) { if (++seg >= seg_thr) return true; }
        else if (rc != 'N') { mis++; seg = 0; if (mis > budget) return false; }
    }
    return mis <= budget;
}
SNK_HD int contam_pos_bytes(const uint8_t* seq, int len, const ContamDev& k)
{
    const int C = k.len;
    if (C == 0) return -1;
    for (int r1 = 0; r1 < k.n13; r1++)
        if (contam_window(seq, len, 0, k.seq, C - r1 - k.edge, r1 + k.edge, k.mis_t[r1], k.seg1_t[r1])) return 0;
    for (int r1 = 0; r1 <= len - C; r1++)
        if (contam_window(seq, len, r1, k.seq, 0, C, k.budget2, k.seg_thr)) return r1;
    for (int r1 = 0; r1 < k.n13; r1++) {
        const int base = len - r1 - k.edge;
        if (contam_window(seq, len, base, k.seq, 0, r1 + k.edge, k.mis_t[r1], k.seg3_t[r1])) return base;
    }
    return -1;
}
// read_filter.cpp:961-1053 global_contam_pos(): scoring walk (match +1, mismatch -200) over three placements of
// the contaminant; score and overlap are reset before each placement only, not between its start positions.
SNK_HD int global_contam_pos(const uint8_t* read, int rl, const uint8_t* ct, int cl, int min_match_len, int mismatch_number)
{
    const int mismatch_score = -200, match_score = 1;
    const int total_mismatch_score = mismatch_number * mismatch_score;
    const int lower_score = (min_match_len - mismatch_number) + total_mismatch_score;
    int total_score = -1000, overlap = 0;
    for (int i = cl - min_match_len; i >= 0; i--) {
        const int j_max = cl - i > rl ? rl : cl - i;
        for (int j = 0; j != j_max; j++) {
            if (read[j] == ct[i + j]) {
                if (total_score > total_mismatch_score) { total_score += match_score; overlap++; }
                else { if (j_max - j < min_match_len) break; total_score = match_score; overlap = 1; }
            } else {
                if (total_score > total_mismatch_score) { total_score += mismatch_score; overlap++; }
                else if (j_max - j < min_match_len) break;
            }
            if (total_score >= lower_score && overlap >= min_match_len) return 0;
        }
    }
    total_score = -1000; overlap = 0;
    for (int i = 0; i <= rl - cl; i++) {
        for (int j = 0; j != cl; j++) {
            if (read[i + j] == ct[j]) {
                if (total_score > total_mismatch_score) { total_score += match_score; overlap++; }
                else { if (cl - j < min_match_len) break; total_score = match_score; overlap = 1; }
            } else {
                if (total_score > total_mismatch_score) { total_score += mismatch_score; overlap++; }
                else if (cl - j < min_match_len) break;
            }
            if (total_score >= lower_score && overlap >= min_match_len) return i + j - overlap + 1;
        }
    }
    total_score = -1000; overlap = 0;
    for (int i = cl > rl ? cl - rl : 0; i <= cl - min_match_len; i++) {
        for (int j = 0; j != cl - i; j++) {
            if (read[rl - (cl - i) + j] == ct[j]) {
                if (total_score > total_mismatch_score) { total_score += match_score; overlap++; }
                else { total_score = match_score; overlap = 1; if (cl - i - j < min_match_len) break; }
            } else {
                if (total_score > total_mismatch_score) { total_score += mismatch_score; overlap++; }
                else if (cl - i - j < min_match_len) break;
            }
            if (total_score >= lower_score && overlap >= min_match_len) return rl - cl + i + j - overlap + 1;
        }
    }
    return -1;
}
// include_global_contam of stat_read (read_filter.cpp:207-249, hasGlobalContams :927-960): either strand of some sequence
SNK_HD bool has_global_contam(const uint8_t* seq, int len, const DevParams& P)
{
    for (int i = 0; i < P.n_gcontams; i++) {
        const GContamDev& g = P.gcontams[i];
        if (global_contam_pos(seq, len, g.fwd, g.len, g.min_match, g.mismatch) >= 0) return true;
        if (global_contam_pos(seq, len, g.rev, g.len, g.min_match, g.mismatch) >= 0) return true;
    }
    return false;
}
SNK_HD bool has_contam(const uint8_t* seq, int len, int mate, const DevParams& P)
{
    for (int i = 0; i < P.n_contams[mate]; i++)
        if (contam_pos_bytes(seq, len, P.contams[mate * SNK_MAX_CONTAMS + i]) >= 0) return true;
    return false;
}

// (out of line like the other cold paths: they must not sit between the hot blocks of phase A)
SNK_HD_NOINLINE uint16_t contam_flags(const uint8_t* seq, int len, int mate, const DevParams& P)
{
    if (P.srna) return 0;
    uint16_t f = 0;
    if (P.n_contams[mate] > 0 && has_contam(seq, len, mate, P)) f |= RF_CONTAM;
    if (P.n_gcontams > 0 && has_global_contam(seq, len, P)) f |= RF_GCONTAM;
    return f;
}

// ---- filtersRNA adapter finders (byte-wise ungapped alignments; reads are short)
// read_filter.cpp:791-862 sRNA_findAdapter: 3' adapter. Alignments in the reference's order: adapter
// offsets 2,1,0 at read position 0, then offset 0 at read positions 1..len-adaRMa; an 'N' of the read is
// neither match nor mismatch; a later accepted alignment replaces the current one only if it has no more
// mismatches and no fewer matches. Returns the read position or -1.
SNK_HD int srna_find_adapter(const uint8_t* read, int len, const uint8_t* ada, int alen, int rma, int rmm, float rer)
{
    if (alen == 0) return -1;
    int start = -1, a1 = 2, mis_best = 0, map_best = 0;
    bool have = false;
    for (int r1 = 0; r1 <= len - rma;) {
        const int l1 = alen - a1, l2 = len - r1;
        const int n = l1 < l2 ? l1 : l2;
        int mis = 0, tot = 0;
        for (int c = 0; c < n; c++) {
            const uint8_t b = read[r1 + c];
            if (b == 'N') continue;
            if (ada[a1 + c] == b) tot++; else mis++;
        }
        if (mis <= rmm && mis + tot >= rma) {
            const float rate = (float)((double)mis / (double)tot);      // `float rate = 1.0*mis/totalMap` (:832)
            if (rate <= rer && (!have || (mis <= mis_best && tot >= map_best))) { start = r1; have = true; mis_best = mis; map_best = tot; }
        }
        if (a1 > 0) a1--; else r1++;
    }
    return start;
}
// read_filter.cpp:863-926 sRNA_hasAdapter: 5' adapter tail. Adapter offsets alen-adaRCtg..0 at read
// position 0, then offset 0 at read positions 1..max(0,len-adaRCtg).
SNK_HD bool srna_has_adapter(const uint8_t* read, int len, const uint8_t* ada, int alen, int rctg, float rar)
{
    if (alen == 0) return false;
    int a1 = alen - rctg;
    if (a1 < 0) a1 = 0;                 // adapters shorter than adaRCtg are rejected by the host (reference: out-of-bounds read)
    const int last = len - rctg < 0 ? 0 : len - rctg;
    for (int r1 = 0; r1 <= last;) {
        const int l1 = alen - a1, l2 = len - r1;
        const int n = l1 < l2 ? l1 : l2;
        int mis = 0, tot = 0, run = 0, best = 0;
        for (int c = 0; c < n; c++) {
            if (ada[a1 + c] == read[r1 + c]) { tot++; run++; best = run > best ? run : best; }
            else { mis++; run = 0; }
        }
        if (mis <= 4 && (best >= rctg || len < 12) &&
            ((double)tot / (double)len >= (double)rar || (double)tot / (double)alen >= (double)rar)) return true;
        if (a1 > 0) a1--; else r1++;
    }
    return false;
}

// ---- bit-plane versions of the two aligners (adapter of uppercase A/C/G/T, at most 64 bases: `fast`).
// q0/q1/qn/ql = the read's planes padded with two zero words. One alignment = a few funnel shifts, one
// XOR/OR tree and popcounts instead of a byte loop: equality with an uppercase adapter base means equal
// 2-bit code, not N, not lowercase.
// window of the read's plane starting at position 32*kw + sft: 32 bases (adapters up to 32 bases) or 64
template <typename W> struct PlaneWin;
template <> struct PlaneWin<uint32_t> {
    SNK_HD_MEMBER uint32_t get(const uint32_t* q, int kw, int sft) { return funnel_r(q[kw], q[kw + 1], (uint32_t)sft); }
    SNK_HD_MEMBER uint32_t low(int n) { return n >= 32 ? 0xFFFFFFFFu : (n <= 0 ? 0u : ((1u << n) - 1u)); }
    SNK_HD_MEMBER int popc(uint32_t x) { return (int)popc32(x); }
    SNK_HD_MEMBER uint32_t ada(uint32_t lo, uint32_t) { return lo; }
};
template <> struct PlaneWin<uint64_t> {
    SNK_HD_MEMBER uint64_t get(const uint32_t* q, int kw, int sft)
    {
        return ((uint64_t)funnel_r(q[kw + 1], q[kw + 2], (uint32_t)sft) << 32) | funnel_r(q[kw], q[kw + 1], (uint32_t)sft);
    }
    SNK_HD_MEMBER uint64_t low(int n) { return n >= 64 ? ~0ull : (n <= 0 ? 0ull : ((1ull << n) - 1ull)); }
    SNK_HD_MEMBER int popc(uint64_t x) { return popc64(x); }
    SNK_HD_MEMBER uint64_t ada(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
};
// does x contain a run of at least T set bits? (T >= 1)
template <typename W>
SNK_HD bool has_run_bits(W x, int T)
{
    if (T > (int)(8 * sizeof(W))) return false;
    for (int k = 1; k < T;) { const int sft = (k < T - k) ? k : T - k; x &= x >> sft; k += sft; }
    return x != 0;
}
// The alignments come in the reference's order: adapter offsets a_first..0 at read position 0, then
// adapter offset 0 at read positions 1..r_last. The position loop is unrolled over plane words (static
// register indices) with a dynamic bit offset inside.
struct SrnaFindState { int start, mis_best, map_best; bool have; };
template <typename W>
SNK_HD void srna_find_eval(SrnaFindState& st, W w0, W w1, W wn, W wl, W A0, W A1, int alen, int a1, int r1, int len, int rma, int rmm, float rer)
{
    const int l1 = alen - a1, l2 = len - r1;
    const W wm = PlaneWin<W>::low(l1 < l2 ? l1 : l2);
    const W neq = (w0 ^ (A0 >> a1)) | (w1 ^ (A1 >> a1)) | wn | wl;
    const int tot = PlaneWin<W>::popc(~neq & wm);
    const int mis = PlaneWin<W>::popc(neq & ~(wn & ~wl) & wm);     // an uppercase 'N' of the read is skipped (:816)
    if (mis <= rmm && mis + tot >= rma) {
        const float rate = (float)((double)mis / (double)tot);      // `float rate = 1.0*mis/totalMap` (:832)
        if (rate <= rer && (!st.have || (mis <= st.mis_best && tot >= st.map_best))) { st.start = r1; st.have = true; st.mis_best = mis; st.map_best = tot; }
    }
}
template <int NW, typename W>
SNK_HD int srna_find_planes(const uint32_t* q0, const uint32_t* q1, const uint32_t* qn, const uint32_t* ql, int len,
                            const AdapterDev& a, int rma, int rmm, float rer)
{
    const int alen = a.len;
    const int r_last = len - rma;
    if (alen == 0 || r_last < 0) return -1;
    const W A0 = PlaneWin<W>::ada(a.a0_lo, a.a0_hi), A1 = PlaneWin<W>::ada(a.a1_lo, a.a1_hi);
    SrnaFindState st = {-1, 0, 0, false};
    {
        const W w0 = PlaneWin<W>::get(q0, 0, 0), w1 = PlaneWin<W>::get(q1, 0, 0), wn = PlaneWin<W>::get(qn, 0, 0), wl = PlaneWin<W>::get(ql, 0, 0);
        for (int a1 = 2; a1 >= 0; a1--) srna_find_eval<W>(st, w0, w1, wn, wl, A0, A1, alen, a1, 0, len, rma, rmm, rer);
    }
#pragma unroll(NW <= 8 ? NW : 1)
    for (int kw = 0; kw < NW; kw++) {
        if (32 * kw <= r_last) {
            const int send = (r_last - 32 * kw) >= 31 ? 32 : (r_last - 32 * kw + 1);
            for (int sft = kw == 0 ? 1 : 0; sft < send; sft++)
                srna_find_eval<W>(st, PlaneWin<W>::get(q0, kw, sft), PlaneWin<W>::get(q1, kw, sft), PlaneWin<W>::get(qn, kw, sft),
                                  PlaneWin<W>::get(ql, kw, sft), A0, A1, alen, 0, 32 * kw + sft, len, rma, rmm, rer);
        }
    }
    return st.start;
}
// min_tot = smallest match count that satisfies one of the two rate tests (see srna_has_planes)
template <typename W>
SNK_HD bool srna_has_eval(W w0, W w1, W wn, W wl, W A0, W A1, int alen, int a1, int r1, int len, int rctg, int min_tot)
{
    const int l1 = alen - a1, l2 = len - r1;
    const W wm = PlaneWin<W>::low(l1 < l2 ? l1 : l2);
    const W neq = (w0 ^ (A0 >> a1)) | (w1 ^ (A1 >> a1)) | wn | wl;
    const W eq = ~neq & wm;
    const int tot = PlaneWin<W>::popc(eq), mis = PlaneWin<W>::popc(neq & wm);
    return mis <= 4 && tot >= min_tot && (len < 12 || has_run_bits<W>(eq, rctg < 1 ? 1 : rctg));
}
template <int NW, typename W>
SNK_HD bool srna_has_planes(const uint32_t* q0, const uint32_t* q1, const uint32_t* qn, const uint32_t* ql, int len,
                            const AdapterDev& a, int rctg, float rar)
{
    const int alen = a.len;
    if (alen == 0) return false;
    // `1.0*totalMap/readLen >= adaRAr || 1.0*totalMap/adptLen >= adaRAr` (:911): with integers below 2^10 and a
    // threshold that is a float, the rounded quotient compares like the exact one, and tot/len >= r is
    // tot >= r*len with the product exact in double; so the test is tot >= ceil(r * min(len, alen)) for
    // r > 0 (r <= 0: always true)
    const int shorter = len < alen ? len : alen;
    const double need = (double)rar * (double)shorter;
    int min_tot = need <= 0.0 ? 0 : (int)need;
    if ((double)min_tot < need) min_tot++;
    const W A0 = PlaneWin<W>::ada(a.a0_lo, a.a0_hi), A1 = PlaneWin<W>::ada(a.a1_lo, a.a1_hi);
    const int a_first = alen - rctg < 0 ? 0 : alen - rctg;
    const int r_last = len - rctg < 0 ? 0 : len - rctg;
    {
        const W w0 = PlaneWin<W>::get(q0, 0, 0), w1 = PlaneWin<W>::get(q1, 0, 0), wn = PlaneWin<W>::get(qn, 0, 0), wl = PlaneWin<W>::get(ql, 0, 0);
        for (int a1 = a_first; a1 >= 0; a1--)
            if (srna_has_eval<W>(w0, w1, wn, wl, A0, A1, alen, a1, 0, len, rctg, min_tot)) return true;
    }
    bool found = false;
#pragma unroll(NW <= 8 ? NW : 1)
    for (int kw = 0; kw < NW; kw++) {
        if (32 * kw <= r_last && !found) {
            const int send = (r_last - 32 * kw) >= 31 ? 32 : (r_last - 32 * kw + 1);
            for (int sft = kw == 0 ? 1 : 0; sft < send; sft++)
                if (srna_has_eval<W>(PlaneWin<W>::get(q0, kw, sft), PlaneWin<W>::get(q1, kw, sft), PlaneWin<W>::get(qn, kw, sft),
                                     PlaneWin<W>::get(ql, kw, sft), A0, A1, alen, 0, 32 * kw + sft, len, rctg, min_tot)) { found = true; break; }
        }
    }
    return found;
}

// stage 4: end scans of fastq_trim (read_filter.cpp:390-429, 454-461): thread 0 scans the head,
// thread kNT-1 the tail and the polyG run. The scans go a word (4 bytes) at a time: `nm` flags (bit 7
// of a byte lane) mark the bytes that END a run; most reads leave after one word.
// 0x80 in every byte lane of w that is NOT a 'G'/'g'
SNK_HD uint32_t not_g_lanes(uint32_t w)
{
    const uint32_t x = (w | 0x20202020u) ^ 0x67676767u;            // zero byte <=> g or G
    return (((x | 0x80808080u) - 0x01010101u) | x) & 0x80808080u;
}
SNK_HD int lead_bytes(uint32_t nm)      // number of byte lanes, from lane 3 downwards, before the first flag
{
    return nm ? clz32(nm) >> 3 : 4;
}
SNK_HD int trail_bytes(uint32_t nm)     // number of byte lanes, from lane 0 upwards, before the first flag
{
    return nm ? ctz32(nm) >> 3 : 4;
}
// run of bytes at the 3' end of row[0..len) whose lanes are not flagged by NM(word); len >= 1
template <class NM>
SNK_HD int tail_run(const uint8_t* row, int len, int limit, NM nm_of)
{
    int run = 0;
    int wi = (len - 1) >> 2, nv = ((len - 1) & 3) + 1;            // word of the last byte, valid bytes in it
    for (;;) {
        const uint32_t nm = nm_of(load4(row + 4 * wi)) << (8 * (4 - nv));   // last valid byte -> lane 3
        int r = lead_bytes(nm);
        if (r > nv) r = nv;
        run += r;
        if (r < nv || wi == 0 || run >= limit) break;
        wi--; nv = 4;
    }
    return run < limit ? run : limit;
}
template <class NM>
SNK_HD int head_run(const uint8_t* row, int len, int limit, NM nm_of)
{
    int run = 0;
    for (int wi = 0;; wi++) {
        const int nv = len - 4 * wi >= 4 ? 4 : len - 4 * wi;
        int r = trail_bytes(nm_of(load4(row + 4 * wi)));
        if (r > nv) r = nv;
        run += r;
        if (r < 4 || 4 * (wi + 1) >= len || run >= limit) break;
    }
    return run < limit ? run : limit;
}
struct TrimPart { int hix, tix, ng; };
SNK_HD void merge_trim(TrimPart& a, const TrimPart& b) { a.hix += b.hix; a.tix += b.tix; a.ng += b.ng; }
// glen = length of the sequence the polyG scan sees (filtersRNA cuts the read at the 3' adapter first)
SNK_HD void trim_part(const uint8_t* seq, const uint8_t* qual, int len, int glen, const DevParams& P, int h, TrimPart& t)
{
    t.hix = t.tix = t.ng = 0;
    if (!P.trimming) return;
    const bool tail = (h == kNT - 1), head = (h == 0);
    if (P.has_lq) {
        // q - phred < thr  <=>  byte < thr + phred (clamped to [0,128]: bytes are < 128)
        if (head && P.bad_head_max > 0) {
            int k = P.bad_head_thr + P.phred; k = k < 0 ? 0 : (k > 128 ? 128 : k);
            const uint32_t kk = (uint32_t)k;
            t.hix = head_run(qual, len, P.bad_head_max, [kk](uint32_t w) { return ~bytes_lt(w, kk) & 0x80808080u; });
        }
        if (tail && P.bad_tail_max > 0) {
            int k = P.bad_tail_thr + P.phred; k = k < 0 ? 0 : (k > 128 ? 128 : k);
            const uint32_t kk = (uint32_t)k;
            t.tix = tail_run(qual, len, P.bad_tail_max, [kk](uint32_t w) { return ~bytes_lt(w, kk) & 0x80808080u; });
        }
    }
    if (tail && P.polyG_tail != -1 && glen > 0) t.ng = tail_run(seq, glen, glen, [](uint32_t w) { return not_g_lanes(w); });
}

// rare path behind ScanPart::qbad: is some quality byte >= 128 or below the Phred base? (the other way to
// be flagged, q >= qb, is legal and only sends the record's histogram through the checked path)
SNK_HD_NOINLINE bool qual_violation(const uint8_t* qual, int len, int phred)
{
    for (int i = 0; i < len; i++)
        if (qual[i] >= 128 || (int)qual[i] < phred) return true;
    return false;
}

// read_filter.cpp:432-438: with adapter trimming on, filtersRNA cuts the read at the 3' adapter (position
// > 2) before the head / tail cuts are applied; returns the sequence length fastq_trim goes on with
SNK_HD int srna_cut_len(const DevParams& P, int ada_pos, int len)
{
    return (P.srna && P.ada_trim && ada_pos > 2 && ada_pos < len) ? ada_pos : len;
}

// stage 5: everything merged -> ReadInfo (predicates of stat_read, read_filter.cpp:289-311, and the
// cut arithmetic of fastq_trim, read_filter.cpp:383-468)
template <int NW>
// ada_pos: adapter_pos() result (filter) or sRNA_findAdapter() result (filtersRNA); has5: sRNA_hasAdapter();
// cur = sequence length fastq_trim applies the cuts to (len, or the 3' adapter position in filtersRNA)
// contam: RF_CONTAM / RF_GCONTAM bits found by the contaminant searches
SNK_HD void finish_read(const ScanPart<NW>& S, bool qviol, bool polyx, int ada_pos, bool has5, uint16_t contam, int cur, const TrimPart& T,
                        int len, int mate, const DevParams& P, ReadInfo& R)
{
    uint16_t flags = 0;
    if (S.viol) flags |= RF_BAD_BASE;
    if (qviol) flags |= RF_BAD_QUAL;
    if (S.qbad) flags |= RF_QSLOW;
    // A and N counts from the planes (only A and N feed predicates: highA, n_ratio)
    // (the planes carry no bit behind the read; a base with no plane bit at all is an A)
    int nN = 0, nA = len;
#pragma unroll(NW <= 8 ? NW : 1)
    for (int k = 0; k < NW; k++) {
        nN += (int)popc32(S.pn[k]);
        nA -= (int)popc32(S.p0[k] | S.p1[k] | S.pn[k]);
    }
    const int nLow = (int)(S.low128 >> 7);
    const int total_q = (int)S.qsum - len * P.phred;
    const float flen = (float)len;
    // float(count)/size with IEEE fp32 division
    const float n_ratio = (float)nN / flen, a_ratio = (float)nA / flen;
    const float lowq_ratio = (float)nLow / flen, mean_q = (float)total_q / flen;
    if (P.n_ratio != -1 && n_ratio >= P.n_ratio) flags |= RF_N;
    if (P.highA_ratio != -1 && a_ratio >= P.highA_ratio) flags |= RF_HIGHA;
    if (polyx) flags |= RF_POLYX;
    flags |= contam;
    if (P.low_qual_ratio != -1 && lowq_ratio >= P.low_qual_ratio) flags |= RF_LOWQ;
    if (lowq_ratio > 1) flags |= RF_LOWQ_GT1;
    if (P.mean_quality != -1 && mean_q < (float)P.mean_quality) flags |= RF_MEANQ;
    int adacut = -1;
    if (P.srna) {
        if (ada_pos == -1) flags |= RF_NO3; else if (ada_pos <= 2) flags |= RF_INSNULL;
        if (has5) flags |= RF_ADAPTER;
    } else if (ada_pos >= 0) { flags |= RF_ADAPTER; adacut = len - ada_pos; }
    int head_hd = -1, head_lq = -1, tail_hd = -1, tail_lq = -1;
    int head_cut = 0, clean_len = len;
    if (P.trimming) {
        int hc = 0, tc = 0;
        if (P.has_hard) { head_hd = P.hard_head[mate]; tail_hd = P.hard_tail[mate]; hc = head_hd; tc = tail_hd; }
        if (P.has_lq) {
            head_lq = T.hix; tail_lq = T.tix;
            if (T.hix > hc) hc = T.hix;
            if (T.tix > tc) tc = T.tix;
        }
        if (P.ada_trim && adacut > 0 && adacut > tc) tc = adacut;
        if (P.polyG_tail != -1 && (float)T.ng >= P.polyG_tail && T.ng > tc) tc = T.ng;
        if (hc + tc > cur) { head_cut = 0; clean_len = 0; }
        else { head_cut = hc; clean_len = cur - hc - tc; }
    }
    R.len = (int16_t)len;
    R.head_cut = (int16_t)head_cut; R.clean_len = (int16_t)clean_len;
    R.head_hdcut = (int16_t)head_hd; R.head_lqcut = (int16_t)head_lq;
    R.tail_hdcut = (int16_t)tail_hd; R.tail_lqcut = (int16_t)tail_lq;
    R.adacut_pos = (int16_t)adacut;
    R.flags = flags;
}

// filtersRNA aligner dispatch: which = 1 -> sRNA_findAdapter with the 3' adapter (ada[1][0]), returns the
// position or -1; which = 0 -> sRNA_hasAdapter with the 5' adapter (ada[0][0]), returns 0/1. Plane
// versions for `fast` adapters, the byte loops otherwise.
template <int NW>
SNK_HD int srna_find(const uint8_t* seq, const uint32_t* q0, const uint32_t* q1, const uint32_t* qn, const uint32_t* ql, int len,
                     const DevParams& P, int which)
{
    const AdapterDev& a = P.ada[which][0];
    const int alen = P.n_adapters[which] > 0 ? a.len : 0;
    if (alen == 0) return which ? -1 : 0;
    if (!a.fast) return which ? srna_find_adapter(seq, len, a.seq, alen, P.ada_rma, P.ada_rmm, P.ada_rer)
                              : (int)srna_has_adapter(seq, len, a.seq, alen, P.ada_rctg, P.ada_rar);
    if (alen <= 32) return which ? srna_find_planes<NW, uint32_t>(q0, q1, qn, ql, len, a, P.ada_rma, P.ada_rmm, P.ada_rer)
                                 : (int)srna_has_planes<NW, uint32_t>(q0, q1, qn, ql, len, a, P.ada_rctg, P.ada_rar);
    return which ? srna_find_planes<NW, uint64_t>(q0, q1, qn, ql, len, a, P.ada_rma, P.ada_rmm, P.ada_rer)
                 : (int)srna_has_planes<NW, uint64_t>(q0, q1, qn, ql, len, a, P.ada_rctg, P.ada_rar);
}

// Exchange policy for the CPU replay / documentation of the protocol: given a callable that returns
// thread h's part, run all kNT parts and merge them. The kernel does the same with lane shuffles.
template <int MAXC>
SNK_HD void scan_read_serial(uint8_t* seq, uint8_t* qual, int len, int nchunks, int mate, const DevParams& P, ReadInfo& R)
{
    constexpr int NW = (MAXC + 1) / 2;
    ScanPart<NW> S, S2;
    scan_chunks<MAXC>(seq, qual, len, nchunks, P, 0, S);
    for (int h = 1; h < kNT; h++) { scan_chunks<MAXC>(seq, qual, len, nchunks, P, h, S2); merge_scan(S, S2); }
    const bool polyx = P.polyX_num != -1 && polyx_hit(S, len, P.polyX_num);
    int ada_pos = -1;
    bool has5 = false;
    if (P.srna) {
        uint32_t q0[NW + 2], q1[NW + 2], qn[NW + 2], ql[NW + 2];
        for (int k = 0; k < NW + 2; k++) { q0[k] = k < NW ? S.p0[k] : 0; q1[k] = k < NW ? S.p1[k] : 0; qn[k] = k < NW ? S.pn[k] : 0; ql[k] = k < NW ? S.pl[k] : 0; }
        ada_pos = srna_find<NW>(seq, q0, q1, qn, ql, len, P, 1);
        has5 = srna_find<NW>(seq, q0, q1, qn, ql, len, P, 0) != 0;
    } else if (P.n_adapters[mate] > 0) {
        uint32_t p0[NW + 2], p1[NW + 2], pb[NW + 2];
        for (int k = 0; k < NW; k++) { p0[k] = S.p0[k]; p1[k] = S.p1[k]; pb[k] = S.pn[k] | S.pl[k] | ~plane_valid(len, k); }
        for (int k = NW; k < NW + 2; k++) { p0[k] = 0; p1[k] = 0; pb[k] = 0xFFFFFFFFu; }
        for (int i = 0; i < P.n_adapters[mate]; i++) {
            const AdapterDev& a = P.ada[mate][i];
            if (a.len == 0) continue;
            if (a.fast && len >= a.len - 1) {
                AdaPart ap, ap2;
                adapter_part<NW>(len, p0, p1, pb, a, 0, ap);
                for (int h = 1; h < kNT; h++) { adapter_part<NW>(len, p0, p1, pb, a, h, ap2); merge_ada(ap, ap2); }
                ada_pos = ada_result(ap);
            } else ada_pos = adapter_pos_bytes(seq, len, a);
            if (ada_pos >= 0) break;
        }
    }
    const int cur = srna_cut_len(P, ada_pos, len);
    TrimPart T, T2;
    trim_part(seq, qual, len, cur, P, 0, T);
    for (int h = 1; h < kNT; h++) { trim_part(seq, qual, len, cur, P, h, T2); merge_trim(T, T2); }
    const uint16_t contam = contam_flags(seq, len, mate, P);
    finish_read<NW>(S, S.qbad && qual_violation(qual, len, P.phred), polyx, ada_pos, has5, contam, cur, T, len, mate, P, R);
}

// ------------------------------------------------------------------ discard cascade
// Returns the category; *mask = pe_dis value; *fs_base = first counter of the category's group
// (or -1). sequence.cpp:198-387 restricted to the categories the engine implements.
SNK_HD int decide_pair(const DevParams& P, const ReadInfo& a, const ReadInfo& b, int* mask, int* fs_base)
{
    bool x, y;
    // sequence.cpp:213-230: tile / fov of fastq1 only; plain counters (fs_base + 1.. are not touched: mask 0)
    if (a.flags & RF_TILE) { *mask = 0; *fs_base = SNK_FS_TILE; return SNK_DROP_TILE; }
    if (a.flags & RF_FOV) { *mask = 0; *fs_base = SNK_FS_FOV; return SNK_DROP_FOV; }
#define SNK_DIS(cat, base) do { if (x || y) { *mask = (x ? 1 : 0) | (y ? 2 : 0); *fs_base = base; return cat; } } while (0)
    if (P.min_len != -1) {
        x = (uint64_t)a.clean_len < (uint64_t)(int64_t)P.min_len; y = (uint64_t)b.clean_len < (uint64_t)(int64_t)P.min_len;
        SNK_DIS(SNK_DROP_SHORT, SNK_FS_SHORT);
    } else if (a.clean_len == 0 || b.clean_len == 0) { *mask = 0; *fs_base = -1; return SNK_DROP_EMPTY; }
    if (P.max_len != -1) {
        x = (uint64_t)a.clean_len > (uint64_t)(int64_t)P.max_len; y = (uint64_t)b.clean_len > (uint64_t)(int64_t)P.max_len;
        SNK_DIS(SNK_DROP_LONG, SNK_FS_LONG);
    }
    if (P.contam_discard) { x = a.flags & RF_GCONTAM; y = b.flags & RF_GCONTAM; SNK_DIS(SNK_DROP_GCONTAM, SNK_FS_GCONTAM); }   // sequence.cpp:262-273
    if (P.contam_discard) { x = a.flags & RF_CONTAM; y = b.flags & RF_CONTAM; SNK_DIS(SNK_DROP_CONTAM, SNK_FS_CONTAM); }   // sequence.cpp:274-288
    x = a.flags & RF_N; y = b.flags & RF_N; SNK_DIS(SNK_DROP_N, SNK_FS_N);
    x = a.flags & RF_HIGHA; y = b.flags & RF_HIGHA; SNK_DIS(SNK_DROP_HIGHA, SNK_FS_HIGHA);
    x = a.flags & RF_POLYX; y = b.flags & RF_POLYX; SNK_DIS(SNK_DROP_POLYX, SNK_FS_POLYX);
    x = a.flags & RF_LOWQ; y = b.flags & RF_LOWQ; SNK_DIS(SNK_DROP_LOWQ, SNK_FS_LOWQ);
    x = a.flags & RF_MEANQ; y = b.flags & RF_MEANQ; SNK_DIS(SNK_DROP_MEANQ, SNK_FS_MEANQ);
    if (!P.ada_trim) { x = a.flags & RF_ADAPTER; y = b.flags & RF_ADAPTER; SNK_DIS(SNK_DROP_ADAPTER, SNK_FS_ADAPTER); }
#undef SNK_DIS
    *mask = 0; *fs_base = -1;
    return SNK_KEEP;
}
// sequence.cpp:76-178
SNK_HD int decide_se(const DevParams& P, const ReadInfo& a, int* fs_base)
{
    if (a.flags & RF_TILE) { *fs_base = SNK_FS_TILE; return SNK_DROP_TILE; }      // sequence.cpp:84-98
    if (a.flags & RF_FOV) { *fs_base = SNK_FS_FOV; return SNK_DROP_FOV; }
    if (P.min_len != -1 && (uint64_t)a.clean_len < (uint64_t)(int64_t)P.min_len) { *fs_base = SNK_FS_SHORT; return SNK_DROP_SHORT; }
    if (P.max_len != -1 && (uint64_t)a.clean_len > (uint64_t)(int64_t)P.max_len) { *fs_base = SNK_FS_LONG; return SNK_DROP_LONG; }
    if (P.contam_discard && (a.flags & RF_CONTAM)) { *fs_base = SNK_FS_CONTAM; return SNK_DROP_CONTAM; }      // sequence.cpp:116-122
    if (P.contam_discard && (a.flags & RF_GCONTAM)) { *fs_base = SNK_FS_GCONTAM; return SNK_DROP_GCONTAM; }   // :123-127
    if (a.flags & RF_N) { *fs_base = SNK_FS_N; return SNK_DROP_N; }
    if (a.flags & RF_HIGHA) { *fs_base = SNK_FS_HIGHA; return SNK_DROP_HIGHA; }
    if (a.flags & RF_POLYX) { *fs_base = SNK_FS_POLYX; return SNK_DROP_POLYX; }
    if (a.flags & RF_LOWQ) { *fs_base = SNK_FS_LOWQ; return SNK_DROP_LOWQ; }
    if (a.flags & RF_MEANQ) { *fs_base = SNK_FS_MEANQ; return SNK_DROP_MEANQ; }
    if ((a.flags & RF_ADAPTER) && !P.ada_trim) { *fs_base = SNK_FS_ADAPTER; return SNK_DROP_ADAPTER; }
    *fs_base = -1;
    return SNK_KEEP;
}

// sequence.cpp:19-75 sRNA_discard
SNK_HD int decide_srna(const DevParams& P, const ReadInfo& a, int* fs_base)
{
    if (P.max_len != -1 && (uint64_t)a.clean_len > (uint64_t)(int64_t)P.max_len) { *fs_base = SNK_FS_LONG; return SNK_DROP_LONG; }
    if (a.flags & RF_LOWQ) { *fs_base = SNK_FS_LOWQ; return SNK_DROP_LOWQ; }
    if (a.flags & RF_NO3) { *fs_base = SNK_FS_NO3ADAPTER; return SNK_DROP_NO3ADAPTER; }
    if (a.flags & RF_INSNULL) { *fs_base = SNK_FS_INSERTNULL; return SNK_DROP_INSERTNULL; }
    if (a.flags & RF_ADAPTER) { *fs_base = SNK_FS_ADAPTER; return SNK_DROP_ADAPTER; }
    if (a.flags & RF_HIGHA) { *fs_base = SNK_FS_HIGHA; return SNK_DROP_HIGHA; }
    if (a.flags & RF_POLYX) { *fs_base = SNK_FS_POLYX; return SNK_DROP_POLYX; }
    if ((uint64_t)a.clean_len < (uint64_t)(int64_t)P.min_len) { *fs_base = SNK_FS_SHORT; return SNK_DROP_SHORT; }   // no -1 test (:68)
    *fs_base = -1;
    return SNK_KEEP;
}

// ------------------------------------------------------------------ trimming-position tables
// Flat index into a file's ts[] block for one record, or -1 (peprocess.cpp:1107-1143).
// which: 0 = PE fq1 (base raw_length), 1 = PE fq2 (base sequence.size()), 2 = SE.
SNK_HD void trim_stat_indices(int which, int slen, int raw_length, int head_hd, int head_lq, int tail_hd,
                              int tail_lq, int adacut, int* head_flat, int* tail_flat)
{
    *head_flat = -1; *tail_flat = -1;
    if (head_hd > 0 || head_lq > 0)
        *head_flat = (head_hd >= head_lq) ? SNK_TS_HT * SNK_MAX_READ_LEN + head_hd : SNK_TS_HLQ * SNK_MAX_READ_LEN + head_lq;
    const bool ada_cond = (which == 2) ? (adacut >= 0) : (adacut > 0);
    if (tail_hd > 0 || tail_lq > 0 || ada_cond) {
        const int base = (which == 1) ? slen : raw_length;
        int flat;
        if (tail_hd >= tail_lq) {
            if (tail_hd >= adacut) flat = SNK_TS_TT * SNK_MAX_READ_LEN + (base - tail_hd + 1);
            else flat = SNK_TS_TA * SNK_MAX_READ_LEN + (base - adacut + 1);
        } else {
            if (tail_lq >= adacut) flat = SNK_TS_TLQ * SNK_MAX_READ_LEN + (base - tail_lq + 1);
            else flat = SNK_TS_TA * SNK_MAX_READ_LEN + (base - adacut + 1);
        }
        if (flat >= 0 && flat < SNK_TS_WORDS) *tail_flat = flat;
    }
    if (*head_flat >= SNK_TS_WORDS) *head_flat = -1;
}

// ------------------------------------------------------------------ per-position histograms
// One histogram item = J consecutive positions of one table; the thread that owns an item is the
// only writer of its counters, so no atomics are needed.
//   quality x position counts: shared memory, 16-bit CounterT cells; cell (q, j) of item x lives in 32-bit word
//                              (q*(J/2) + j/2) * X + x, half j&1 (qcell_index): the lanes of a warp own consecutive items,
//                              so whatever their q rows are they hit 32 different banks (X is a multiple of 32)
//   base x position counts:    packed J x 8 bit per symbol while walking a tile (BaseAcc), then
//                              added to the owner's 5*J register counters (BaseCnt) for the whole launch.
// RAW items count every record of the tile: rows are padded by scan_chunks (bases 0, qualities in the
// dump bin), so the owner walks whole rows with no length test. The CLEAN tables are not counted
// directly: the clean set is the raw set minus the dropped records and the trimmed-off ends, so the
// second group of items accumulates that (usually small) difference, "removed - added", from a
// per-tile list of DELTA entries, and the flush adds raw - delta to the clean tables.
struct BaseAcc { uint32_t a, c, g, t, n; };
// Per-thread base counters for one item, whole flush interval: pairs of 16-bit counters packed in
// 32-bit registers ([symbol A,C,G,T,N][pair p]: sub-position 2p in the low half, 2p+1 in the high
// half). raw counts records (<= kQCounterMax per interval, no carry between halves); del holds
// "removed - added" biased by 0x8000 per half, so that neither adds nor subtracts ever carry or
// borrow across the halves while at most kQCounterMax records are counted per interval.
template <int J> struct BaseCnt { uint32_t raw[5][J / 2], del[5][J / 2]; };
constexpr uint32_t kDelBias = 0x80008000u;
template <int J>
SNK_HD void base_cnt_reset(BaseCnt<J>& c)
{
#pragma unroll
    for (int b = 0; b < 5; b++)
#pragma unroll
        for (int p = 0; p < J / 2; p++) { c.raw[b][p] = 0; c.del[b][p] = kDelBias; }
}

SNK_HD void base_acc_add(BaseAcc& acc, uint32_t s /* masked: bytes outside the item are 0 */)
{
    const uint32_t f = s & 0xDFDFDFDFu;
    const uint32_t m1 = (f >> 1) & 0x01010101u, m2 = (f >> 2) & 0x01010101u, m3 = (f >> 3) & 0x01010101u;
    const uint32_t valid = (f >> 6) & 0x01010101u;
    acc.n += m3;
    acc.c += m1 & ~m2;
    acc.g += m1 & m2 & ~m3;
    acc.t += ~m1 & m2 & ~m3 & 0x01010101u;
    acc.a += valid & ~(m1 | m2 | m3);
}
// bytes 2p and 2p+1 of w, each zero-extended into a 16-bit half
SNK_HD uint32_t byte_pair(uint32_t w, int p)
{
#ifdef __CUDA_ARCH__
    return __byte_perm(w, 0u, 0x4140u + 0x0202u * (uint32_t)p);
#else
    return ((w >> (16 * p)) & 0xFFu) | (((w >> (16 * p + 8)) & 0xFFu) << 16);
#endif
}
// add (sign = +1) or subtract (-1) the packed per-tile counts to/from dst (raw or del) and clear them
template <int J>
SNK_HD void base_acc_spill(BaseAcc& acc, uint32_t (*dst)[J / 2], int sign = 1)
{
    const uint32_t packed[5] = {acc.a, acc.c, acc.g, acc.t, acc.n};
#pragma unroll
    for (int b = 0; b < 5; b++)
#pragma unroll
        for (int p = 0; p < J / 2; p++) dst[b][p] += (uint32_t)sign * byte_pair(packed[b], p);
    acc.a = acc.c = acc.g = acc.t = acc.n = 0;
}

// Loads the J bytes (low J bytes of the result) that item w of a record starting at byte `off` of
// its row block covers (record positions J*w .. J*w+J-1).
template <int J>
SNK_HD uint32_t hist_load_word(const uint8_t* row, int off, int w)
{
    const int byte0 = off + J * w;
    const int al = byte0 & ~3, sh = 8 * (byte0 & 3);
    uint32_t v = load4(row + al);
    if (sh + 8 * J > 32) v = funnel_r(v, load4(row + al + 4), sh);     // the item straddles a word boundary
    else if (sh) v >>= sh;
    return v;
}

// Per-record descriptor, one 32-bit word: record length n (bits 0-9), byte address of the record's
// first base inside the tile's row block = r*stride + first byte (bits 10-30), bit 31 = some quality
// of the record lies outside the shared-memory bins (checked path). 0 = skip.
SNK_HD uint32_t hist_desc(int n, uint32_t addr, bool slow) { return n <= 0 ? 0u : ((uint32_t)n | (addr << 10) | (slow ? 0x80000000u : 0u)); }

// Delta entry: positions [start, end) of the record at `addr` are REMOVED from (or, add flag, ADDED
// to) the clean set relative to the raw set. d0 = hist_desc(end, addr, slow), d1 = start | add << 31.
struct DeltaEnt { uint32_t d0, d1; };
constexpr uint32_t kDeltaAdd = 0x80000000u;
// Entries of one record (0..2): dropped -> remove everything; kept and only 3'-trimmed -> remove the
// tail; kept with a 5' cut (clean positions shift) -> remove everything, add the clean record.
SNK_HD int delta_entries(const ReadInfo& ri, bool keep, uint32_t row0, DeltaEnt* e)
{
    if (ri.len <= 0) return 0;
    const bool slow = (ri.flags & RF_QSLOW) != 0;
    if (keep && ri.head_cut == 0) {
        if (ri.clean_len >= ri.len) return 0;
        e[0].d0 = hist_desc(ri.len, row0, slow); e[0].d1 = (uint32_t)(ri.clean_len < 0 ? 0 : ri.clean_len);
        return 1;
    }
    e[0].d0 = hist_desc(ri.len, row0, slow); e[0].d1 = 0;
    if (!keep || ri.clean_len <= 0) return 1;
    e[1].d0 = hist_desc(ri.clean_len, row0 + (uint32_t)ri.head_cut, slow); e[1].d1 = kDeltaAdd;
    return 2;
}

// byte lanes lo <= j < hi of a word (lo, hi may lie outside [0,4])
SNK_HD uint32_t lane_mask(int lo, int hi)
{
    const uint32_t mh = hi >= 4 ? 0xFFFFFFFFu : (hi <= 0 ? 0u : ((1u << (8 * hi)) - 1u));
    const uint32_t ml = lo >= 4 ? 0xFFFFFFFFu : (lo <= 0 ? 0u : ((1u << (8 * lo)) - 1u));
    return mh & ~ml;
}
template <int J>
SNK_HD void base_update_range(uint32_t s, int lo, int hi, BaseAcc& acc)
{
    const uint32_t jmask = J >= 4 ? 0xFFFFFFFFu : ((1u << (8 * J)) - 1u);
    base_acc_add(acc, s & lane_mask(lo, hi) & jmask);
}

// element index of quality cell (q, j) of item x in a table of row pitch X items
template <int J>
SNK_HD uint32_t qcell_index(uint32_t q, uint32_t j, uint32_t x, uint32_t X) { return ((q * (uint32_t)(J / 2) + (j >> 1)) * X + x) * 2u + (j & 1u); }
// byte offset of sub-position j (relative to sub-position 0 of the same item and q); rowstep = X * 4 bytes
SNK_HD int qj_off(int j, int rowstep) { return (j >> 1) * rowstep + (j & 1) * 2; }
// Quality cells: qcells = the quality table as bytes; the cell of (byte value b, sub-position j) is
// at qcells + cell0 + qj_off(j, rowstep) + b*bstep, where cell0 already folds in the item and the Phred base
// and bstep = (J/2) * rowstep. Units start at an even sub-position, so qj_off(j0 + j) = qj_off(j0) + qj_off(j).
// The J cells belong to different sub-positions, so they never alias: load them all, then store them
// all - one shared-memory round trip per record.
template <typename CounterT, int J>
SNK_HD void qual_update_all(uint32_t q, uint8_t* qcells, int cell0, int rowstep, int bstep)
{
    CounterT* cell[J];
    CounterT val[J];
#pragma unroll
    for (int j = 0; j < J; j++) {
        cell[j] = reinterpret_cast<CounterT*>(qcells + (cell0 + qj_off(j, rowstep)) + (int)byte_of(q, j) * bstep);
        val[j] = *cell[j];
    }
#pragma unroll
    for (int j = 0; j < J; j++) *cell[j] = (CounterT)(val[j] + 1);
}
// sub-positions lo <= j < hi only, inc = +1 / -1 (cells wrap: delta cells are read as signed)
template <typename CounterT, int J>
SNK_HD void qual_update_range(uint32_t q, int lo, int hi, int inc, uint8_t* qcells, int cell0, int rowstep, int bstep)
{
    CounterT* cell[J];
    CounterT val[J];
#pragma unroll
    for (int j = 0; j < J; j++) {
        cell[j] = reinterpret_cast<CounterT*>(qcells + (cell0 + qj_off(j, rowstep)) + (int)byte_of(q, j) * bstep);
        val[j] = (j >= lo && j < hi) ? *cell[j] : (CounterT)0;
    }
#pragma unroll
    for (int j = 0; j < J; j++)
        if (j >= lo && j < hi) *cell[j] = (CounterT)(val[j] + inc);
}

// Checked path (records whose qualities may fall outside [0,qb), RF_QSLOW): sub-positions lo <= j < hi
// of item w of the record at `off`. Bins kept in shared memory take cell_inc; the others go straight
// to the slot's global tables: g_inc into file_base's table and, if given, into mirror_base's too
// (a raw record counts for the clean table as well until a delta entry removes it). Anything outside
// [0,SNK_QBINS) raises the error flag. Returns error bits.
template <typename CounterT, int J>
SNK_HD uint32_t qual_update_checked(const uint8_t* qual, int off, int w, int lo, int hi, int phred, int qb, int cell_inc,
                                    CounterT* qhist /* cell (q 0, j 0) of the item */, int qstride /* row pitch X */, long long g_inc,
                                    unsigned long long* file_base, unsigned long long* mirror_base)
{
    const uint32_t q = hist_load_word<J>(qual, off, w);
    uint32_t err = 0;
#pragma unroll
    for (int j = 0; j < J; j++) {
        if (j >= lo && j < hi) {
            const int qq = (int)((q >> (8 * j)) & 0xFFu) - phred;
            if ((unsigned)qq < (unsigned)qb) {
                CounterT& cell = qhist[qcell_index<J>((uint32_t)qq, (uint32_t)j, 0u, (uint32_t)qstride)];
                cell = (CounterT)(cell + cell_inc);
            }
            else if ((unsigned)qq < (unsigned)SNK_QBINS) {
                const size_t cell = SNK_FILE_QS_OFF + (size_t)(J * w + j) * SNK_QBINS + qq;
                unsigned long long* bases[2] = {file_base, mirror_base};
                for (int t = 0; t < 2; t++) {
                    unsigned long long* F = bases[t];
                    if (!F) continue;
#ifdef __CUDA_ARCH__
                    atomicAdd(F + cell, (unsigned long long)g_inc);
                    if (qq >= 20) atomicAdd(F + SNK_FILE_GS_OFF + SNK_GS_Q20, (unsigned long long)g_inc);
                    if (qq >= 30) atomicAdd(F + SNK_FILE_GS_OFF + SNK_GS_Q30, (unsigned long long)g_inc);
#else
                    F[cell] += (unsigned long long)g_inc;
                    if (qq >= 20) F[SNK_FILE_GS_OFF + SNK_GS_Q20] += (unsigned long long)g_inc;
                    if (qq >= 30) F[SNK_FILE_GS_OFF + SNK_GS_Q30] += (unsigned long long)g_inc;
#endif
                }
            } else err |= ERR_BAD_QUAL;
        }
    }
    return err;
}

// ---- phase B work units. A q-unit owns JN of the J sub-positions of raw item (w of mate m) and of
// its delta item; a b-unit owns the base counters of one item for the records r0, r0+rstep, ...
// (and the delta entries k = r0, r0+rstep, ...). With two units per item every thread of a CTA
// sized for phase A (two threads per read) has phase-B work.
template <typename CounterT, int J, int JN>
SNK_HD void unit_q_raw(const uint8_t* rows_q, uint32_t stride, uint32_t cnt, int w, int j0, uint8_t* qcells, int cell0_raw, int rowstep, int bstep)
{
    const uint8_t* pq = rows_q + J * w;
    const uint32_t sh = 8u * (uint32_t)j0;
    const int c_raw = cell0_raw + qj_off(j0, rowstep);
    // the row word of the next record is loaded before the current one's cells are updated: the compiler cannot move
    // a shared-memory load above the counter stores by itself (it cannot prove that rows and cells do not alias)
    // (the load behind the last record reads the row after the tile's rows: always inside the CTA's shared memory / the
    // replay's buffers, value unused)
    uint32_t w0 = load4(pq);
    for (uint32_t r = 0; r < cnt; r++) {
        const uint32_t w1 = load4(pq + (size_t)(r + 1) * stride);
        qual_update_all<CounterT, JN>(w0 >> sh, qcells, c_raw, rowstep, bstep);
        w0 = w1;
    }
}
template <typename CounterT, int J, int JN>
SNK_HD void unit_q_delta(const uint8_t* rows_q, const DeltaEnt* dl, uint32_t nd, int w, int j0, uint8_t* qcells, int cell0_del, int rowstep, int bstep)
{
    const uint32_t sh = 8u * (uint32_t)j0;
    const int c_del = cell0_del + qj_off(j0, rowstep);
    const int first = J * w + j0;
    for (uint32_t k = 0; k < nd; k++) {
        const DeltaEnt d = dl[k];
        const int hi = (int)(d.d0 & 0x3FFu) - first, lo = (int)(d.d1 & 0x3FFu) - first;
        if (hi <= 0 || lo >= JN) continue;
        qual_update_range<CounterT, JN>(hist_load_word<J>(rows_q, (int)((d.d0 >> 10) & 0x1FFFFFu), w) >> sh, lo, hi,
                                        (d.d1 & kDeltaAdd) ? -1 : 1, qcells, c_del, rowstep, bstep);
    }
}
// delta entries of a b-unit (shared by the fast and the checked path)
template <int J>
SNK_HD void unit_b_delta(const uint8_t* rows_s, const DeltaEnt* dl, uint32_t nd, int w, uint32_t r0, uint32_t rstep, BaseCnt<J>& bc)
{
    BaseAcc acc = {0, 0, 0, 0, 0};
    uint32_t since = 0;
    const int first = J * w;
    for (uint32_t k = r0; k < nd; k += rstep) {
        const DeltaEnt d = dl[k];
        const int hi = (int)(d.d0 & 0x3FFu) - first, lo = (int)(d.d1 & 0x3FFu) - first;
        if (hi <= 0 || lo >= J) continue;
        const uint32_t s = hist_load_word<J>(rows_s, (int)((d.d0 >> 10) & 0x1FFFFFu), w);
        if (d.d1 & kDeltaAdd) {                  // rare: a kept record with a 5' cut
            BaseAcc plus = {0, 0, 0, 0, 0};
            base_update_range<J>(s, lo, hi, plus);
            base_acc_spill<J>(plus, bc.del, -1);
        } else {
            base_update_range<J>(s, lo, hi, acc);
            if (++since == 255) { since = 0; base_acc_spill<J>(acc, bc.del); }       // packed 8-bit lanes must not wrap
        }
    }
    base_acc_spill<J>(acc, bc.del);
}
template <int J>
SNK_HD void unit_b_raw(const uint8_t* rows_s, uint32_t stride, uint32_t cnt, int w, uint32_t r0, uint32_t rstep, BaseCnt<J>& bc)
{
    BaseAcc acc = {0, 0, 0, 0, 0};
    const uint8_t* ps = rows_s + J * w;
    if (cnt <= 255u * rstep) {
        for (uint32_t r = r0; r < cnt; r += rstep) base_acc_add(acc, load4(ps + (size_t)r * stride));
    } else {
        uint32_t since = 0;
        for (uint32_t r = r0; r < cnt; r += rstep) {
            base_acc_add(acc, load4(ps + (size_t)r * stride));
            if (++since == 255) { since = 0; base_acc_spill<J>(acc, bc.raw); }
        }
    }
    base_acc_spill<J>(acc, bc.raw);
}
// Checked variants (some record of the tile has qualities outside the shared-memory bins, or a row was
// not scanned): explicit record lengths from the raw descriptors, out-of-bin qualities straight to
// the slot's global tables. Returns error bits.
template <typename CounterT, int J>
SNK_HD uint32_t unit_q_checked(const uint8_t* rows_q, const uint32_t* desc, uint32_t cnt, const DeltaEnt* dl, uint32_t nd, int w,
                               int j0, int jn, int phred, int qb, CounterT* cell_raw /* cell (q 0, j 0) of the raw item */, uint32_t nraw,
                               int qstride, unsigned long long* f_raw, unsigned long long* f_clean)
{
    uint32_t err = 0;
    const int first = J * w;
    for (uint32_t r = 0; r < cnt; r++) {
        const uint32_t d0 = desc[r];
        int hi = (int)(d0 & 0x3FFu) - first;
        if (hi > j0 + jn) hi = j0 + jn;
        if (hi <= j0) continue;
        err |= qual_update_checked<CounterT, J>(rows_q, (int)((d0 >> 10) & 0x1FFFFFu), w, j0, hi, phred, qb, 1, cell_raw, qstride, 1ll, f_raw, f_clean);
    }
    for (uint32_t k = 0; k < nd; k++) {
        const DeltaEnt d = dl[k];
        int hi = (int)(d.d0 & 0x3FFu) - first, lo = (int)(d.d1 & 0x3FFu) - first;
        if (hi > j0 + jn) hi = j0 + jn;
        if (lo < j0) lo = j0;
        if (hi <= lo) continue;
        const bool add = (d.d1 & kDeltaAdd) != 0;
        err |= qual_update_checked<CounterT, J>(rows_q, (int)((d.d0 >> 10) & 0x1FFFFFu), w, lo, hi, phred, qb, add ? -1 : 1, cell_raw + 2u * nraw,
                                                qstride, add ? 1ll : -1ll, f_clean, nullptr);
    }
    return err;
}
template <int J>
SNK_HD void unit_b_checked(const uint8_t* rows_s, const uint32_t* desc, uint32_t cnt, const DeltaEnt* dl, uint32_t nd, int w,
                           uint32_t r0, uint32_t rstep, BaseCnt<J>& bc)
{
    BaseAcc acc = {0, 0, 0, 0, 0};
    uint32_t since = 0;
    const int first = J * w;
    for (uint32_t r = r0; r < cnt; r += rstep) {
        const uint32_t d0 = desc[r];
        const int hi = (int)(d0 & 0x3FFu) - first;
        if (hi <= 0) continue;
        base_update_range<J>(hist_load_word<J>(rows_s, (int)((d0 >> 10) & 0x1FFFFFu), w), 0, hi, acc);
        if (++since == 255) { since = 0; base_acc_spill<J>(acc, bc.raw); }
    }
    base_acc_spill<J>(acc, bc.raw);
    unit_b_delta<J>(rows_s, dl, nd, w, r0, rstep, bc);
}

// ------------------------------------------------------------------ tile decomposition
// A launch covers reads [first, first+n) of the input. Reads are cut into tiles of at most R reads
// that never straddle a slot-block boundary (blocks of slot_block reads, counted from read 0 of the
// input), so that one tile belongs to exactly one statistics slot.
struct TileMap {
    uint64_t first;      // global index of the batch's first read
    uint32_t n;          // reads in the batch
    uint32_t R;          // tile capacity
    uint64_t sb;         // slot_block
    uint32_t n0;         // reads of the batch that fall into its first (partial) block
    uint32_t tiles0;     // tiles covering those
    uint32_t tpb;        // tiles per full block
    uint32_t ntiles;
};
SNK_HD TileMap make_tile_map(uint64_t first, uint32_t n, uint32_t R, uint64_t sb)
{
    TileMap m;
    m.first = first; m.n = n; m.R = R; m.sb = sb;
    const uint64_t room = sb - first % sb;
    m.n0 = (uint64_t)n < room ? n : (uint32_t)room;
    m.tiles0 = (m.n0 + R - 1) / R;
    const uint64_t tpb = (sb + R - 1) / R;
    m.tpb = tpb > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)tpb;
    const uint32_t rest = n - m.n0;
    const uint64_t full = rest / sb, tail = rest % sb;
    m.ntiles = m.tiles0 + (uint32_t)(full * tpb) + (uint32_t)((tail + R - 1) / R);
    return m;
}
SNK_HD void tile_range(const TileMap& m, uint32_t t, uint32_t* start, uint32_t* cnt)
{
    if (t < m.tiles0) {
        *start = t * m.R;
        const uint32_t left = m.n0 - *start;
        *cnt = left < m.R ? left : m.R;
        return;
    }
    const uint32_t tt = t - m.tiles0;
    const uint32_t blk = tt / m.tpb, j = tt % m.tpb;
    const uint64_t s = (uint64_t)m.n0 + (uint64_t)blk * m.sb + (uint64_t)j * m.R;
    uint64_t c = m.sb - (uint64_t)j * m.R;
    if (c > m.R) c = m.R;
    if (c > m.n - s) c = m.n - s;
    *start = (uint32_t)s; *cnt = (uint32_t)c;
}
SNK_HD int slot_of(uint64_t gi, uint64_t sb, int n_slots) { return (int)((gi / sb) % (uint64_t)n_slots); }

} // namespace snkcore
