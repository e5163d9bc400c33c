// dev_params.h — host-side preparation of the device parameter block (shared by engine.cu and the
// CPU replay harness in tests/coretest).
#pragma once
#include <cmath>
#include <climits>
#include <cstring>
#include "filter_core.cuh"

namespace snkcore {

inline int float_to_int_x86(float f)
{
    // the reference assigns float quotients to int; cvttss2si gives INT_MIN for NaN / overflow
    if (!(f == f) || f >= 2147483648.0f || f < -2147483648.0f) return INT_MIN;
    return (int)f;
}

inline void prepare_adapter(const snk_params& p, int mate, int idx, AdapterDev& a)
{
    memset(&a, 0, sizeof(a));
    const int A = p.adapter_len[mate][idx];
    const int adaMis = p.ada_mis[mate], adaEdge = p.ada_edge[mate];
    const float adaMR = p.ada_mr[mate];
    a.len = A;
    memcpy(a.seq, p.adapter[mate][idx], (size_t)A);
    if (A == 0) return;
    const float misGrad5 = (float)((A - 5) / (adaMis + 1));          // read_filter.cpp:714
    const float misGrad = (float)((A - adaEdge) / (adaMis + 1));     // :715
    a.seg_thr = (int)ceilf((float)A * adaMR);                         // :717
    a.budget2 = adaMis;
    a.edge = adaEdge;
    a.n3 = A - adaEdge > 0 ? A - adaEdge : 0;
    for (int r1 = 1; r1 <= 5; r1++) a.budget1[r1 - 1] = float_to_int_x86((float)(A - r1) / misGrad5);   // :724
    for (int r1 = 0; r1 < a.n3 && r1 < SNK_MAX_ADAPTER_LEN; r1++) a.budget3[r1] = float_to_int_x86((float)r1 / misGrad);   // :769
    bool fast = A <= 64;
    for (int i = 0; i < A; i++) {
        const char ch = (char)a.seq[i];
        if (ch != 'A' && ch != 'C' && ch != 'G' && ch != 'T') fast = false;
    }
    a.fast = fast ? 1 : 0;
    int k = A < 32 ? A : 32;
    if (a.seg_thr < k) k = a.seg_thr;     // no run-accept can complete inside the prefilter window
    if (k < 0) k = 0;
    a.pre_k = k;
    a.pre_mask = k >= 32 ? 0xFFFFFFFFu : ((1u << k) - 1u);
    uint64_t b0 = 0, b1 = 0;
    for (int i = 0; i < 64 && i < A; i++) {
        const uint64_t ch = a.seq[i];
        b0 |= ((ch >> 1) & 1ull) << i;
        b1 |= ((ch >> 2) & 1ull) << i;
    }
    a.a0_lo = (uint32_t)b0; a.a0_hi = (uint32_t)(b0 >> 32);
    a.a1_lo = (uint32_t)b1; a.a1_hi = (uint32_t)(b1 >> 32);
}

// hasContam's per-offset tables (read_filter.cpp:603-626, 683-686)
inline void prepare_contam(const snk_params& p, int mate, int idx, ContamDev& k)
{
    memset(&k, 0, sizeof(k));
    const int C = p.contam_len[mate][idx];
    const int adaMis = p.ada_mis[mate], adaEdge = p.ada_edge[mate];      // read 2 is analysed with adaMis2 / adaEdge2 (sequence.cpp:183-188)
    k.len = C;
    memcpy(k.seq, p.contam[mate][idx], (size_t)C);
    if (C == 0) return;
    k.seg_thr = p.contam_seg_thr[mate][idx];
    k.budget2 = adaMis;
    k.edge = adaEdge;
    k.n13 = C - adaEdge > 0 ? C - adaEdge : 0;
    const float misGrad = (float)((C - adaEdge) / (adaMis + 1));
    const float segGrad = (k.seg_thr - 7 + 1 == 0) ? 0.0f : (float)((C - adaEdge) / (k.seg_thr - 7 + 1));
    for (int r1 = 0; r1 < k.n13 && r1 < SNK_MAX_ADAPTER_LEN; r1++) {
        k.mis_t[r1] = float_to_int_x86((float)r1 / misGrad);
        k.seg1_t[r1] = segGrad != 0 ? float_to_int_x86(7 + (float)r1 / segGrad) : 7;
        k.seg3_t[r1] = float_to_int_x86(7 + (float)r1 / segGrad);
    }
}
// all contaminant records of a run, [2][SNK_MAX_CONTAMS]; DevParams::contams must point to a copy the kernel can read
inline void prepare_contams(const snk_params& p, ContamDev* out)
{
    for (int m = 0; m < 2; m++)
        for (int i = 0; i < SNK_MAX_CONTAMS; i++) {
            if (i < p.n_contams[m]) prepare_contam(p, m, i, out[m * SNK_MAX_CONTAMS + i]);
            else memset(&out[m * SNK_MAX_CONTAMS + i], 0, sizeof(ContamDev));
        }
}

// reversecomplementary() of a contaminant (read_filter.cpp:1055-1075); false on an unrecognized base
inline bool revcomp_contam(const char* a, int n, uint8_t* out)
{
    for (int k = 0; k < n; k++) {
        int ch = (unsigned char)a[n - 1 - k];
        if (ch >= 'a' && ch <= 'z') ch -= 32;
        switch (ch) {
            case 'A': out[k] = 'T'; break; case 'T': out[k] = 'A'; break;
            case 'G': out[k] = 'C'; break; case 'C': out[k] = 'G'; break;
            case 'N': out[k] = 'N'; break;
            default: return false;
        }
    }
    return true;
}
inline void prepare_gcontams(const snk_params& p, GContamDev* out)
{
    for (int i = 0; i < SNK_MAX_CONTAMS; i++) {
        memset(&out[i], 0, sizeof(GContamDev));
        if (i >= p.n_gcontams) continue;
        out[i].len = p.gcontam_len[i]; out[i].min_match = p.gcontam_min_match[i]; out[i].mismatch = p.gcontam_mismatch[i];
        memcpy(out[i].fwd, p.gcontam[i], (size_t)p.gcontam_len[i]);
        revcomp_contam(p.gcontam[i], p.gcontam_len[i], out[i].rev);       // validity is checked by params_check
    }
}

inline void prepare_params(const snk_params& p, DevParams& d)
{
    memset(&d, 0, sizeof(d));
    d.contam_discard = p.contam_discard;
    d.n_contams[0] = p.n_contams[0]; d.n_contams[1] = p.n_contams[1];
    d.contams = nullptr;
    d.n_gcontams = p.n_gcontams;
    d.gcontams = nullptr;
    d.is_pe = p.is_pe;
    d.phred = p.quality_phred;
    d.low_qual = p.low_qual;
    d.low_qual_ratio = p.low_qual_ratio;
    d.mean_quality = p.mean_quality;
    d.n_ratio = p.n_ratio; d.highA_ratio = p.highA_ratio; d.polyG_tail = p.polyG_tail;
    d.polyX_num = p.polyX_num;
    d.min_len = p.min_read_length; d.max_len = p.max_read_length;
    d.ada_trim = p.ada_trim;
    d.has_hard = p.has_hard_trim;
    d.has_lq = p.has_trim_bad_head || p.has_trim_bad_tail;
    d.trimming = p.has_hard_trim || d.has_lq || p.index_remove || p.ada_trim || p.contam_trim || p.polyG_tail != -1;
    d.cutback = p.ada_trim || p.contam_trim || p.has_hard_trim || p.has_trim_bad_head || p.has_trim_bad_tail;
    for (int m = 0; m < 2; m++) { d.hard_head[m] = p.hard_head[m]; d.hard_tail[m] = p.hard_tail[m]; }
    d.bad_head_thr = p.has_trim_bad_head ? p.bad_head_thr : 0; d.bad_head_max = p.has_trim_bad_head ? p.bad_head_max : 0;
    d.bad_tail_thr = p.has_trim_bad_tail ? p.bad_tail_thr : 0; d.bad_tail_max = p.has_trim_bad_tail ? p.bad_tail_max : 0;
    d.srna = p.srna;
    d.ada_rctg = p.ada_rctg; d.ada_rma = p.ada_rma; d.ada_rmm = p.ada_rmm;
    d.ada_rar = p.ada_rar; d.ada_rer = p.ada_rer;
    d.n_slots = p.n_slots;
    d.slot_block = p.slot_block;
    int qb = p.max_base_quality + 1;
    if (qb < 1) qb = 1;
    if (qb > SNK_QBINS) qb = SNK_QBINS;
    d.qb = qb;
    for (int m = 0; m < 2; m++) {
        d.n_adapters[m] = p.n_adapters[m];
        for (int i = 0; i < p.n_adapters[m]; i++) prepare_adapter(p, m, i, d.ada[m][i]);
    }
}


} // namespace snkcore
